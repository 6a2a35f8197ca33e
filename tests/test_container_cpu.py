import numpy as np

from bitswap_b200.container import extract_blocks, unextract_blocks


def test_block_tiling_roundtrip_and_crop():
    rs = np.random.RandomState(0)
    for h, w in ((32, 32), (100, 70), (64, 97), (224, 225)):
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        blocks, hh, ww = extract_blocks(img)
        assert (hh, ww) == (h - h % 32, w - w % 32) and blocks.shape == ((hh // 32) * (ww // 32), 32, 32, 3)
        assert np.array_equal(blocks[1], img[0:32, 32:64]) if ww >= 64 else True      # row-major block order
        assert np.array_equal(unextract_blocks(blocks, hh, ww), img[:hh, :ww])
