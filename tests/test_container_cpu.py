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


def test_container_headers_are_validated_before_any_device_work():
    """demo_decompress.py:216-222 trusts its input file; here a truncated or inconsistent container is a ValueError raised while
    the headers are parsed -- before a stream set is sized from them or anything is launched (so no codec is needed)."""
    import pytest
    from bitswap_b200.container import decompress_images

    good_tail = [1, 0, 6, 64, 96]                             # head_lo, head_hi, nblocks, h, w: 2 x 3 blocks
    bad = [np.zeros(4, np.uint32),                             # shorter than the tail itself
           np.array([7, 7] + [1, 0, 5, 64, 96], np.uint32),    # nblocks does not match h x w
           np.array([7, 7] + [1, 0, 6, 64, 100], np.uint32),   # w not a multiple of 32
           np.array([7, 7] + [1, 0, 0, 0, 96], np.uint32),     # empty image
           np.array([7, 7] + [1, 0, 1 << 21, 32 << 11, 32 << 10], np.uint32),   # consistent but absurdly large
           np.zeros((2, 8), np.uint32)]                        # not a flat word array
    for c in bad:
        with pytest.raises(ValueError):
            decompress_images(None, [np.array([7, 7] + good_tail, np.uint32), c])
