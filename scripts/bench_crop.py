"""BASELINE.json configs[4] shape on one GPU: 100 variable-size synthetic images coded as chained 32x32 block streams
with the imagenetcrop model family (nz=4, W=256, conditional x-scale), through bitswap_b200.container, next to
gzip / bz2 / lzma / PNG / WebP on the host (benchmark_compress.py:64-103).  Synthetic smooth images, random-init weights:
the rates are NOT the paper's, the point is the path and its throughput."""
import bz2, gzip, io, json, lzma, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bitswap_b200 import synthetic
from bitswap_b200.config import preset
from bitswap_b200.codec import PipelinedCodec, Bins
from bitswap_b200.container import compress_images, decompress_images

cfg = preset("imagenetcrop4")
rs = np.random.RandomState(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
images = []
for i in range(n):
    h, w = rs.randint(224, 513, 2)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 60 * np.sin(xx * rs.uniform(0.01, 0.1) + yy * rs.uniform(0.01, 0.1) + rs.uniform(0, 6)) +
                    rs.normal(0, 6, (h, w)) for _ in range(3)], axis=-1)
    images.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
sd = synthetic.synthetic_state_dict(cfg, seed=50, varied=False)
zend, zcen = synthetic.synthetic_bins(cfg, seed=0)
codec = PipelinedCodec(cfg, sd, Bins(cfg, zend, zcen), n, lanes=4)
t0 = time.perf_counter(); conts = compress_images(codec, images); torch.cuda.synchronize(); t1 = time.perf_counter()
back = decompress_images(codec, conts); torch.cuda.synchronize(); t2 = time.perf_counter()
crop = [im[:im.shape[0] - im.shape[0] % 32, :im.shape[1] - im.shape[1] % 32] for im in images]
assert all(np.array_equal(a, b) for a, b in zip(crop, back))
dims = sum(c.size for c in crop)
blocks = sum(c.shape[0] * c.shape[1] // 1024 for c in crop)
res = {"images": n, "blocks": blocks, "Mpixel": dims / 3e6, "encode_s": t1 - t0, "decode_s": t2 - t1,
       "encode_Mpixel_s": dims / 3e6 / (t1 - t0), "decode_Mpixel_s": dims / 3e6 / (t2 - t1),
       "bitswap_bits_per_dim_incl_trimmed_initial_bits": sum(32 * (len(c) - 3) for c in conts) / dims, "roundtrip_ok": True}
raw = [c.tobytes() for c in crop]
res["gzip"] = sum(8 * len(gzip.compress(r)) for r in raw) / dims
res["bz2"] = sum(8 * len(bz2.compress(r)) for r in raw) / dims
res["lzma"] = sum(8 * len(lzma.compress(r)) for r in raw) / dims
try:
    import PIL.Image as pimg
    def enc(c, fmt, **kw):
        b = io.BytesIO(); pimg.fromarray(c).save(b, format=fmt, **kw); return 8 * len(b.getvalue())
    res["png"] = sum(enc(c, "PNG", optimize=True) for c in crop) / dims
    res["webp"] = sum(enc(c, "WebP", lossless=True, quality=100) for c in crop) / dims
except Exception as e:
    res["pil"] = str(e)
print(json.dumps(res))
