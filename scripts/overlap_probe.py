"""Experiment: lanes (sub-batches on separate streams) x dual-stream priority chaining inside each codec."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bitswap_b200 import synthetic
from bitswap_b200.config import preset
from bitswap_b200.codec import PipelinedCodec, Bins
from bitswap_b200.streams import StreamSet

cfg = preset("cifar8"); B = 1024
sd = synthetic.synthetic_state_dict(cfg, seed=50, varied=False)
zend, zcen = synthetic.synthetic_bins(cfg, seed=0)
bins = Bins(cfg, zend, zcen)
x = torch.from_numpy(synthetic.synthetic_images(cfg, B, seed=7)).cuda()
w, head = synthetic.initial_words(4096, seed=100)
for lanes, dual in [(4, 0), (6, 0), (8, 0)]:
    ss = StreamSet(B, 6144); ss.fill(w, head)
    pc = PipelinedCodec(cfg, sd, bins, B, lanes=lanes)
    pc.set_dual_stream(bool(dual))
    out = torch.empty_like(x)
    for _ in range(2):
        pc.encode(ss, x); pc.decode(ss, B, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        pc.encode(ss, x); pc.decode(ss, B, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ss.raise_on_error()
    print(f"lanes={lanes} dual={dual}: {ms:.1f} ms/step -> {B*1024/ms/1e3:.3f} Mpixel/s enc+dec, roundtrip {torch.equal(out, x)}", flush=True)
    del pc
