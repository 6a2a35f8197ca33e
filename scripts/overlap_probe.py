"""Experiment: two half-batch codecs on two CUDA streams (conv on tensor pipe || row tables on FP64 pipe)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bitswap_b200 import synthetic
from bitswap_b200.config import preset
from bitswap_b200.model import Model
from bitswap_b200.codec import BitSwapCodec, Bins
from bitswap_b200.streams import StreamSet

cfg = preset("cifar8"); B = 1024
sd = synthetic.synthetic_state_dict(cfg, seed=50, varied=False)
zend, zcen = synthetic.synthetic_bins(cfg, seed=0)
bins = Bins(cfg, zend, zcen)
x = torch.from_numpy(synthetic.synthetic_images(cfg, B, seed=7)).cuda()
w, head = synthetic.initial_words(4096, seed=100)

def run(nsplit, prio):
    ss = StreamSet(B, 6144); ss.fill(w, head)
    n = B // nsplit
    codecs, streams = [], []
    for i in range(nsplit):
        m = Model.from_config(cfg, max_batch=n, use_tensor_cores=True).load_state_dict(sd); m.compress()
        codecs.append(BitSwapCodec(cfg, m, bins, n))
        streams.append(torch.cuda.Stream(priority=(-1 if (prio and i % 2) else 0)))
    out = torch.empty_like(x)
    def step():
        ev = torch.cuda.Event(); ev.record()
        for i in range(nsplit):
            streams[i].wait_event(ev)
            with torch.cuda.stream(streams[i]):
                codecs[i].encode(ss, x[i*n:(i+1)*n], first=i*n)
                codecs[i].decode(ss, n, first=i*n, out=out[i*n:(i+1)*n])
        for s in streams: torch.cuda.current_stream().wait_stream(s)
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ok = torch.equal(out, x)
    print(f"nsplit={nsplit} prio={prio}: {ms:.1f} ms/step -> {B*1024/ms/1e3:.3f} Mpixel/s enc+dec, roundtrip {ok}", flush=True)
    del codecs

for nsplit, prio in [(4, 0), (8, 0), (16, 0)]:
    run(nsplit, prio)
