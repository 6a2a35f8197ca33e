"""Diagnostic: step the Bit-Swap sender level by level with the low-level C-ABI calls and compare
against the golden reference trace + the torch oracle's nets."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from bitswap_b200 import synthetic
from bitswap_b200.config import preset
from bitswap_b200.model import Model
from bitswap_b200.codec import Bins
from bitswap_b200.streams import StreamSet
from bitswap_b200._lib import lib, check, cuda_stream_ptr
import ctypes

name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
g = json.load(open(f"tests/golden/bitswap_{name}.json"))
cfg = preset(name)
sd = synthetic.synthetic_state_dict(cfg, seed=50, varied=True)
m = Model.from_config(cfg, max_batch=1).load_state_dict(sd); m.compress()
orc = O.ModelOracle(cfg, sd)
zend, zcen = synthetic.synthetic_bins(cfg, seed=0)
bins = Bins(cfg, zend, zcen)
ss = StreamSet(1, 8192)
w, head = synthetic.initial_words(g["nwords"], seed=100)
ss.fill(w, head)
imgs = synthetic.synthetic_images(cfg, g["nimg"], seed=7)
S, q = cfg.zsupport, cfg.quantbits
def ptrs(level):
    a, b, c = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    check(lib().bsw_bins_device_ptrs(bins.handle, level, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
    return a.value, b.value, c.value
ti = 1
def note(tag):
    global ti
    n, h, f = ss.sizes()
    st = ss.export_lists()[0]
    want = g["trace"][ti]
    ok = (len(st) == want[1], O.CState.from_list(st).digest() == want[2])
    print(f"{tag:24s} len {len(st)} want {want[1]} {'OK' if ok[0] else 'LEN-DIFF'} {'sha OK' if ok[1] else 'sha diff'}")
    ti += 1
zr = torch.arange(cfg.zdim)
for xi in range(g["nimg"]):
    x = torch.from_numpy(imgs[xi].reshape(-1).astype(np.int64))
    xc = ((x.double() - 127.5) / 127.5)
    zsym = None
    for zi in range(cfg.nz):
        given = zcen[zi - 1, zr, zsym.long().cpu()] if zi > 0 else xc
        mu, sc = m.infer(zi)(given.cuda())
        mo, so = orc.infer(zi)(given.unsqueeze(0))
        print(f"  infer{zi}: |dmu| {float((mu.cpu()-mo[0]).abs().max()):.2e} |dsc| {float((sc.cpu()-so[0]).abs().max()):.2e}")
        mu32, sc32 = mu.float().contiguous(), sc.float().contiguous()
        out = torch.zeros(cfg.zdim, dtype=torch.int16, device="cuda")
        ze, _, xe = ptrs(zi)
        check(lib().bsw_logistic_pop(ss.handle, 0, 1, mu32.data_ptr(), cfg.zdim, sc32.data_ptr(), cfg.zdim, ze, S, out.data_ptr(), cfg.zdim, S, 31, q, cuda_stream_ptr()))
        note(f"img{xi} pop z{zi+1}")
        z = zcen[zi, zr, out.long().cpu()]
        mu, sc = m.generate(zi)(z.cuda())
        mo, so = orc.generate(zi)(z.unsqueeze(0))
        print(f"  gen{zi}: |dmu| {float((mu.cpu()-mo[0]).abs().max()):.2e} |dsc| {float((sc.cpu()-so[0]).abs().max()):.2e}")
        mu32, sc32 = mu.float().contiguous(), sc.float().contiguous()
        if zi > 0:
            zl, _, _ = ptrs(zi - 1)
            sy = zsym.to(torch.int16).contiguous()
            check(lib().bsw_logistic_push(ss.handle, 0, 1, mu32.data_ptr(), cfg.zdim, sc32.data_ptr(), cfg.zdim, zl, S, sy.data_ptr(), cfg.zdim, S, 31, q, cuda_stream_ptr()))
        else:
            sy = x.to(torch.int16).cuda().contiguous()
            check(lib().bsw_logistic_push(ss.handle, 0, 1, mu32.data_ptr(), cfg.xdim, sc32.data_ptr(), cfg.xdim, xe, 0, sy.data_ptr(), cfg.xdim, 256, 31, 8, cuda_stream_ptr()))
        note(f"img{xi} push {'z%d' % zi if zi else 'x'}")
        zsym = out
    # prior via materialised tables
    P = torch.empty((cfg.zdim, S), dtype=torch.int32, device="cuda"); C = torch.empty((cfg.zdim, S + 1), dtype=torch.int32, device="cuda")
    zero, one = torch.zeros(1, dtype=torch.float64, device="cuda"), torch.ones(1, dtype=torch.float64, device="cuda")
    e_raw = zend[-1].contiguous().cuda()
    check(lib().bsw_logistic_tables(e_raw.data_ptr(), S - 1, zero.data_ptr(), one.data_ptr(), 0, cfg.zdim, S, 31, q, P.data_ptr(), C.data_ptr(), cuda_stream_ptr()))
    s32 = zsym.to(torch.int32).contiguous()
    check(lib().bsw_ans_push(ss.handle, 0, 1, P.data_ptr(), C.data_ptr(), 0, 0, s32.data_ptr(), cfg.zdim, S, 31, cuda_stream_ptr()))
    note(f"img{xi} push prior")
