"""Debug aid: where do the affine-row kernels (per lanes-per-row setting), the generic two-phase kernels and the fused
exact kernels disagree on a hostile level?  Prints the first differing (stream, row) with its parameters."""
import ctypes
import sys
import os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from bitswap_b200 import synthetic
from bitswap_b200._lib import lib, check, cuda_stream_ptr
from bitswap_b200.streams import StreamSet
import test_ans_gpu as T

dev = "cuda"
L, S, q = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (2048, 1024, 10)))
B = 5
ends, mu, sc = T._level_case(L, S, q, B, seed=17)
sc[:, ::5] = np.float32((2. / 255.) / 8.)
mu[:, 1::9] *= 4
rs = np.random.RandomState(23)
sym = rs.randint(0, S, size=(B, L)).astype(np.int16)
sym[:, :4] = [0, S - 1, 1, S // 2]
states = []
for b in range(B):
    w, head = synthetic.initial_words(7000 + b, seed=900 + b)
    states.append([int(v) for v in w] + [head])

def fused():
    e_pad = torch.from_numpy(T._pad_big(ends)).to(dev)
    dmu, dsc, dsym = torch.from_numpy(mu).to(dev), torch.from_numpy(sc).to(dev), torch.from_numpy(sym).to(dev)
    ss = StreamSet(B, 1 << 14)
    ss.import_lists(states)
    out = torch.zeros((B, L), dtype=torch.int16, device=dev)
    check(lib().bsw_logistic_pop(ss.handle, 0, B, dmu.data_ptr(), L, dsc.data_ptr(), L, e_pad.data_ptr(), S, out.data_ptr(), L, S, 31, q, cuda_stream_ptr()))
    ss.raise_on_error()
    a, popped = ss.export_lists(), out.cpu().numpy().copy()
    check(lib().bsw_logistic_push(ss.handle, 0, B, dmu.data_ptr(), L, dsc.data_ptr(), L, e_pad.data_ptr(), S, dsym.data_ptr(), L, S, 31, q, cuda_stream_ptr()))
    ss.raise_on_error()
    return a, popped, ss.export_lists()

ref = fused()
gen = T._run_2p(ends, mu, sc, S, q, states, sym, mode=0)
def cmp(name, got):
    okp = [got[0][b] == ref[0][b] for b in range(B)]
    okq = [got[2][b] == ref[2][b] for b in range(B)]
    d = np.argwhere(got[1] != ref[1])
    print(f"{name}: pop states equal {okp}, push states equal {okq}, differing popped symbols {len(d)}")
    if len(d):
        # rows are popped from L-1 down: the first divergence in coding order is the LARGEST row index of the stream
        b = d[0][0]
        r = max(x[1] for x in d if x[0] == b)
        print(f"   stream {b}: first divergence at row {r}: mu {mu[b, r]!r} sigma {sc[b, r]!r} got {got[1][b, r]} exact {ref[1][b, r]}; "
              f"grid a {ends[r, 0]!r} d {(ends[r, -1] - ends[r, 0]) / (S - 2)!r}")
cmp("generic two-phase", gen)
for lpr in (32, 8, 4, 2):
    check(lib().bsw_rows6_set_lanes_per_row(lpr))
    check(lib().bsw_rows6_set_verify(1))
    st = np.zeros(4, dtype=np.uint64)
    got = T._run_2p(ends, mu, sc, S, q, states, sym, mode=1)
    check(lib().bsw_rows6_verify_read(st.ctypes.data))
    check(lib().bsw_rows6_set_verify(0))
    print(f"LPR {lpr} verify (LPR 32/4 kernels only): mismatches {int(st[0])}, worst {int(st[1])}/1000 of window, checked {int(st[2])}, exact path {int(st[3])}")
    cmp(f"affine verify-build LPR {lpr}", got)
    got = T._run_2p(ends, mu, sc, S, q, states, sym, mode=1)
    cmp(f"affine LPR {lpr}", got)
check(lib().bsw_rows6_set_lanes_per_row(0))
