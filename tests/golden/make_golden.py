"""Generates tests/golden/* by running the REFERENCE's own classes.

Run in the build container only (needs /root/reference, read-only):
    python tests/golden/make_golden.py
The outputs are committed; nothing at test time reads /root/reference.

What is pinned (SURVEY.md 8c -- the reference has no golden vectors itself):
  ans_kat.json      KAT-1..3 of the reference `ANS` (cifar_compress.py:12-67): tables, push, pop-first,
                    round-trips, underflow behaviour
  tables_small.npz  ANS.__init__ on small float64 pmfs, incl. an argmax tie row
  pmfs_small.npz    logistic_cdf + pmf assembly (rand.py:67-68, cifar_compress.py:182-184) on torch CPU
  bins.npz          rand.Bins / rand.ImageBins outputs (float32 top-level endpoints!)
  model_tiny*.npz   reference Model.infer/generate (model/cifar_train.py, imagenetcrop_train.py) mu/scale
                    for our synthetic state_dict (loaded with load_state_dict(strict=True))
  bitswap_tiny.json state trace of the reference-classes sender/receiver loop on the tiny config
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("BITSWAP_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
tb = types.ModuleType("tensorboardX")
tb.SummaryWriter = object
sys.modules["tensorboardX"] = tb

from utils.torch.rand import Bins, ImageBins, logistic_cdf      # noqa: E402  (reference)
import cifar_compress as RC                                     # noqa: E402  (reference)
from model.cifar_train import Model as RefModel                 # noqa: E402
from model.imagenetcrop_train import Model as RefCropModel      # noqa: E402

from bitswap_b200.config import preset                          # noqa: E402
from bitswap_b200 import synthetic                              # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ANS = RC.ANS


def digest(st):
    b = np.array(st[:-1], dtype=np.uint64).astype(np.uint32).tobytes() + np.array([st[-1]], dtype=np.uint64).tobytes()
    return hashlib.sha256(b).hexdigest()[:16]


def init_state(n):
    st = list(map(int, np.random.RandomState(100).randint(1 << 16, (1 << 32) - 1, size=n, dtype=np.uint32)))
    st[-1] <<= 32
    return st


def ans_kats():
    out = []
    for name, L, S, q, N, seed in [("KAT-1", 64, 16, 4, 32, 0), ("KAT-2", 3072, 256, 8, 4096, 1),
                                   ("KAT-3", 2048, 1024, 10, 4096, 2)]:
        rs = np.random.RandomState(seed)
        pm = rs.dirichlet(np.ones(S) * 0.5, size=L)
        sym = rs.randint(0, S, size=L)
        a = ANS(torch.from_numpy(pm), 31, q)
        st0 = init_state(N)
        pushed = a.encode(st0.copy(), torch.from_numpy(sym))
        popped, psym = a.decode(st0.copy())
        rt1, rsym = a.decode(pushed.copy())
        rt2 = a.encode(popped.copy(), psym)
        assert rt1 == st0 and np.all(rsym.numpy() == sym) and rt2 == st0
        rec = dict(name=name, L=L, S=S, q=q, N=N, seed=seed, init_sha=digest(st0),
                   push_len=len(pushed), push_head=hex(pushed[-1]), push_sha=digest(pushed),
                   pop_len=len(popped), pop_head=hex(popped[-1]), pop_sha=digest(popped),
                   pop_syms_sha=hashlib.sha256(psym.numpy().astype(np.int64).tobytes()).hexdigest()[:16],
                   pop_syms_head=[int(v) for v in psym.numpy()[:6]],
                   P_row0_head=[int(v) for v in a.pmfs[0, :6]],
                   P_sha=hashlib.sha256(a.pmfs.astype(np.int64).tobytes()).hexdigest()[:16],
                   C_sha=hashlib.sha256(a.cdfs.astype(np.int64).tobytes()).hexdigest()[:16])
        if name == "KAT-2":     # underflow: too few initial words -> IndexError from x.pop(-2)
            try:
                a.decode(init_state(512))
                rec["underflow_N512"] = "no error"
            except IndexError:
                rec["underflow_N512"] = "IndexError"
        out.append(rec)
    json.dump(out, open(os.path.join(OUT, "ans_kat.json"), "w"), indent=1)


def tables_small():
    rs = np.random.RandomState(11)
    pm = rs.dirichlet(np.ones(16) * 0.3, size=12)
    pm[3] = 1.0 / 16                       # fully tied row: argmax must pick index 0
    pm[4, :] = 0.0; pm[4, 5] = 0.5; pm[4, 9] = 0.5        # two-way tie at the maximum
    a = ANS(torch.from_numpy(pm), 31, 4)
    pm2 = rs.dirichlet(np.ones(256) * 0.05, size=6)       # very peaked rows: many P == 1
    b = ANS(torch.from_numpy(pm2), 31, 8)
    np.savez_compressed(os.path.join(OUT, "tables_small.npz"), pm_a=pm, P_a=a.pmfs, C_a=a.cdfs,
                        pm_b=pm2, P_b=b.pmfs, C_b=b.cdfs)


def ref_pmfs(endpoints, mu, scale):
    cdfs = logistic_cdf(endpoints.t(), mu, scale).t()
    pmfs = cdfs[:, 1:] - cdfs[:, :-1]
    return torch.cat((cdfs[:, 0].unsqueeze(1), pmfs, 1. - cdfs[:, -1].unsqueeze(1)), dim=1)


def pmfs_small():
    rs = np.random.RandomState(5)
    L, S = 24, 64
    lo, hi = -6 - rs.uniform(0, 1, L), 6 + rs.uniform(0, 1, L)
    ends = torch.from_numpy(np.linspace(lo, hi, S + 1, axis=1)[:, 1:-1])
    mu = torch.from_numpy(rs.normal(0, 2, L).astype(np.float32)).double()
    sc = torch.from_numpy(rs.uniform(0.1, 1.0, L).astype(np.float32)).double()
    pm = ref_pmfs(ends, mu, sc)
    xe = ImageBins(torch.float64, "cpu", 8).endpoints()
    xmu = torch.from_numpy(rs.uniform(-1, 1, 8).astype(np.float32)).double()
    xsc = torch.from_numpy(rs.uniform(0.002, 0.7, 8).astype(np.float32)).double()
    xpm = ref_pmfs(xe, xmu, xsc)
    prior = ref_pmfs(ends, torch.zeros(1, dtype=torch.float64), torch.ones(1, dtype=torch.float64))
    np.savez_compressed(os.path.join(OUT, "pmfs_small.npz"), ends=ends.numpy(), mu=mu.numpy(), sc=sc.numpy(),
                        pm=pm.numpy(), xe=xe.numpy(), xmu=xmu.numpy(), xsc=xsc.numpy(), xpm=xpm.numpy(),
                        prior=prior.numpy())


def bins():
    b = Bins(torch.zeros((1, 1, 4)), torch.ones((1, 1, 4)), 6)
    ib = ImageBins(torch.float64, "cpu", 3)
    b64 = Bins(torch.zeros(2, dtype=torch.float64) + 0.25, torch.ones(2, dtype=torch.float64) * 0.5, 4)
    np.savez_compressed(os.path.join(OUT, "bins.npz"), top_end=b.endpoints().numpy(), top_cen=b.centres().numpy(),
                        img_end=ib.endpoints().numpy(), img_cen=ib.centres().numpy(),
                        b64_end=b64.endpoints().numpy(), b64_cen=b64.centres().numpy())


def ref_model(cfg, sd):
    cls = RefCropModel if cfg.cond_xscale else RefModel
    m = cls(xs=cfg.xs, nz=cfg.nz, zchannels=cfg.zchannels, nprocessing=cfg.nprocessing,
            kernel_size=cfg.kernel_size, resdepth=cfg.resdepth, reswidth=cfg.reswidth, root_process=False)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == \
           [(k, s) for k, s, _ in synthetic.state_dict_spec(cfg)], "state_dict layout mismatch"
    m.load_state_dict(sd, strict=True)
    m.eval()
    m.compress()
    return m


def model_golden(name):
    cfg = preset(name)
    sd = synthetic.synthetic_state_dict(cfg, seed=50, varied=True)
    m = ref_model(cfg, sd)
    zend, zcen = synthetic.synthetic_bins(cfg, seed=0)
    rs = np.random.RandomState(3)
    out = {}
    x = synthetic.synthetic_images(cfg, 1, seed=9)[0].reshape(-1).astype(np.int64)
    xcen = ImageBins(torch.float64, "cpu", cfg.xdim).centres()
    given = xcen[torch.arange(cfg.xdim), torch.from_numpy(x)]
    out["x"] = x.astype(np.uint8)
    with torch.no_grad():
        for i in range(cfg.nz):
            if i > 0:
                sym = rs.randint(0, cfg.zsupport, cfg.zdim)
                out[f"zsym_in_infer{i}"] = sym.astype(np.int16)
                given = zcen[i - 1, torch.arange(cfg.zdim), torch.from_numpy(sym)]
            mu, sc = m.infer(i)(given=given)
            out[f"infer{i}_mu"], out[f"infer{i}_scale"] = mu.numpy(), sc.numpy()
            sym = rs.randint(0, cfg.zsupport, cfg.zdim)
            out[f"zsym_in_gen{i}"] = sym.astype(np.int16)
            z = zcen[i, torch.arange(cfg.zdim), torch.from_numpy(sym)]
            mu, sc = m.generate(i)(given=z)
            out[f"gen{i}_mu"], out[f"gen{i}_scale"] = mu.numpy(), sc.numpy()
    np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), **out)


def bitswap_golden(name, nimg=2, nwords=600):
    """Reference classes + a loop restating cifar_compress.py:175-250 / :283-317 (the script-level
    compress() hard-codes cuda/dataset/checkpoint paths and cannot be called)."""
    cfg = preset(name)
    sd = synthetic.synthetic_state_dict(cfg, seed=50, varied=True)
    m = ref_model(cfg, sd)
    zend, zcen = synthetic.synthetic_bins(cfg, seed=0)
    xb = ImageBins(torch.float64, "cpu", cfg.xdim)
    xend, xcen = xb.endpoints(), xb.centres()
    zr, xr = torch.arange(cfg.zdim), torch.arange(cfg.xdim)
    q, bits = cfg.quantbits, cfg.ansbits
    imgs = synthetic.synthetic_images(cfg, nimg, seed=7)
    w, head = synthetic.initial_words(nwords, seed=100)
    state = [int(v) for v in w] + [head]
    initial = state.copy()
    trace = [("init", len(state), digest(state))]
    zero, one = torch.zeros(1, dtype=torch.float64), torch.ones(1, dtype=torch.float64)
    with torch.no_grad():
        for xi in range(nimg):
            x = torch.from_numpy(imgs[xi].reshape(-1).astype(np.int64))
            for zi in range(cfg.nz):
                inp = zcen[zi - 1, zr, zsym] if zi > 0 else xcen[xr, x]
                mu, scale = m.infer(zi)(given=inp)
                state, zsymtop = ANS(ref_pmfs(zend[zi], mu, scale), bits, q).decode(state)
                trace.append((f"img{xi} pop z{zi+1}", len(state), digest(state)))
                z = zcen[zi, zr, zsymtop]
                mu, scale = m.generate(zi)(given=z)
                pm = ref_pmfs(zend[zi - 1] if zi > 0 else xend, mu, scale)
                state = ANS(pm, bits, q if zi > 0 else 8).encode(state, zsym if zi > 0 else x)
                trace.append((f"img{xi} push {'z%d' % zi if zi > 0 else 'x'}", len(state), digest(state)))
                zsym = zsymtop
            state = ANS(ref_pmfs(zend[-1], zero, one), bits, q).encode(state, zsymtop)
            trace.append((f"img{xi} push prior", len(state), digest(state)))
        final = state.copy()
        # receiver
        for xi in reversed(range(nimg)):
            state, zsymtop = ANS(ref_pmfs(zend[-1], zero, one), bits, q).decode(state)
            for zi in reversed(range(cfg.nz)):
                z = zcen[zi, zr, zsymtop]
                mu, scale = m.generate(zi)(given=z)
                pm = ref_pmfs(zend[zi - 1] if zi > 0 else xend, mu, scale)
                state, sym = ANS(pm, bits, q if zi > 0 else 8).decode(state)
                inp = zcen[zi - 1, zr, sym] if zi > 0 else xcen[xr, sym]
                mu, scale = m.infer(zi)(given=inp)
                state = ANS(ref_pmfs(zend[zi], mu, scale), bits, q).encode(state, zsymtop)
                zsymtop = sym
            assert torch.all(torch.from_numpy(imgs[xi].reshape(-1).astype(np.int64)) == zsymtop)
        assert state == initial
    json.dump(dict(config=name, nimg=nimg, nwords=nwords, trace=trace, final_len=len(final),
                   final_head=hex(final[-1]), final_tail_words=[int(v) for v in final[-9:-1]],
                   net_bits_per_dim=32.0 * (len(final) - len(initial)) / (cfg.xdim * nimg)),
              open(os.path.join(OUT, f"bitswap_{name}.json"), "w"), indent=1)


if __name__ == "__main__":
    torch.set_num_threads(4)
    ans_kats()
    tables_small()
    pmfs_small()
    bins()
    for n in ("tiny", "tiny3"):
        model_golden(n)
        bitswap_golden(n)
    print("golden files written to", OUT)


def logp_golden():
    """Reference log-densities (utils/torch/rand.py:23-64) on random inputs incl. edge pixels and vanishing bins."""
    from utils.torch.rand import logistic_logp, discretized_logistic_logp
    rs = np.random.RandomState(21)
    mu = torch.from_numpy(rs.normal(0, 1.5, (3, 4, 50))); sc = torch.from_numpy(rs.uniform(0.1, 1.0, (3, 4, 50)))
    x = torch.from_numpy(rs.normal(0, 2, (3, 4, 50)))
    lp = logistic_logp(mu, sc, x)
    xm = torch.from_numpy(rs.uniform(-1, 1, (5, 300))); xs = torch.from_numpy(rs.uniform(0.002, 0.7, (5, 300)))
    xx = torch.from_numpy(rs.randint(0, 256, (5, 300)).astype(np.float64)); xx[0, :6] = torch.tensor([0, 255, 0, 255, 1, 254.])
    xs[1, :40] = 1e-4
    dl = discretized_logistic_logp(xm, xs, xx)
    np.savez_compressed(os.path.join(OUT, "logp.npz"), mu=mu.numpy(), sc=sc.numpy(), x=x.numpy(), lp=lp.numpy(),
                        xm=xm.numpy(), xs=xs.numpy(), xx=xx.numpy(), dl=dl.numpy())


if __name__ == "__main__":
    logp_golden()
