"""Deterministic synthetic assets (weights, bins, images, initial ANS words).

The reference's checkpoints, discretisation bins and datasets are downloads
(README.md:123-135) that do not exist offline, so every parity and bench run
uses the seeded assets built here.  All generators use numpy RandomState so the
same bytes come out on the build container and on the GPU box.

Formats follow the reference:
  * state_dict keys/shapes  -- model/cifar_train.py:86-308
  * zendpoints [nz,zdim,2^q-1], zcentres [nz,zdim,2^q] f64 -- discretization.py:21-27,81-99
  * initial ANS words       -- cifar_compress.py:157-158
"""
import numpy as np
import torch

from .config import CodecConfig


def state_dict_spec(cfg: CodecConfig):
    """Ordered (key, shape, kind) list of the reference Model.state_dict().
    kind in {'v', 'gain_log', 'gain_lin', 'b', 'gen_std'}; 'gain_lin' marks
    ResNetLayer.conv2 (loggain=False, utils/torch/modules.py:226-227)."""
    C, W, zc, k = cfg.xs[0], cfg.reswidth, cfg.zchannels, cfg.kernel_size
    rd = cfg.level_resdepth
    spec = []

    def conv(prefix, o, i, ks, lin=False):
        spec.append((prefix + ".v", (o, i, ks, ks), "v"))
        spec.append((prefix + ".gain", (o,), "gain_lin" if lin else "gain_log"))
        spec.append((prefix + ".b", (o,), "b"))

    def resblock(prefix, ks, n):
        for l in range(1, n + 1):
            conv(f"{prefix}.res{W}layer{l}.conv1", W, W, ks)
            conv(f"{prefix}.res{W}layer{l}.conv2", W, W, ks, lin=True)

    if not cfg.cond_xscale:
        spec.append(("gen_std", tuple(cfg.xs), "gen_std"))
    conv("infer_in.1", W, 4 * C, 5)
    resblock("infer_res0.0", 5, cfg.nprocessing)
    resblock("infer_res1.0", k, rd[0])
    conv("infer_mu", zc, W, k)
    conv("infer_std", zc, W, k)
    for name in ("deepinfer", "deepgen"):
        for j in range(cfg.nz - 1):
            conv(f"{name}_in.{j}.0", W, zc, k)
        for j in range(cfg.nz - 1):
            resblock(f"{name}_res.{j}.0", k, rd[j + 1])
        for j in range(cfg.nz - 1):
            conv(f"{name}_mu.{j}.0", zc, W, k)
        for j in range(cfg.nz - 1):
            conv(f"{name}_std.{j}.0", zc, W, k)
    conv("gen_in.0", W, zc, k)
    resblock("gen_res1.0", k, rd[0])
    resblock("gen_res0.0", 5, cfg.nprocessing)
    conv("gen_mu.0", 4 * C, W, k)
    if cfg.cond_xscale:
        conv("gen_std.0", 4 * C, W, k)
    return spec


def synthetic_state_dict(cfg: CodecConfig, seed: int = 50, varied: bool = True):
    """Seeded random weights in the reference checkpoint layout.

    varied=False reproduces the *distribution* of the reference's default init
    (utils/torch/modules.py:68-73: v~N(0,0.05), gain 0 (loggain) or 1, b=0,
    gen_std=0).  varied=True additionally draws non-trivial gains, biases and
    x-scales so every term of every epilogue is exercised by the parity tests."""
    rs = np.random.RandomState(seed)
    sd = {}
    for key, shape, kind in state_dict_spec(cfg):
        if kind == "v":
            a = rs.normal(0.0, 0.05, size=shape)
        elif kind == "gain_log":
            a = rs.uniform(-0.3, 0.3, size=shape) if varied else np.zeros(shape)
        elif kind == "gain_lin":
            a = rs.uniform(0.6, 1.0, size=shape) if varied else np.ones(shape)
        elif kind == "b":
            a = rs.uniform(-0.1, 0.1, size=shape) if varied else np.zeros(shape)
        else:  # gen_std
            a = rs.uniform(-2.5, 0.5, size=shape) if varied else np.zeros(shape)
        sd[key] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return sd


def synthetic_bins(cfg: CodecConfig, seed: int = 0):
    """(zendpoints [nz,zdim,2^q-1], zcentres [nz,zdim,2^q]) float64 tensors.

    Top level: equal-mass Logistic(0,1) bins computed in float32 exactly as
    discretization.py:25-27 does through rand.Bins (rand.py:96-128).  Lower
    levels: uniform-width grids between per-dimension sample extrema, which is
    what discretize_kbins(strategy='uniform') yields (discretization.py:105-118);
    the extrema themselves are synthetic (lo=-6-U, hi=6+U), SURVEY.md 8d."""
    from .rand import Bins
    nz, zdim, S = cfg.nz, cfg.zdim, cfg.zsupport
    zendpoints = np.zeros((nz, zdim, S - 1))
    zcentres = np.zeros((nz, zdim, S))
    zbins = Bins(torch.zeros((1, 1, zdim)), torch.ones((1, 1, zdim)), cfg.quantbits)
    zendpoints[nz - 1] = zbins.endpoints().numpy()
    zcentres[nz - 1] = zbins.centres().numpy()
    rs = np.random.RandomState(seed)
    for zi in range(nz - 1):
        lo = -6.0 - rs.uniform(0, 1, size=zdim)
        hi = 6.0 + rs.uniform(0, 1, size=zdim)
        edges = np.linspace(lo, hi, S + 1, axis=1)          # [zdim, S+1]
        zendpoints[zi] = edges[:, 1:-1]
        zcentres[zi] = (edges[:, :-1] + edges[:, 1:]) / 2
    return torch.from_numpy(zendpoints), torch.from_numpy(zcentres)


def synthetic_images(cfg: CodecConfig, n: int, seed: int = 7, kind: str = "uniform"):
    """uint8 [n, C, 32, 32].  'uniform': iid bytes (worst case, maximal
    renormalisation traffic); 'smooth': low-entropy field."""
    C, H, W = cfg.xs
    rs = np.random.RandomState(seed)
    if kind == "uniform":
        return rs.randint(0, 256, size=(n, C, H, W)).astype(np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    ph = rs.uniform(0, 2 * np.pi, size=(n, C, 1, 1))
    fr = rs.uniform(0.1, 0.6, size=(n, C, 1, 1))
    img = 128 + 40 * np.sin(fr * xx + ph) * np.cos(fr * yy - ph) + rs.normal(0, 4, size=(n, C, H, W))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def initial_words(nwords: int, seed: int = 100):
    """The reference's 'random initial bits' (cifar_compress.py:93,157-158):
    nwords draws of randint(2^16, 2^32-1, uint32); the last becomes the head
    (<< 32).  Returns (words uint32 [nwords-1], head python int)."""
    w = np.random.RandomState(seed).randint(low=1 << 16, high=(1 << 32) - 1, size=nwords, dtype=np.uint32)
    return w[:-1].copy(), int(w[-1]) << 32
