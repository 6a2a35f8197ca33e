"""ELBO pieces against the reference's own log-density functions (golden) and the assembly on the torch oracle."""
import os

import numpy as np
import torch

from bitswap_b200 import synthetic
from bitswap_b200.config import preset
from bitswap_b200.elbo import discretized_logistic_logp, elbo, logistic_logp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_logp_functions_match_reference():
    g = np.load(os.path.join(GOLDEN, "logp.npz"))
    t = torch.from_numpy
    assert np.abs(logistic_logp(t(g["mu"]), t(g["sc"]), t(g["x"])).numpy() - g["lp"]).max() < 1e-12
    assert np.abs(discretized_logistic_logp(t(g["xm"]), t(g["xs"]), t(g["xx"])).numpy() - g["dl"]).max() < 1e-12


def test_elbo_assembly_on_oracle_nets():
    """With random-init nets the ELBO must upper-bound nothing in particular, but its pieces must be finite, the
    reconstruction term must equal the sum of per-pixel log-probs, and the value is deterministic given eps."""
    from oracle import oracle as O
    cfg = preset("tiny")
    m = O.ModelOracle(cfg, synthetic.synthetic_state_dict(cfg, seed=50, varied=True))
    x = torch.from_numpy(synthetic.synthetic_images(cfg, 3, seed=2))
    gen = torch.Generator().manual_seed(1)
    eps = [torch.randn(3, cfg.zdim, generator=gen, dtype=torch.float64) for _ in range(cfg.nz)]
    a = elbo(m, x, cfg=cfg, eps=eps)
    b = elbo(m, x, cfg=cfg, eps=eps)
    assert torch.equal(a["elbo_bits_per_dim"], b["elbo_bits_per_dim"]) and torch.isfinite(a["elbo_bits_per_dim"]).all()
    assert a["logenc"].shape == (cfg.nz, 3) and a["logdec"].shape == (cfg.nz, 3) and (a["logrecon"] < 0).all()
    total = -a["logrecon"] + (-a["logdec"] + a["logenc"]).sum(0)
    assert torch.allclose(total / cfg.xdim, a["elbo_bits_per_dim"])
