"""Batched ELBO of the hierarchical VAE on the GPU nets (the number the reference logs next to the code length).

Reference: Model.loss (model/cifar_train.py:441-490) as used by the compression scripts for the
"net bits - ELBO" gap (cifar_compress.py:169-173,258,262), with the log-densities of utils/torch/rand.py:23-64.
Works with any object exposing infer(i)(given) / generate(i)(given) on flat [B, dim] inputs in compressing mode
(bitswap_b200.model.Model on the GPU; a torch-CPU restatement of the nets in the tests).
"""
import math

import torch
import torch.nn.functional as F


def _softplus(x):
    return -F.logsigmoid(-x)        # utils/torch/modules.py:112-114


def logistic_logp(mu, scale, x):
    """log-density of Logistic(mu, scale) at x (rand.py:23-27), elementwise."""
    y = -(x - mu) / scale
    return -y - torch.log(scale) - 2 * _softplus(-y)


def discretized_logistic_logp(mu, scale, x):
    """log-probability of pixel values x in [0,255] under the discretised logistic with bins of width 2/255 on [-1,1]
    (rand.py:31-64, after the PixelCNN++ loss): edge bins use the cdf tails, vanishing bins the density."""
    xr = (x - 127.5) / 127.5
    inv = 1. / scale
    xc = xr - mu
    plus_in, min_in, mid_in = inv * (xc + 1. / 255.), inv * (xc - 1. / 255.), inv * xc
    cdf_delta = torch.sigmoid(plus_in) - torch.sigmoid(min_in)
    log_cdf_plus = plus_in - _softplus(plus_in)
    log_one_minus_cdf_min = -_softplus(min_in)
    log_pdf_mid = mid_in - torch.log(scale) - 2. * _softplus(mid_in)
    inner = torch.where(cdf_delta > 1e-5, torch.log(torch.clamp(cdf_delta, min=1e-12)), log_pdf_mid - math.log(127.5))
    out = torch.where(xr > .999, log_one_minus_cdf_min, inner)
    return torch.where(xr < -.999, log_cdf_plus, out)


def logistic_eps(shape, device, bound=1e-5, generator=None):
    u = torch.rand(shape, device=device, generator=generator).clamp_(min=bound, max=1 - bound)
    return torch.log(u) - torch.log1p(-u)


@torch.no_grad()
def elbo(model, x, cfg=None, eps=None, generator=None):
    """x: uint8 [B, C, 32, 32].  Returns dict with per-image tensors in bits: logrecon [B], logenc [nz,B], logdec [nz,B],
    elbo_bits_per_dim [B] = (-logrecon + sum(-logdec + logenc)) / xdim (cifar_compress.py:172,258).
    eps: optional list of nz noise tensors [B, zdim] (Logistic(0,1)); drawn with `generator` otherwise."""
    cfg = cfg or model.cfg
    B = x.shape[0]
    dev = x.device
    xf = x.reshape(B, -1).double()
    given = (xf - 127.5) / 127.5
    log2e = math.log2(math.e)
    logenc = torch.zeros((cfg.nz, B), dtype=torch.float64, device=dev)
    logdec = torch.zeros((cfg.nz, B), dtype=torch.float64, device=dev)
    logrecon = None
    z = None
    for i in range(cfg.nz):
        mu, scale = model.infer(i)(given if i == 0 else z)
        mu, scale = mu.to(dev), scale.to(dev)
        e = eps[i].to(dev) if eps is not None else logistic_eps(mu.shape, dev, generator=generator).double()
        z_next = mu + scale * e                                            # cifar_train.py:455-458
        logenc[i] = logistic_logp(mu, scale, z_next).sum(dim=1)           # :462-463
        mu, scale = model.generate(i)(z_next)
        mu, scale = mu.to(dev), scale.to(dev)
        if i == 0:
            logrecon = discretized_logistic_logp(mu, scale, xf).sum(dim=1)  # :472-474
        else:
            logdec[i - 1] = logistic_logp(mu, scale, z).sum(dim=1)         # :477-478
        z = z_next
    zero, one = torch.zeros(1, dtype=torch.float64, device=dev), torch.ones(1, dtype=torch.float64, device=dev)
    logdec[cfg.nz - 1] = logistic_logp(zero, one, z).sum(dim=1)            # :483-484
    logrecon, logenc, logdec = logrecon * log2e, logenc * log2e, logdec * log2e
    total = -logrecon + (-logdec + logenc).sum(dim=0)
    return dict(logrecon=logrecon, logenc=logenc, logdec=logdec, elbo_bits_per_dim=total / cfg.xdim)
