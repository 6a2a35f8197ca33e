"""Distribution helpers with the reference's call signatures.

Drop-in for the subset of utils/torch/rand.py the compression scripts use
(reference: utils/torch/rand.py:67-72 logistic_cdf/icdf, :78-128 Bins,
:134-153 ImageBins).  These build the one-off bin tables that are *inputs* of
the hot path; the per-symbol logistic-CDF work of the hot path itself is fused
into the CUDA coder (csrc/ans_kernels.cu) and never goes through here.
"""
import torch


def logistic_cdf(x, mu, scale):
    """CDF of Logistic(mu, scale) at x (rand.py:67-68)."""
    return torch.sigmoid((x - mu) / scale)


def logistic_icdf(p, mu, scale):
    """Quantile function of Logistic(mu, scale) (rand.py:71-72)."""
    return mu + scale * torch.log(p / (1. - p))


class Bins:
    """Equal-mass discretisation bins of Logistic(mu, scale) with 2^precision
    bins (rand.py:78-128).  endpoints(): shape mu.shape+[2^p-1] (the +-inf outer
    endpoints are implicit); centres(): mu.shape+[2^p].  dtype/device follow mu."""

    def __init__(self, mu, scale, precision):
        self.mu, self.scale, self.precision = mu, scale, precision
        self.nbins = 1 << precision
        self.type, self.device, self.shape = mu.dtype, mu.device, list(mu.shape)

    def _quantiles(self, probs):
        nd = len(self.shape)
        probs = probs.view([-1] + [1] * nd).expand([-1] + self.shape)        # [n, *shape]
        q = logistic_icdf(probs, self.mu, self.scale)
        return q.permute(list(range(1, nd + 1)) + [0])                       # [*shape, n]

    def endpoints(self):
        probs = torch.arange(1., self.nbins, dtype=self.type, device=self.device) / self.nbins
        return self._quantiles(probs)

    def centres(self):
        probs = (torch.arange(end=self.nbins, dtype=self.type, device=self.device) + .5) / self.nbins
        return self._quantiles(probs)


class ImageBins:
    """Pixel bins of the discretised logistic in [-1,1] (rand.py:134-153):
    255 inner endpoints ((k-127.5)/127.5 - 1/255, k=1..255) and 256 centres
    ((k-127.5)/127.5), identical for every one of `shape` dimensions."""

    def __init__(self, type, device, shape):
        self.type, self.device, self.shape = type, device, [shape]

    def endpoints(self):
        k = torch.arange(1, 256, dtype=self.type, device=self.device)
        e = ((k - 127.5) / 127.5) - 1. / 255.
        return e[None,].expand(self.shape + [-1])

    def centres(self):
        k = torch.arange(0, 256, dtype=self.type, device=self.device)
        c = (k - 127.5) / 127.5
        return c[None,].expand(self.shape + [-1])
