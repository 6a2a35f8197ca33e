"""Batched, device-resident Bit-Swap / BB-ANS codec (host-side handle of `bsw_codec`).

Replaces the sender/receiver loop bodies of the reference's compress()
(cifar_compress.py:175-250 and :283-352): one call codes one image per stream for
a whole batch of independent chains; the ANS state never leaves the GPU between
latent levels.  Chain more images onto the same streams by calling again.
"""
import ctypes

import numpy as np
import torch

from ._lib import lib, check, cuda_stream_ptr
from .config import CodecConfig
from .model import Model
from .streams import StreamSet

BITSWAP, BBANS = 0, 1


class Bins:
    """Device copy of the discretisation tables (zendpoints [nz,zdim,2^q-1], zcentres [nz,zdim,2^q],
    the layout discretize() returns, discretization.py:99)."""

    def __init__(self, cfg: CodecConfig, zendpoints, zcentres):
        ze = np.ascontiguousarray(zendpoints.detach().cpu().numpy() if torch.is_tensor(zendpoints) else zendpoints, dtype=np.float64)
        zc = np.ascontiguousarray(zcentres.detach().cpu().numpy() if torch.is_tensor(zcentres) else zcentres, dtype=np.float64)
        assert ze.shape == (cfg.nz, cfg.zdim, cfg.zsupport - 1) and zc.shape == (cfg.nz, cfg.zdim, cfg.zsupport)
        self._h = ctypes.c_void_p()
        check(lib().bsw_bins_create(ctypes.byref(self._h), cfg.nz, cfg.zdim, cfg.quantbits, cfg.xdim, ze.ctypes.data, zc.ctypes.data))

    @property
    def handle(self):
        return self._h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().bsw_bins_destroy(h)
            except Exception:
                pass
            self._h = None


class BitSwapCodec:
    def __init__(self, cfg: CodecConfig, model: Model, bins: Bins, max_batch: int):
        self.cfg, self.model, self.bins, self.max_batch = cfg, model, bins, int(max_batch)
        self._h = ctypes.c_void_p()
        check(lib().bsw_codec_create(ctypes.byref(self._h), model.handle, bins.handle, self.max_batch))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().bsw_codec_destroy(h)
            except Exception:
                pass
            self._h = None

    def encode(self, streams: StreamSet, x: torch.Tensor, first=0, scheme=BITSWAP):
        """x: uint8 CUDA tensor [count, C, 32, 32]; pushes one image onto each of `count` streams. Async."""
        assert x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous()
        count = x.shape[0]
        check(lib().bsw_codec_encode(self._h, streams.handle, first, count, x.data_ptr(), scheme, cuda_stream_ptr()))

    def decode(self, streams: StreamSet, count: int, first=0, scheme=BITSWAP, out=None):
        """Pops one image from each stream; returns uint8 CUDA tensor [count, C, 32, 32]. Async."""
        if out is None:
            out = torch.empty((count,) + tuple(self.cfg.xs), dtype=torch.uint8, device="cuda")
        check(lib().bsw_codec_decode(self._h, streams.handle, first, count, out.data_ptr(), scheme, cuda_stream_ptr()))
        return out

    @property
    def last_launches(self):
        return int(lib().bsw_codec_last_launches(self._h))

    CATEGORIES = ("misc", "conv_in", "conv_dense3x3", "conv_dense5x5", "conv_head", "pop_z", "push_z", "pop_x",
                  "push_x", "prior", "rows_z", "rows_x")

    def set_two_phase(self, on=True):
        """True (default): parallel row-table kernel + serial coder; False: fused one-warp-per-stream kernels."""
        check(lib().bsw_codec_set_two_phase(self._h, int(on)))

    def profile(self, enable=-1):
        """Per-kernel-category device time since profiling was enabled: {name: (ms, launches)}.
        enable: True/False resets and starts/stops, -1 just reads."""
        ms = np.zeros(12, dtype=np.float64)
        n = np.zeros(12, dtype=np.int64)
        check(lib().bsw_codec_profile(self._h, int(enable), ms.ctypes.data, n.ctypes.data))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.CATEGORIES)}
