"""Drop-in `ANS` with the reference's constructor and method signatures.

Reference: class ANS, cifar_compress.py:12-67 (five more identical copies in the
other *_compress.py scripts).  Same arguments, same list-in/list-out state, same
exception types; the arithmetic runs in csrc/ans_kernels.cu through the C ABI
(bsw_ans_tables / bsw_ans_push / bsw_ans_pop).  This class is the compatibility
surface -- one stream, host list in and out per call, like the reference.  The
fast path is the batched, device-resident codec in codec.py.
"""
import ctypes

import numpy as np
import torch

from ._lib import lib, check, cuda_stream_ptr, on_device
from .streams import StreamSet


class ANS:
    def __init__(self, pmfs, bits=31, quantbits=8):
        if not torch.cuda.is_available():
            raise RuntimeError("bitswap_b200.ANS needs a CUDA device (no CPU fallback)")
        self.device = pmfs.device
        self.bits, self.quantbits = bits, quantbits
        self.seq_len, self.support = pmfs.shape
        # the tables live where the pmfs live (the reference: cdfs on pmfs.device, cifar_compress.py:28-43); every
        # launch below runs on THAT device, whatever the caller's current device is
        dev = pmfs.device if pmfs.is_cuda else torch.device("cuda", torch.cuda.current_device())
        self._dev = dev
        pm = pmfs.to(device=dev, dtype=torch.float64).contiguous()
        self._P = torch.empty((self.seq_len, self.support), dtype=torch.int32, device=dev)
        self._C = torch.empty((self.seq_len, self.support + 1), dtype=torch.int32, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        with on_device(dev.index):
            check(lib().bsw_ans_tables(pm.data_ptr(), self.seq_len, self.support, bits, quantbits,
                                       self._P.data_ptr(), self._C.data_ptr(), err.data_ptr(), cuda_stream_ptr()))
        assert int(err.item()) == 0, "cdf table does not sum to 2^bits"      # cifar_compress.py:45-46

    # integer tables as the reference exposes them (numpy int64 on the host)
    @property
    def pmfs(self):
        return self._P.cpu().numpy().astype(np.int64) & 0xffffffff

    @property
    def cdfs(self):
        return self._C.cpu().numpy().astype(np.int64) & 0xffffffff

    def _run(self, x, push, symbols=None):
        ss = StreamSet(1, len(x) + self.seq_len + 64, device=self._dev)
        ss.import_lists([x])
        sym = torch.empty(self.seq_len, dtype=torch.int32, device=self._dev)
        if push:
            s_in = torch.as_tensor(symbols).reshape(-1)
            # the reference indexes self.pmfs[i, s] for i, s in enumerate(symbols) (cifar_compress.py:49-50): a wrong length
            # or a symbol outside the support is an IndexError there, and must not become an out-of-bounds device read here
            if s_in.numel() != self.seq_len:
                raise IndexError(f"ANS.encode: {s_in.numel()} symbols for a table of {self.seq_len} rows")
            if s_in.numel() and (int(s_in.min()) < 0 or int(s_in.max()) >= self.support):
                raise IndexError(f"ANS.encode: symbol outside [0, {self.support})")
            sym.copy_(s_in.to(torch.int32))
        fn = lib().bsw_ans_push if push else lib().bsw_ans_pop
        with on_device(self._dev.index):
            check(fn(ss.handle, 0, 1, self._P.data_ptr(), self._C.data_ptr(), 0, 0, sym.data_ptr(),
                     self.seq_len, self.support, self.bits, cuda_stream_ptr()))
        ss.raise_on_error()
        x[:] = ss.export_lists()[0]          # the reference mutates and returns the same list
        return x, sym

    def encode(self, x, symbols):
        return self._run(x, True, symbols)[0]

    def decode(self, x):
        x, sym = self._run(x, False)
        return x, sym.to(torch.int64).to(self.device)
