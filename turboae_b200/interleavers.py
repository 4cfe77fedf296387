"""Interleaver / DeInterleaver with the reference's nn.Module surface (reference interleavers.py:6-48),
executing ``out[b,i,f] = in[b,p[i],f]`` as one gather kernel (``tae_interleave_f32``).  Bit-exact."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _as_perm(p_array) -> np.ndarray:
    p = np.asarray(p_array).reshape(-1).astype(np.int64)
    return p


def inverse_permutation(p: np.ndarray) -> np.ndarray:
    """rp[p[i]] = i (reference interleavers.py:29-33)."""
    rp = np.zeros_like(p)
    rp[p] = np.arange(len(p), dtype=p.dtype)
    return rp


class _Gather(torch.autograd.Function):
    """out[b,i,f] = in[b,idx[i],f]; backward is the gather with the inverse index."""

    @staticmethod
    def forward(ctx, x, idx_dev, inv_dev):
        ctx.save_for_backward(idx_dev, inv_dev)
        return gather_rows(x, idx_dev)

    @staticmethod
    def backward(ctx, g):
        idx_dev, inv_dev = ctx.saved_tensors
        return gather_rows(g.contiguous(), inv_dev), None, None


def gather_rows(x: torch.Tensor, idx_dev: torch.Tensor) -> torch.Tensor:
    _lib.require_cuda(x, "Interleaver input")
    if x.dim() != 3:
        raise _lib.TaeError("Interleaver expects a (B, L, F) tensor, got shape %s" % (tuple(x.shape),))
    if x.dtype != torch.float32:
        raise _lib.TaeError("Interleaver expects float32, got %s" % x.dtype)
    B, L, F = x.shape
    if L != idx_dev.numel():
        raise _lib.TaeError("Interleaver: block length %d != permutation length %d" % (L, idx_dev.numel()))
    x = x.contiguous()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().tae_interleave_f32(_lib.ptr(x), _lib.ptr(out), _lib.ptr(idx_dev), B, L, F,
                                                  _lib.stream_ptr(x.device)))
    return out


class _PermHolder(torch.nn.Module):
    def __init__(self, args, p_array):
        super().__init__()
        self.args = args
        self._p = None
        self._dev = {}
        self.set_parray(p_array)

    def set_parray(self, p_array):
        p = _as_perm(p_array)
        if self._p is not None and len(p) == len(self._p) and np.array_equal(p, self._p):
            return                                   # Channel_AE.forward re-sets the same p every call (channel_ae.py:32-36)
        if sorted(p.tolist()) != list(range(len(p))):
            raise _lib.TaeError("p_array is not a permutation of 0..%d" % (len(p) - 1))
        self._p = p
        self._rp = inverse_permutation(p)
        self._dev = {}
        self.p_array = torch.LongTensor(p)                      # reference attribute (interleavers.py:10)
        self.reverse_p_array = torch.LongTensor(self._rp)       # reference attribute (interleavers.py:33)

    def device_index(self, device):
        """(perm, inverse perm) as int32 tensors on `device`, uploaded once per permutation."""
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self._p.astype(np.int32)).to(device),
                              torch.from_numpy(self._rp.astype(np.int32)).to(device))
        return self._dev[key]


class Interleaver(_PermHolder):
    """reference interleavers.py:6-21."""

    def forward(self, inputs):
        p, rp = self.device_index(inputs.device)
        return _Gather.apply(inputs, p, rp)


class DeInterleaver(_PermHolder):
    """reference interleavers.py:24-48."""

    def forward(self, inputs):
        p, rp = self.device_index(inputs.device)
        return _Gather.apply(inputs, rp, p)
