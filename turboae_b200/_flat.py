"""Flat-parameter cache: the C ABI takes all weights of a codec as ONE contiguous fp32 buffer in a canonical
order.  The nn.Parameters stay owned by PyTorch (optimisers step them in place); the flat copy (and anything
derived from it, e.g. the bf16 tensor-core image) is rebuilt only when some parameter's storage or version changes."""
from __future__ import annotations

import torch

# Parameter mutations that autograd's version counter does not see (``p.data.copy_(...)``: the reference's own Lookahead
# optimizer, optimizers.py:29, writes the slow weights back that way) are caught by a process-wide epoch that every
# optimizer step bumps (post-step hook registered in turboae_b200/__init__.py; launch.py also wraps Lookahead.step, which does
# not go through torch's hook machinery).  Anything else that writes through ``.data`` must call ``invalidate_all()``.
_EPOCH = [0]


def invalidate_all() -> None:
    """Force every FlatCache to rebuild its flat copy (and the derived bf16 images) at the next forward."""
    _EPOCH[0] += 1


class FlatCache:
    def __init__(self):
        self._key = None
        self.flat = None
        self.derived = {}

    def invalidate(self) -> None:
        self._key = None
        self.derived = {}

    def get(self, params):
        key = (_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in params)
        if key != self._key:
            with torch.no_grad():
                self.flat = torch.cat([p.detach().reshape(-1).to(torch.float32) for p in params]).contiguous()
            self._key = key
            self.derived = {}
        return self.flat


class Workspace:
    """Grow-only device scratch buffer (borrowed by the library for the duration of a call)."""

    def __init__(self):
        self._buf = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self._buf is None or self._buf.numel() < nbytes or self._buf.device != device:
            self._buf = torch.zeros(max(nbytes, 256), dtype=torch.uint8, device=device)
        return self._buf


class ParallelShim(torch.nn.Module):
    """Stands where the reference puts ``torch.nn.DataParallel(sub_module)`` (reference encoders.py:343-349,
    decoders.py:194-199): it only contributes the ``.module.`` level to the state_dict keys that all shipped
    checkpoints carry.  Multi-GPU is one process per GPU with whole codewords per rank instead."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


def unwrap(m):
    return m.module if isinstance(m, ParallelShim) else m
