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
    """The flat fp32 image of a module's parameters.  Steady state (training: every optimizer step moves the version counters)
    costs ONE multi-tensor copy into the persistent flat buffer through cached views -- no per-parameter torch calls and no
    re-allocation, so the image keeps its address (CUDA-graph replays and the packed images derived from it stay valid).  The
    views are rebuilt only when the parameter set itself changes (other sizes or another device)."""

    def __init__(self):
        self._key = None
        self._sig = None
        self._views = None
        self.flat = None
        self.derived = {}

    def invalidate(self) -> None:
        self._key = None
        self.derived = {}

    def get(self, params):
        key = (_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in params)
        if key != self._key:
            with torch.no_grad():
                dev = params[0].device if params else None
                sig = (dev,) + tuple(p.numel() for p in params)
                if sig != self._sig:
                    self.flat = torch.empty(sum(sig[1:]), dtype=torch.float32, device=dev)
                    self._views = [v.view(p.shape) for v, p in zip(self.flat.split_with_sizes(list(sig[1:])), params)]
                    self._sig = sig
                torch._foreach_copy_(self._views, list(params))        # (no_grad: plain copies out of the leaves)
            self._key = key
            self.derived = {}
        return self.flat


def _drop_ordered_hook(module, incompatible_keys):     # (module-level: a whole-module pickle / deepcopy must survive it)
    module._drop_ordered()


class OrderedParameters:
    """Mixin: ``ordered_parameters()`` -- the module's parameters in the canonical flat order of include/turboae_b200.h --
    walks the module tree once (``_walk_ordered_parameters``) and keeps the list; it is dropped whenever the Parameter objects
    can have been replaced (``_apply`` = .to() / .cuda() / dtype casts, ``load_state_dict`` at any level of the tree,
    ``set_parallel``)."""

    def _drop_ordered(self, *unused):
        self.__dict__["_ordered"] = None

    def _watch_ordered(self):
        self._drop_ordered()
        self.register_load_state_dict_post_hook(_drop_ordered_hook)

    def _apply(self, fn, *a, **k):
        self._drop_ordered()
        return super()._apply(fn, *a, **k)

    def ordered_parameters(self):
        ps = self.__dict__.get("_ordered")
        if ps is None:
            ps = self.__dict__["_ordered"] = self._walk_ordered_parameters()
        return ps


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
