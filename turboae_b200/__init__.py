"""turboae_b200 -- B200 (sm_100a) implementation of the TurboAE CNN encode / turbo-decode hot path.

Drop-in for the reference's ``encoders.ENC_interCNN`` / ``decoders.DEC_LargeCNN`` (and the
``Interleaver`` / ``DeInterleaver`` / ``SameShapeConv1d`` building blocks): same constructor,
``forward`` surface, parameter names and checkpoint keys, backed by hand-written CUDA kernels in
``libturboae_b200.so`` (C ABI in ``include/turboae_b200.h``).  There is no CPU fallback: without
the library or without a CUDA device every forward raises.
"""
from .interleavers import Interleaver, DeInterleaver          # noqa: F401
from .cnn_utils import SameShapeConv1d, DenseSameShapeConv1d  # noqa: F401
from .encoders import ENCBase, ENC_interCNN                   # noqa: F401
from .decoders import DEC_LargeCNN, DEC_LargeRNN              # noqa: F401
from . import channel, shard, graphs                          # noqa: F401
from ._flat import invalidate_all                             # noqa: F401


def _register_step_hook():
    # weight caches follow every optimizer step, including in-place writes through ``.data`` that bump no version counter
    try:
        from torch.optim.optimizer import register_optimizer_step_post_hook
        register_optimizer_step_post_hook(lambda opt, args, kwargs: invalidate_all())
    except Exception:  # pragma: no cover -- very old torch: FlatCache still follows the version counters
        pass


_register_step_hook()

__version__ = "0.1.0"
