"""Run an UNMODIFIED reference script (main.py, or anything that imports encoders / decoders) with the hot path
swapped for turboae_b200:

    cd <a working dir with ./logs ./tmp ./models>
    python -m turboae_b200.launch --reference /path/to/turboae main.py -encoder TurboAE_rate3_cnn ... -num_epoch 0

What it does: puts the reference checkout on sys.path, imports the reference's own
`encoders` / `decoders` / `interleavers` / `cnn_utils` modules and replaces the four attributes
`encoders.ENC_interCNN`, `decoders.DEC_LargeCNN`, `interleavers.Interleaver|DeInterleaver` with this package's
classes, then runs the script with `runpy` as `__main__`.  main.py picks its classes with
`from encoders import ENC_interCNN as ENC` at call time (main.py:35-36, 75-76), so no reference file is edited.

The reference targets PyTorch 1.0 / numpy < 1.20; `_modernise()` restores the handful of names newer libraries
removed (numpy.float/int/complex, fractions.gcd, an importable `matplotlib`, torch.load without map_location on a
CPU-only host, commpy's Python-2 `array(map(...))` interleaver, Lookahead.zero_grad).  That is environment glue, not a change of behaviour.
"""
from __future__ import annotations

import argparse
import fractions
import math
import os
import runpy
import sys
import types


def _modernise():
    import numpy as np
    import torch
    for a, t in (("complex", complex), ("float", float), ("int", int)):
        if not hasattr(np, a):
            setattr(np, a, t)                                   # commpy/channels.py:19
    if not hasattr(fractions, "gcd"):
        fractions.gcd = math.gcd                                # commpy/channelcoding/gfields.py:8
    try:
        import matplotlib  # noqa: F401
    except Exception:
        class _Stub(types.ModuleType):
            def __getattr__(self, k):
                return None
        for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.patches", "matplotlib.mlab"):
            sys.modules.setdefault(n, _Stub(n))                 # commpy/channelcoding/convcode.py:9-11
    try:
        # Python 2's `array(map(...))` (commpy/channelcoding/interleavers.py:19) yields a 0-d object array on Python 3: the classical
        # turbo encoders (-encoder Turbo_rate3_757 / Turbo_rate3_lte, README.md:94-98) would fail in turbo_encode.  Same permutation.
        from commpy.channelcoding import interleavers as _ci

        def interlv(self, in_array):
            return np.asarray(in_array)[np.asarray(self.p_array)]
        _ci._Interleaver.interlv = interlv
    except Exception:  # pragma: no cover -- a checkout without the vendored commpy
        pass
    try:
        # the reference's Lookahead (optimizers.py:10-19) never calls Optimizer.__init__, so it has no `defaults`, which
        # torch >= 2 reads in Optimizer.zero_grad (trainer.py:41): clear the gradients through the wrapped optimizer (same parameters)
        import optimizers as _ro

        def zero_grad(self, set_to_none=True):
            return self.optimizer.zero_grad(set_to_none=set_to_none)
        _ro.Lookahead.zero_grad = zero_grad
    except Exception:  # pragma: no cover -- a checkout without optimizers.py
        pass
    if not torch.cuda.is_available():
        _load = torch.load

        def load(f, *a, **k):
            k.setdefault("map_location", "cpu")                 # main.py:166 has none; checkpoints hold CUDA storages
            return _load(f, *a, **k)
        torch.load = load


def fused_adam_default() -> None:
    """``torch.optim.Adam(params, lr=...)`` as the reference writes it (main.py:196-213) picks torch's multi-tensor ("foreach")
    implementation: a dozen launches and a Python loop over the parameter list per step.  With all parameters on a CUDA device the
    single-kernel implementation (``fused=True``: same update rule, torch's own kernel) is chosen instead unless the caller said
    otherwise; the training step through the unmodified trainer.py is host-bound, so this is wall-clock time."""
    import torch
    if getattr(torch.optim.Adam.__init__, "_tae_fused_default", False):
        return
    _init = torch.optim.Adam.__init__

    def __init__(self, params, *args, **kwargs):
        params = list(params)                       # (the reference passes filter(...) generators)
        if kwargs.get("fused") is None and kwargs.get("foreach") is None and not kwargs.get("differentiable", False):
            leaves = [q for g in params for q in (g["params"] if isinstance(g, dict) else [g])]
            if leaves and all(torch.is_tensor(q) and q.is_cuda and torch.is_floating_point(q) for q in leaves):
                kwargs["fused"] = True
        _init(self, params, *args, **kwargs)
    __init__._tae_fused_default = True
    torch.optim.Adam.__init__ = __init__


def install(reference_root: str) -> None:
    """Swap the hot-path classes inside the (already importable) reference modules."""
    reference_root = os.path.abspath(reference_root)
    if not os.path.isfile(os.path.join(reference_root, "decoders.py")):
        raise FileNotFoundError("no reference checkout at %s (decoders.py not found)" % reference_root)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    _modernise()
    import turboae_b200 as T
    import interleavers as ref_interleavers
    import cnn_utils as ref_cnn_utils
    import encoders as ref_encoders
    import decoders as ref_decoders
    ref_interleavers.Interleaver = T.Interleaver
    ref_interleavers.DeInterleaver = T.DeInterleaver
    ref_encoders.ENC_interCNN = T.ENC_interCNN
    ref_decoders.DEC_LargeCNN = T.DEC_LargeCNN
    ref_decoders.DEC_LargeRNN = T.DEC_LargeRNN
    # the reference's Lookahead (optimizers.py:10-44) is an Optimizer that never calls Optimizer.__init__, so torch's step
    # hooks do not see it, and it writes the slow weights back with fast.data.copy_(slow): tell the weight caches
    try:
        import optimizers as ref_optimizers
        _step, _sync = ref_optimizers.Lookahead.step, ref_optimizers.Lookahead.update_lookahead

        def step(self, closure=None):
            out = _step(self, closure)
            T.invalidate_all()
            return out

        def update_lookahead(self):
            _sync(self)
            T.invalidate_all()
        ref_optimizers.Lookahead.step, ref_optimizers.Lookahead.update_lookahead = step, update_lookahead
    except Exception:  # pragma: no cover -- a checkout without optimizers.py
        pass
    # SameShapeConv1d is imported by name into encoders/decoders at their import time; the replaced classes above
    # build this package's own conv stacks, so cnn_utils is left as is for the out-of-scope variants.
    del ref_cnn_utils


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m turboae_b200.launch", add_help=True)
    ap.add_argument("--reference", default=os.environ.get("TURBOAE_REF", "."), help="path of the turboae checkout")
    ap.add_argument("--seed", type=int, default=None, help="seed numpy / torch before the script starts (the reference sets "
                    "no seed; with one, two runs draw identical bits and noise)")
    ap.add_argument("--stock", action="store_true", help="do NOT swap the hot-path classes: run the reference's own modules "
                    "(only the library-compatibility shim is applied); the comparison arm of scripts/run_reference_dropin.py")
    ap.add_argument("--no-tf32", action="store_true", help="torch's own CUDA convolutions / matmuls in true fp32 (cudnn.allow_tf32 "
                    "defaults to True): makes the --stock arm the reference's fp32 arithmetic (it was written for torch 1.0)")
    ap.add_argument("--no-fused-adam", action="store_true", help="leave torch.optim.Adam's implementation choice alone (default: "
                    "fused=True when every parameter is on a CUDA device)")
    ap.add_argument("--device-channel", action="store_true", help="draw the AWGN noise of trainer.train / validate / test on the GPU "
                    "(turboae_b200.channel.DeviceNoise in place of channels.generate_noise: same distributions, another random "
                    "stream) instead of torch.randn on the CPU + an upload per step")
    ap.add_argument("script", help="reference script to run, e.g. main.py")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    if a.no_tf32:
        import torch
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    if a.stock:
        sys.path.insert(0, os.path.abspath(a.reference))
        _modernise()
    else:
        install(a.reference)
    if not (a.stock or a.no_fused_adam):
        fused_adam_default()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    seed = a.seed
    if world > 1 and not a.stock:
        # data-parallel under torchrun: one process per GPU, whole codewords per rank; the reference's
        # `loss.backward(); optimizer.step()` (trainer.py:74-76) is kept, gradients are averaged by an optimizer pre-step hook
        import torch
        import torch.distributed as dist
        from . import shard
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        os.environ["TURBOAE_B200_SHARD"] = "1"          # ENC_interCNN: batch-global power statistics across ranks
        shard.install_optimizer_hook()
        # main.py sets no seed: every rank would build different random weights (and, for -is_interleave > 1, draw a different
        # permutation) and averaged gradients never bring such replicas together.  All ranks therefore start from ONE seed
        # (rank 0 draws it unless --seed is given), so model construction and the interleaver agree; at the first call of
        # trainer.train / validate / test rank 0's parameters are broadcast once more (belt and braces) and the generators are
        # re-seeded per rank, so that every rank then draws its OWN bits and noise (trainer.py:53-60).
        seed_t = torch.zeros(1, dtype=torch.int64, device="cuda")
        if dist.get_rank() == 0:
            seed_t[0] = seed if seed is not None else int.from_bytes(os.urandom(4), "little")
        dist.broadcast(seed_t, 0)
        seed = int(seed_t.item())
        import trainer as ref_trainer                       # main.py binds train / validate / test from here (main.py:17)
        state = {"synced": False}

        def _wrap(fn, model_pos):
            def wrapped(*args, **kwargs):
                if not state["synced"]:
                    shard.sync_replicas(args[model_pos])
                    shard.seed_everything(seed * 1000003 + 7919 * (dist.get_rank() + 1))
                    state["synced"] = True
                return fn(*args, **kwargs)
            wrapped.__wrapped__ = fn
            return wrapped
        ref_trainer.train = _wrap(ref_trainer.train, 1)          # train(epoch, model, optimizer, args, ...)
        ref_trainer.validate = _wrap(ref_trainer.validate, 0)    # validate(model, optimizer, args, ...)
        ref_trainer.test = _wrap(ref_trainer.test, 0)            # test(model, args, ...)
    if seed is not None:
        from . import shard
        shard.seed_everything(seed)
    if a.device_channel and not a.stock:
        import torch
        import trainer as ref_trainer                       # `from channels import generate_noise` (trainer.py:10) binds the name here
        from .channel import DeviceNoise
        if torch.cuda.is_available() and "--no-cuda" not in a.script_args:
            key = (seed if seed is not None else int.from_bytes(os.urandom(4), "little")) * 1000003 + int(os.environ.get("RANK", "0"))
            ref_trainer.generate_noise = DeviceNoise(ref_trainer.generate_noise, torch.device("cuda", torch.cuda.current_device()), key)
    script = a.script if os.path.isabs(a.script) else os.path.join(os.path.abspath(a.reference), a.script)
    for d in ("logs", "tmp"):
        os.makedirs(d, exist_ok=True)                            # main.py:106, 248 write there
    sys.argv = [script] + a.script_args
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
