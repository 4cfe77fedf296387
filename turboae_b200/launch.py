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
CPU-only host).  That is environment glue, not a change of behaviour.
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
    if not torch.cuda.is_available():
        _load = torch.load

        def load(f, *a, **k):
            k.setdefault("map_location", "cpu")                 # main.py:166 has none; checkpoints hold CUDA storages
            return _load(f, *a, **k)
        torch.load = load


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
    ap.add_argument("script", help="reference script to run, e.g. main.py")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    if a.stock:
        sys.path.insert(0, os.path.abspath(a.reference))
        _modernise()
    else:
        install(a.reference)
    if a.seed is not None:
        import numpy as np
        import torch
        np.random.seed(a.seed)
        torch.manual_seed(a.seed)
    world = int(os.environ.get("WORLD_SIZE", "1"))
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
    script = a.script if os.path.isabs(a.script) else os.path.join(os.path.abspath(a.reference), a.script)
    for d in ("logs", "tmp"):
        os.makedirs(d, exist_ok=True)                            # main.py:106, 248 write there
    sys.argv = [script] + a.script_args
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
