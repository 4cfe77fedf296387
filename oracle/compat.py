"""Environment glue that lets the UNMODIFIED reference (/root/reference,
PyTorch-1.0 / numpy<1.20 era) import under Python 3.12 / numpy 2 / torch 2.11.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py in the build
container; ``/root/reference`` does not exist on the GPU box).  It edits no
reference file; it only restores names that newer libraries removed:

* ``numpy.complex/float/int`` aliases     (reference commpy/channels.py:19)
* ``fractions.gcd``                       (reference commpy/channelcoding/gfields.py:8)
* stub ``matplotlib`` modules             (reference commpy/channelcoding/convcode.py:9-11)
* ``torch.load(map_location='cpu')``      (reference main.py:166 has none; the
                                           shipped checkpoints hold CUDA storages)
"""
import fractions
import math
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("TURBOAE_REF", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "decoders.py"))


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        return None


_installed = False


def install() -> None:
    """Idempotently install the shim and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for a, t in (("complex", complex), ("float", float), ("int", int)):
        if not hasattr(np, a):
            setattr(np, a, t)
    if not hasattr(fractions, "gcd"):
        fractions.gcd = math.gcd
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.collections",
              "matplotlib.patches", "matplotlib.mlab"):
        if n not in sys.modules:
            sys.modules[n] = _Stub(n)
    if not torch.cuda.is_available():
        _l = torch.load

        def _load(f, *a, **k):
            k.setdefault("map_location", "cpu")
            return _l(f, *a, **k)

        torch.load = _load
    _installed = True


def reference_args(argv):
    """Parse ``argv`` (list of CLI tokens) with the reference's own get_args()."""
    install()
    import get_args as _ga  # reference get_args.py:4
    old = sys.argv
    try:
        sys.argv = ["main.py"] + list(argv)
        return _ga.get_args()
    finally:
        sys.argv = old
