"""Build libturboae_b200.so in-tree with nvcc for sm_100a (no GPU needed: cross-compiles)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.environ.get("TURBOAE_B200_LIB") or os.path.join(LIBDIR, "libturboae_b200.so")
SOURCES = ["tae_api.cu", "tae_f32.cu", "tae_dec_pair.cu", "tae_channel.cu", "tae_gru.cu", "tae_wgrad.cu", "tae_gru_tc.cu", "tae_x3.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "turboae_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    extra = os.environ.get("TURBOAE_B200_NVCC_FLAGS", "").split()      # e.g. -DTAE_TIMELINE=1 for scripts/dec_timeline.py
    cmd = [find_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
        ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed (exit %d): %s" % (r.returncode, " ".join(cmd)))
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
