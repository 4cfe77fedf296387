#!/usr/bin/env python
"""Stage the UNMODIFIED reference checkout into the git-ignored baseline/_ref/ so that it travels to the GPU box
(gpurun ships git-ignored files; /root/reference itself does not exist there).

    python scripts/stage_reference.py [--src /root/reference]

Nothing is edited: Python sources, the vendored commpy package and the four shipped checkpoints are copied byte for
byte (docs / results / training logs are left out).  A manifest with the sha256 of every staged file is written to
baseline/_ref/MANIFEST.json; tests/test_reference_dropin.py and bench.py --impl reference use the staged copy when it
exists and say so.  baseline/_ref/ is never imported by turboae_b200/ (the product has no reference dependency)."""
import argparse
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("TURBOAE_REF", "/root/reference"))
    ap.add_argument("--dst", default=os.path.join(ROOT, "baseline", "_ref"))
    a = ap.parse_args()
    if not os.path.isfile(os.path.join(a.src, "decoders.py")):
        raise SystemExit("no reference checkout at %s" % a.src)
    if os.path.isdir(a.dst):
        shutil.rmtree(a.dst)
    os.makedirs(a.dst)
    manifest = {}
    for base, dirs, files in os.walk(a.src):
        rel = os.path.relpath(base, a.src)
        top = rel.split(os.sep)[0]
        if top in ("docs", "results", "tmp", ".git", "__pycache__"):
            dirs[:] = []
            continue
        dirs[:] = [d for d in dirs if d not in ("__pycache__", ".git")]
        for f in files:
            if not (f.endswith(".py") or f.endswith(".pt") or f in ("README.md", "LICENSE")):
                continue
            s = os.path.join(base, f)
            d = os.path.join(a.dst, rel, f) if rel != "." else os.path.join(a.dst, f)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            manifest[os.path.relpath(d, a.dst)] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    json.dump({"source": a.src, "files": manifest}, open(os.path.join(a.dst, "MANIFEST.json"), "w"), indent=0, sort_keys=True)
    print("staged %d files (%.1f MB) into %s" % (len(manifest), sum(os.path.getsize(os.path.join(a.dst, k)) for k in manifest) / 1e6, a.dst))


if __name__ == "__main__":
    main()
