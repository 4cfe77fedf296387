"""Throughput of the split-operand tensor path (precision 'f16x3'): encoder and decoder, next to the fp32 CUDA-core path and the bf16
fused kernel.  python scripts/x3_bench.py [B]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_codec  # noqa: E402


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    out = {"B": B}
    for cfg in ("c1", "c3"):
        m, w, p = build_codec(cfg)
        u = torch.randint(0, 2, (B, 100, 1), device="cuda").float()
        with torch.no_grad():
            for prec in ("f16x3", "bf16", "fp32"):
                m.enc.precision = prec
                ms = timed(lambda: m.enc(u), n=3 if prec == "fp32" else 10)
                out["%s_enc_%s_cw_per_s" % (cfg, prec)] = B / ms * 1e3
            m.enc.precision = "f16x3"
            r = m.enc(u) + torch.randn(B, 100, 3, device="cuda")
            for prec in ("f16x3", "bf16"):
                ms = timed(lambda: m.dec.decode(r, precision=prec), n=3)
                out["%s_dec_%s_cw_per_s" % (cfg, prec)] = B / ms * 1e3
                out["%s_dec_%s_ms" % (cfg, prec)] = ms
            y3 = m.dec.decode(r, precision="f16x3")
            yb = m.dec.decode(r, precision="bf16")
            out["%s_hard_disagreement_x3_vs_bf16" % cfg] = float((torch.round(y3) != torch.round(yb)).float().mean())
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
