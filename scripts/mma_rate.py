#!/usr/bin/env python
"""MMA issue-rate microbenchmark with optional concurrent shared-memory traffic (development tool; GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from turboae_b200 import _lib
lib = _lib.load()
n = 16384
src = torch.zeros(64 * 7168, dtype=torch.uint8, device="cuda")
for cg, N, grid, tma, sts in [(2, 112, 1, 0, 0), (2, 112, 74, 0, 0), (2, 112, 74, 1, 0), (2, 112, 74, 0, 1), (2, 112, 74, 0, 2), (2, 112, 74, 0, 4),
                              (2, 112, 74, 1, 2), (1, 112, 148, 0, 0), (1, 112, 148, 1, 0), (1, 112, 148, 0, 2), (2, 64, 74, 1, 0), (2, 64, 74, 0, 2)]:
    cyc = torch.zeros(256, dtype=torch.int64, device="cuda")
    _lib.check(lib.tae_debug_probe_rate(cg, N, n, 0, grid, _lib.ptr(cyc), _lib.ptr(src), tma, sts, _lib.stream_ptr()))
    torch.cuda.synchronize()
    c = cyc[:grid].float()
    tot = c.mean().item()
    print("cta_group %d N %3d grid %3d tma %d sts_warps %d : %.1f cycles/MMA | MMA operand B/clk %.1f | tma B/clk %.1f | sts B/clk %.1f" % (
        cg, N, grid, tma, sts, tot / n, (4096 + (N // cg) * 32) * n / tot, cyc[128].item() / tot, cyc[192].item() / tot))
