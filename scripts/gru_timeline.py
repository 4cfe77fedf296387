#!/usr/bin/env python
"""clock64 timeline of the tensor-core GRU recurrence (cluster 0, first 64 steps) for one layer-direction launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from turboae_b200 import _lib
lib = _lib.load()
dev = "cuda"
B, L, H = int(os.environ.get("RNN_B", "18944")), 200, 100
for cin, grp in ((7, 7), (200, 100)):
    R = lib.tae_gru_rows_per_block(B)
    n_in = (cin // grp) * (8 * ((grp + 7) // 8)) // 8
    x = (torch.randn(lib.tae_gru_tile_bytes(B, L, n_in, R) // 2, device=dev) * 0.5).to(torch.bfloat16)
    w = [torch.randn(3 * H, cin, device=dev) * 0.1, torch.randn(3 * H, H, device=dev) * 0.1, torch.zeros(3 * H, device=dev), torch.zeros(3 * H, device=dev)]
    packed = torch.empty(lib.tae_gru_packed_bytes(H, cin, grp), dtype=torch.uint8, device=dev)
    _lib.check(lib.tae_gru_pack_bf16(*[_lib.ptr(t) for t in w], _lib.ptr(packed), H, cin, grp, _lib.stream_ptr()))
    out = torch.empty(lib.tae_gru_tile_bytes(B, L, 26, R), dtype=torch.uint8, device=dev)
    ws = torch.zeros(256, dtype=torch.uint8, device=dev)
    tl = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
    run = lambda: _lib.check(lib.tae_gru_direction_bf16(_lib.ptr(packed), _lib.ptr(x), _lib.ptr(out), B, L, H, cin, grp, R, 26, 0, 0, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    run(); torch.cuda.synchronize()
    lib.tae_debug_gru_timeline(_lib.ptr(tl))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    lib.tae_debug_gru_timeline(None)
    t = tl.cpu().numpy().reshape(64, 8)
    print("R=%d " % R, end=""); print("in=%d: launch %.3f ms for %d steps = %.2f us/step" % (cin, e0.elapsed_time(e1), L, 1e3 * e0.elapsed_time(e1) / L))
    base = t[20, 0]
    names = ["mma:wait_start", "mma:rdy", "mma:issued", "load:start", "load:done", "epi:acc0", "epi:done", "epi:acc1"]
    for s in range(20, 24):
        print("  step %d: " % s + "  ".join("%s=%d" % (n, t[s, i] - base) for i, n in enumerate(names)))
    d = np.diff(t[10:60, 1])
    print("  cycles/step (rdy to rdy): mean %.0f min %d max %d" % (d.mean(), d.min(), d.max()))
