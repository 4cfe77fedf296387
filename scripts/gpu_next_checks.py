#!/usr/bin/env python
"""First GPU checks of the next round (not collected by pytest on purpose: they have never run on a GPU yet).

1. gru_pair_kernel with MORE than one block pair per cluster (batch > 18 944 at 128 rows per block): the persistent loop
   carries the hidden-state buffer parity and the barrier phases across pairs.
2. Run `torchrun --nproc-per-node 8 bench.py --gpus 8 --steps 5 --warmup 3` separately.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from turboae_b200 import _lib
from oracle import turboae_oracle as O

lib = _lib.load()
DEV = "cuda"
B, L, H, cin = 40000, 6, 100, 7
rs = np.random.RandomState(0)
q = lambda a: torch.from_numpy(a).to(torch.bfloat16).float().numpy()
sc = 1.0 / np.sqrt(H)
w_ih, w_hh = q((rs.uniform(-1, 1, (3 * H, cin)) * 2 * sc).astype(np.float32)), q((rs.uniform(-1, 1, (3 * H, H)) * 2 * sc).astype(np.float32))
b_ih, b_hh = (rs.uniform(-1, 1, 3 * H) * sc).astype(np.float32), (rs.uniform(-1, 1, 3 * H) * sc).astype(np.float32)
x = q(rs.standard_normal((B, L, cin)).astype(np.float32))
ref = O.gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse=False)
R = lib.tae_gru_rows_per_block(B)
assert R == 128 and (B + R - 1) // R > 2 * 74, "want several block pairs per cluster"
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
xpad = np.zeros((B, L, 8), np.float32); xpad[:, :, :cin] = x
xt = torch.empty(lib.tae_gru_tile_bytes(B, L, 1, R), dtype=torch.uint8, device=DEV)
xd = t(xpad)
_lib.check(lib.tae_gru_tiles_from_f32(_lib.ptr(xd), _lib.ptr(xt), B, L, 8, R, _lib.stream_ptr()))
packed = torch.empty(lib.tae_gru_packed_bytes(H, cin, cin), dtype=torch.uint8, device=DEV)
wd = [t(w_ih), t(w_hh), t(b_ih), t(b_hh)]
_lib.check(lib.tae_gru_pack_bf16(*[_lib.ptr(v) for v in wd], _lib.ptr(packed), H, cin, cin, _lib.stream_ptr()))
out = torch.zeros(lib.tae_gru_tile_bytes(B, L, 26, R), dtype=torch.uint8, device=DEV)
ws = torch.zeros(256, dtype=torch.uint8, device=DEV)
_lib.check(lib.tae_gru_direction_bf16(_lib.ptr(packed), _lib.ptr(xt), _lib.ptr(out), B, L, H, cin, cin, R, 26, 0, 0, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
got = np.zeros((B, L, H), np.float32)
for f0 in range(0, H, 8):
    wsel = torch.zeros(8, 2 * H, device=DEV)
    for f in range(min(8, H - f0)):
        wsel[f, f0 + f] = 1.0
    y8 = torch.empty(B, L, 8, device=DEV)
    _lib.check(lib.tae_gru_linear_f32(_lib.ptr(out), _lib.ptr(wsel), _lib.ptr(torch.zeros(8, device=DEV)), _lib.ptr(y8), B, L, 2 * H, H, 8, R, _lib.stream_ptr()))
    torch.cuda.synchronize()
    got[:, :, f0:f0 + 8] = y8.cpu().numpy()[:, :, :min(8, H - f0)]
err = np.abs(got - ref)
per_pair = err.reshape(-1, 256 if B % 256 == 0 else 1, L, H).max(axis=(1, 2, 3)) if B % 256 == 0 else None
print("multi-pair GRU: max |dh| %.3e mean %.3e (expected < 2e-2)" % (err.max(), err.mean()))
if per_pair is not None:
    print("worst block pairs:", np.argsort(-per_pair)[:5], per_pair.max())
assert err.max() < 2e-2
print("OK")
