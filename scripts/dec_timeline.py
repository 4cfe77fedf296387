#!/usr/bin/env python
"""Dump the clock64 timeline of one CTA pair of the fused decoder (development tool; run on the GPU box)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import make_args
import turboae_b200 as T
from turboae_b200 import _lib
from oracle import turboae_oracle as O
B = 50000
args = make_args(batch_size=B)
dec = T.DEC_LargeCNN(args, O.make_perm(100, 0)).cuda().eval()
rec = torch.randn(B, 100, 3, device="cuda")
tl = torch.zeros(72 * 4 * 8 + 148 * 4 + 128, dtype=torch.int64, device="cuda")
lib = _lib.load()
with torch.no_grad():
    dec(rec); torch.cuda.synchronize()
    lib.tae_debug_set_timeline(_lib.ptr(tl))
    dec(rec); torch.cuda.synchronize()
    lib.tae_debug_set_timeline(None)
full = tl.cpu().numpy()
t = full[:72 * 4 * 8].reshape(72, 4, 8)
cta = full[72 * 4 * 8:72 * 4 * 8 + 148 * 4].reshape(148, 4)
gs = full[72 * 4 * 8 + 148 * 4:]
print('group start deltas (cycles):', np.diff(gs[:68]).tolist())
print('first group: start->first mma wait done', t[0,0,1]-gs[0], 'last mma issued -> next group start', gs[1]-t[71,3,2])
start = cta[:, 0].min()
print('per-CTA: start offset us (min/max) %.1f %.1f ; end offset us (min/max) %.1f %.1f ; duration us (min/mean/max) %.1f %.1f %.1f ; cycles (min/max) %d %d => clk GHz %.3f' % (
    (cta[:,0]-start).min()/1e3, (cta[:,0]-start).max()/1e3, (cta[:,1]-start).min()/1e3, (cta[:,1]-start).max()/1e3,
    (cta[:,1]-cta[:,0]).min()/1e3, (cta[:,1]-cta[:,0]).mean()/1e3, (cta[:,1]-cta[:,0]).max()/1e3, cta[:,2].min(), cta[:,2].max(),
    cta[:,2].mean()/ (cta[:,1]-cta[:,0]).mean()))
np.save(os.path.join(ROOT, 'gpurun_out', 'timeline_cta.npy'), cta)
t0 = t[0, 0, 0]
np.save(os.path.join(ROOT, "gpurun_out", "timeline.npy"), t)
names = ["mma_wait0", "mma_wait1", "mma_issued", "e0_wait0", "e0_acc", "e0_done", "e7_acc", "e7_done"]
for step in range(14, 16):
    for m in range(4):
        print("step %2d tile %d: " % (step, m) + " ".join("%s=%7d" % (n, t[step, m, k] - t0) for k, n in enumerate(names)))
