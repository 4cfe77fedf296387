#!/usr/bin/env python
"""clock64 timeline of one CTA pair of the fused decoder (development tool; run on the GPU box with a library built with
-DTAE_TIMELINE=1: `TURBOAE_B200_LIB=.../libtae_timeline.so TURBOAE_B200_NVCC_FLAGS=-DTAE_TIMELINE=1 python -m turboae_b200.build`).

Writes gpurun_out/<TAG>.md: per (layer, tile) of one stack in steady state -- issuer wait, issue time, when the accumulators
completed, epilogue duration -- and where the tensor pipe idles: the bubble at every layer transition of the first group
(completion of tile 0 of step s+1 minus completion of tile 3 of step s minus the MMA time of that tile)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import make_args
import turboae_b200 as T
from turboae_b200 import _lib
from oracle import turboae_oracle as O

TAG = os.environ.get("TL_TAG", "dec_timeline")
B = int(os.environ.get("TL_B", "50000"))
args = make_args(batch_size=B)
dec = T.DEC_LargeCNN(args, O.make_perm(100, 0)).cuda().eval()
rec = torch.randn(B, 100, 3, device="cuda")
N_STEP = 72
tl = torch.zeros(N_STEP * 4 * 8 + 148 * 4 + 128 + 5 * 4 * 2 * 16 * 2, dtype=torch.int64, device="cuda")
lib = _lib.load()
with torch.no_grad():
    dec(rec); torch.cuda.synchronize()
    lib.tae_debug_set_timeline(_lib.ptr(tl))
    dec(rec); torch.cuda.synchronize()
    lib.tae_debug_set_timeline(None)
full = tl.cpu().numpy()
t = full[:N_STEP * 4 * 8].reshape(N_STEP, 4, 8).astype(np.int64)
cta = full[N_STEP * 4 * 8:N_STEP * 4 * 8 + 148 * 4].reshape(148, 4)
gs = full[N_STEP * 4 * 8 + 148 * 4:N_STEP * 4 * 8 + 148 * 4 + 128]
wt = full[N_STEP * 4 * 8 + 148 * 4 + 128:].reshape(5, 4, 2, 16, 2).astype(np.int64)
out = []
P = out.append
n_groups = int((gs > 0).sum())
gd = np.diff(gs[:n_groups])
P("# fused decoder timeline (%s), cluster 0, B = %d" % (TAG, B))
P("")
P("groups run by cluster 0: %d; cycles per group: min %d median %d max %d" % (n_groups, gd.min(), int(np.median(gd)), gd.max()))
start = cta[:, 0].min()
dur = (cta[:, 1] - cta[:, 0])
P("per-CTA duration us min/mean/max %.1f / %.1f / %.1f ; clock %.3f GHz ; CTA start skew %.1f us, end skew %.1f us" % (
    dur.min() / 1e3, dur.mean() / 1e3, dur.max() / 1e3, cta[:, 2].mean() / dur.mean(),
    (cta[:, 0] - start).max() / 1e3, ((cta[:, 1] - start).max() - (cta[:, 1] - start).min()) / 1e3))
# columns: 0 mma_wait0, 1 mma_wait1 (inputs ready), 2 mma_issued, 3 e0_wait0, 4 e0_acc, 5 e0_done (arrived), 6 e0 tmem loads done, 7 e0 stores done
gi = 2 if n_groups > 3 else 0          # the kernel stamps the third group of cluster 0 (steady state)
t0 = t[0, 0, 0]
MMA = {0: 3 * 56, 5: 7 * 8}          # cycles of a tile's MMAs: layer 0 (3 k-steps of N=112), Linear (7 k-steps of N=16)
def mma_cycles(layer):
    return MMA.get(layer, 32 * 56)
ideal = 12 * 4 * sum(mma_cycles(l) for l in range(6))
P("ideal tensor cycles per group (MMA issue slots only): %d ; measured median group %d => busy fraction %.3f" % (ideal, int(np.median(gd)), ideal / np.median(gd)))
P("")
P("## one stack in steady state (stack 6 of the first group; cycles relative to the stack's first issuer wait)")
P("")
P("| layer | tile | issuer wait | issue | acc ready (epi saw) | epi wait | epi tmem ld | epi compute+store | epi total |")
P("|---|---|---|---|---|---|---|---|---|")
st = 6
base = t[st * 6, 0, 0]
for layer in range(6):
    s = st * 6 + layer
    for m in range(4):
        r = t[s, m]
        P("| %d | %d | %d..%d (%d) | ..%d (%d) | %d | %d | %d | %d | %d |" % (
            layer, m, r[0] - base, r[1] - base, r[1] - r[0], r[2] - base, r[2] - r[1], r[4] - base if r[4] else -1,
            r[4] - r[3] if r[4] else -1, r[6] - r[4] if r[6] else -1, r[7] - r[6] if r[7] and r[6] else -1, r[5] - r[4] if r[5] else -1))
P("")
P("## tensor-pipe bubbles of the first group: acc-ready(step s+1, tile 0) - acc-ready(step s, tile 3) - MMA cycles of that tile")
P("")
acc = t[:, :, 4]
rows = []
for s in range(N_STEP - 1):
    layer_next = (s + 1) % 6
    # layers whose epilogue warp 0 does not stamp per tile (Linear) use tile 0's stamp for all tiles
    a3 = acc[s, 3] if acc[s, 3] else acc[s, 0]
    a0n = acc[s + 1, 0]
    rows.append((s, s % 6, a0n - a3 - mma_cycles(layer_next)))
by_kind = {}
for s, layer, b in rows:
    by_kind.setdefault(layer, []).append(b)
P("| transition after layer | count | mean bubble | min | max |")
P("|---|---|---|---|---|")
for layer in sorted(by_kind):
    v = np.array(by_kind[layer])
    P("| %d -> %d | %d | %.0f | %d | %d |" % (layer, (layer + 1) % 6, len(v), v.mean(), v.min(), v.max()))
tot = sum(b for _, _, b in rows)
P("")
P("sum of bubbles over the group: %d cycles (%.1f %% of the group); group prologue (group start -> first inputs ready): %d ; last issue -> next group start: %d" % (
    tot, 100.0 * tot / np.median(gd), t[0, 0, 1] - gs[gi], gs[gi + 1] - t[N_STEP - 1, 3, 2]))
P("")
P("## per-tile within-layer gaps (conv layers): acc-ready(tile m+1) - acc-ready(tile m) - 1792")
g = []
for s in range(N_STEP):
    if s % 6 in (1, 2, 3, 4):
        for m in range(3):
            g.append(acc[s, m + 1] - acc[s, m] - 32 * 56)
g = np.array(g)
P("mean %.0f, p50 %d, p90 %d, max %d cycles (x %d tile transitions per group = %d cycles)" % (g.mean(), np.percentile(g, 50), np.percentile(g, 90), g.max(), len(g), g.sum()))
P("")
P("## per-warp epilogue of stack 6 (leader CTA / peer CTA; each SM has its own clock64): duration acc-seen -> reported, per column part (mean over the 4 lane quadrants), and spread of the report times")
P("")
P("| layer | tile | leader: part 0 / 1 / 2 / 3 duration | leader: last report - first report | leader: last report - warp 0 report | peer: part 0 / 1 / 2 / 3 duration | peer spread |")
P("|---|---|---|---|---|---|---|")
for layer in range(5):
    for m in range(4):
        row = []
        for cta in range(2):
            w = wt[layer, m, cta]
            d = (w[:, 1] - w[:, 0]).reshape(4, 4)          # ew = part * 4 + quadrant-ish (warp & 3)
            row.append(" / ".join("%d" % x for x in d.mean(axis=1)))
            row.append("%d" % (w[:, 1].max() - w[:, 1].min()))
            if cta == 0:
                row.append("%d" % (w[:, 1].max() - w[0, 1]))
        P("| %d | %d | %s |" % (layer, m, " | ".join(row)))
text = "\n".join(out)
print(text)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", TAG + ".md"), "w").write(text + "\n")
np.save(os.path.join(ROOT, "gpurun_out", TAG + ".npy"), full)
