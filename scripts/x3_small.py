"""One small encode + decode through the split-operand kernel (for compute-sanitizer / debugging): python scripts/x3_small.py [B]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_codec, gen_inputs  # noqa: E402
from oracle import turboae_oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 5
m, w, p = build_codec("c1")
u, noise = gen_inputs(77, B, 100, 0.0)
m.enc.precision = "f16x3"
with torch.no_grad():
    codes = m.enc(torch.from_numpy(u).cuda())
    r = codes + torch.from_numpy(noise).cuda()
    y = m.dec.decode(r, precision="f16x3")
torch.cuda.synchronize()
ref = O.dec_forward(r.cpu().numpy(), w, p)
print("x3 small run: B=%d max|dcodes|=%.2e max|dy|=%.2e" % (B, float(np.abs(codes.cpu().numpy() - O.enc_forward(u, w, p)).max()),
                                                         float(np.abs(y.cpu().numpy() - ref).max())))
