"""Two training steps (forward + backward + Adam) of enc2/dec5 on the tensor-core path, for compute-sanitizer / debugging:
python scripts/train_small.py [B]   (B = 745 exercises the two-launch backward with the overlapped weight gradients)"""
import os
import sys

import torch
import torch.nn.functional as Fn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_args  # noqa: E402
import turboae_b200 as T  # noqa: E402
from oracle import turboae_oracle as O  # noqa: E402  (permutation helper only)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 23
dev = torch.device("cuda", 0)
torch.manual_seed(3)
args = make_args(batch_size=B)
p = O.make_perm(100, 0)
enc, dec = T.ENC_interCNN(args, p).to(dev), T.DEC_LargeCNN(args, p).to(dev)
opt = torch.optim.Adam(list(enc.parameters()) + list(dec.parameters()), lr=1e-4, fused=True)
for _ in range(2):
    opt.zero_grad()
    u = torch.randint(0, 2, (B, 100, 1), device=dev).float()
    loss = Fn.binary_cross_entropy(torch.clamp(dec(enc(u) + torch.randn(B, 100, 3, device=dev)), 0.0, 1.0), u)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("train small run: B=%d loss %.4f" % (B, float(loss)))
