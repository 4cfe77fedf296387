"""Error of the three decode precisions against the numpy oracle (fp32 restatement of the reference, itself within 3e-6 of an fp64
evaluation) on seeded batches: max / 99.99th percentile / count above 1e-4 of |posterior difference| and of the codes."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_codec, gen_inputs  # noqa: E402
from oracle import turboae_oracle as O  # noqa: E402


def main():
    out = {}
    for cfg, B, snr in (("c1", 297, 0.0), ("c1", 3000, 0.0), ("c1", 400, 2.0), ("c3", 1000, 0.0)):
        m, w, p = build_codec(cfg)
        u, noise = gen_inputs(900 + B, B, 100, snr)
        ref_codes = O.enc_forward(u, w, p)
        r = (ref_codes + noise).astype(np.float32)
        ref_y = O.dec_forward(r, w, p)
        rd = torch.from_numpy(r).cuda()
        ud = torch.from_numpy(u).cuda()
        with torch.no_grad():
            for prec in ("f16x3", "fp32") + (("bf16",) if B <= 400 else ()):
                m.enc.precision = prec
                c = m.enc(ud).cpu().numpy()
                y = m.dec.decode(rd, precision=prec).cpu().numpy()
                e, ec = np.abs(y - ref_y), np.abs(c - ref_codes)
                out["%s_B%d_%s" % (cfg, B, prec)] = {"y_max": float(e.max()), "y_p9999": float(np.quantile(e, 0.9999)),
                                                      "y_n_gt_1e-4": int((e > 1e-4).sum()), "codes_max": float(ec.max()),
                                                      "hard_flips": int((np.round(y) != np.round(ref_y)).sum())}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
