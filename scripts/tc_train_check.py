#!/usr/bin/env python
"""GPU bring-up checks of the tensor-core training path (run under gpurun): (1) the weight-gradient kernel on synthetic
group images against torch einsum, (2) forward-with-stash against the inference kernel and the fp32 activations,
(3) decoder gradients against the fp32 CUDA-core training path."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.nn.functional as Fn
from helpers import build_codec, gen_inputs
from turboae_b200 import _lib, train_tc

dev = torch.device("cuda:0")
lib = _lib.load()
L, CW = 100, 5
ROWS, CB = 516, 8256


def to_image(x, groups):
    """x (B, L, C) float -> bf16 group image bytes [groups][ceil(C/8)][516][8]"""
    B, L_, Cc = x.shape
    nch = (Cc + 7) // 8
    img = torch.zeros(groups, nch, ROWS, 8, dtype=torch.bfloat16, device=x.device)
    xp = torch.zeros(B, L_, nch * 8, dtype=torch.float32, device=x.device)
    xp[:, :, :Cc] = x
    for b in range(B):
        g, c = divmod(b, CW)
        r0 = 2 + c * (L_ + 2)
        img[g, :, r0:r0 + L_, :] = xp[b].view(L_, nch, 8).permute(1, 0, 2).to(torch.bfloat16)
    return img


def from_image(img, B, Cc):
    groups, nch = img.shape[0], img.shape[1]
    out = torch.zeros(B, L, nch * 8, dtype=torch.float32, device=img.device)
    for b in range(B):
        g, c = divmod(b, CW)
        r0 = 2 + c * (L + 2)
        out[b] = img[g, :, r0:r0 + L, :].permute(1, 0, 2).reshape(L, nch * 8).float()
    return out[:, :, :Cc]


def check_wgrad():
    torch.manual_seed(0)
    B, units = 13, 100
    groups = lib.tae_train_groups(L, B)
    g = torch.randn(B, L, units, device=dev)
    x = torch.randn(B, L, units, device=dev)
    gi, xi = to_image(g, groups), to_image(x, groups)
    gq, xq = from_image(gi, B, units), from_image(xi, B, units)
    xpad = Fn.pad(xq, (0, 0, 2, 2))
    ref = torch.stack([torch.einsum("blo,blc->oc", gq, xpad[:, t:t + L]) for t in range(5)], dim=2)     # (o, c, t)
    ref_b = gq.sum(dim=(0, 1))
    res = {}
    for swap in (0, 1):
        lib.tae_debug_wgrad_swap(swap)
        dw = torch.zeros(units, units, 5, device=dev)
        db = torch.zeros(units, device=dev)
        jobs = []
        for (c0, nc, ncols, nv, bias) in ((0, 8, 64, 64, None), (8, 5, 48, 36, db.data_ptr())):
            jobs.append(_lib.TaeWgradJob(gi.data_ptr(), xi.data_ptr(), dw.data_ptr(), bias, 13, c0, nc, 5, ncols, units, nv, c0 * 8,
                                         5 * units, 5, 1, 0, groups, 0))
        try:
            train_tc.run_wgrad(jobs, dev)
            torch.cuda.synchronize()
            res[swap] = {"dw_max_err": float((dw - ref).abs().max()), "dw_ref_max": float(ref.abs().max()),
                         "db_max_err": float((db - ref_b).abs().max())}
        except Exception as e:   # noqa: BLE001
            res[swap] = {"error": str(e)}
    lib.tae_debug_wgrad_swap(0)
    return res


def check_train(B=23, fresh=False):
    m, w, p = build_codec("c1", batch_size=B)
    if fresh:                                  # default initialisation: loss ~ 0.69, well-conditioned gradients
        import turboae_b200 as T
        from helpers import make_args
        torch.manual_seed(0)
        m.dec = T.DEC_LargeCNN(make_args(batch_size=B), p).to(dev)
    u, noise = gen_inputs(99, B, L, 0.0)
    ud, nd = torch.from_numpy(u).to(dev), torch.from_numpy(noise).to(dev)
    with torch.no_grad():
        rec = (m.enc(ud) + nd).contiguous()
        y_inf = m.dec(rec)
    res = {}
    grads = {}
    for prec in ("fp32", "bf16"):
        m.dec.train_precision = prec
        m.zero_grad()
        r = rec.clone().requires_grad_(True)
        out = m.dec(r)
        loss = Fn.binary_cross_entropy(torch.clamp(out, 0.0, 1.0), ud)
        loss.backward()
        torch.cuda.synchronize()
        grads[prec] = ({k: v.grad.detach().clone() for k, v in m.dec.named_parameters()}, r.grad.detach().clone(), out.detach(), float(loss))
    res["out_vs_inference_max"] = float((grads["bf16"][2] - y_inf).abs().max())
    res["loss_fp32"], res["loss_bf16"] = grads["fp32"][3], grads["bf16"][3]
    gr, gb = grads["fp32"][1], grads["bf16"][1]
    res["d_received"] = {"rel_l2": float((gb - gr).norm() / gr.norm()), "cos": float(Fn.cosine_similarity(gb.flatten(), gr.flatten(), dim=0))}
    worst = []
    for k in grads["fp32"][0]:
        a, b = grads["fp32"][0][k], grads["bf16"][0][k]
        rel = float((a - b).norm() / (a.norm() + 1e-30))
        cos = float(Fn.cosine_similarity(a.flatten(), b.flatten(), dim=0))
        worst.append((rel, cos, k, float(a.norm())))
    worst.sort(reverse=True)
    res["params_worst"] = worst[:8]
    res["params_best"] = worst[-3:]
    res["n_bad"] = sum(1 for r_ in worst if r_[0] > 0.05)
    res["n_params"] = len(worst)
    # stash check: layer outputs of stack 0 against the fp32 conv stack
    buf = list(m.dec.__dict__["_tc_buffers"].values())[0]
    groups = buf.groups
    ysz = groups * 13 * CB
    with torch.no_grad():
        x0 = torch.cat([rec[:, :, 0:1], rec[:, :, 1:2], torch.zeros(B, L, 5, device=dev)], dim=2)
        h = x0
        errs = []
        from turboae_b200.cnn_utils import _conv_layer
        from turboae_b200._flat import unwrap
        for j, conv in enumerate(unwrap(m.dec.dec1_cnns[0]).cnns):
            h = _conv_layer(h.contiguous(), conv.weight.detach().contiguous(), conv.bias.detach().contiguous(), True)
            img = buf.stash_y[j * ysz:(j + 1) * ysz].view(torch.bfloat16).view(groups, 13, ROWS, 8)
            got = from_image(img, B, 100)
            errs.append(float((got - h).abs().max()))
        res["stash_y_stack0_max_err"] = errs
        ximg = buf.stash_x[:groups * CB].view(torch.bfloat16).view(groups, 1, ROWS, 8)
        res["stash_x_stack0_max_err"] = float((from_image(ximg, B, 7) - x0).abs().max())
    return res


def probe_wgrad():
    """Where does the hardware look for element (row, channel) of an MN-major operand?  One group, one tap, one job."""
    res = []
    for swap in (0, 1):
        lib.tae_debug_wgrad_swap(swap)
        for which in ("A", "B"):
            for (r, ch) in ((2, 0), (3, 0), (2, 1), (2, 8), (9, 0), (10, 3), (18, 0), (100, 50), (513, 63)):
                a = torch.ones(1, 13, ROWS, 8, dtype=torch.bfloat16, device=dev)
                b = torch.ones(1, 13, ROWS, 8, dtype=torch.bfloat16, device=dev)
                t = a if which == "A" else b
                t.zero_()
                t[0, ch // 8, r, ch % 8] = 1.0
                d = torch.zeros(104, 64, device=dev)
                job = _lib.TaeWgradJob(a.data_ptr(), b.data_ptr(), d.data_ptr(), None, 13, 0, 8, 1, 64, 104, 64, 0, 64, 1, 0, 0, 1, 0)
                train_tc.run_wgrad([job], dev)
                torch.cuda.synchronize()
                nz = d.nonzero().tolist()
                vec = d[:, 0] if which == "A" else d[0, :]
                hits = [(i, float(v)) for i, v in enumerate(vec.tolist()) if v != 0.0]
                res.append({"swap": swap, "op": which, "row": r, "ch": ch, "n_nonzero": len(nz), "hits": hits[:6]})
    lib.tae_debug_wgrad_swap(0)
    return res


if __name__ == "__main__":
    out = {}
    which = sys.argv[1:] or ["wgrad", "train"]
    if "wgrad" in which:
        out["wgrad"] = check_wgrad()
        print(json.dumps(out["wgrad"]), flush=True)
    if "probe" in which:
        for r in probe_wgrad():
            print(json.dumps(r), flush=True)
    if "train" in which:
        out["train"] = check_train()
    if "fresh" in which:
        out["fresh"] = check_train(B=203, fresh=True)
    print(json.dumps(out, indent=1))
