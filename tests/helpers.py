"""Shared test plumbing: reference-shaped ``args``, codec construction from the committed weight fixtures and
the seeded input generator of tests/golden/make_golden.py.  Nothing here touches /root/reference."""
import os
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def make_args(**over):
    """The fields of the reference's argparse Namespace (get_args.py:8-229) the hot path reads, with the
    reference defaults / BASELINE config-1 values."""
    a = dict(encoder="TurboAE_rate3_cnn", decoder="TurboAE_rate3_cnn", enc_num_unit=100, enc_num_layer=2,
             enc_kernel_size=5, dec_num_unit=100, dec_num_layer=5, dec_kernel_size=5, num_iteration=6, num_iter_ft=5,
             extrinsic=1, code_rate_k=1, code_rate_n=3, block_len=100, batch_size=500, no_cuda=False,
             is_parallel=1, enc_act="elu", dec_act="linear", no_code_norm=False, enc_truncate_limit=0.0,
             precompute_norm_stats=False, is_variable_block_len=False, train_channel_mode="block_norm",
             test_channel_mode="block_norm", is_interleave=1, dropout=0.0, tae_precision=None,
             enc_quantize_level=2, enc_value_limit=1.0, enc_grad_limit=0.01, enc_clipping="both", dec_rnn="gru")
    a.update(over)
    return SimpleNamespace(**a)


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


class Codec(torch.nn.Module):
    """`enc` / `dec` children like the reference's Channel_AE (channel_ae.py:10-18), so the shipped
    checkpoints' keys ('enc.…', 'dec.…') load with strict=True."""

    def __init__(self, args, p):
        super().__init__()
        import turboae_b200 as T
        self.enc = T.ENC_interCNN(args, p)
        self.dec = T.DEC_LargeCNN(args, p)
        self.enc.set_parallel()
        self.dec.set_parallel()


def build_codec(cfg="c1", device="cuda", **over):
    from oracle import turboae_oracle as O
    w = load_npz("weights_%s.npz" % cfg)
    if cfg == "c1s":                                  # the binarised checkpoint (dta_steq2_...): README.md:84-88
        over.setdefault("train_channel_mode", "block_norm_ste")
    args = make_args(enc_num_layer=5 if cfg == "c3" else 2, **over)
    p = O.make_perm(args.block_len, 0)
    m = Codec(args, p)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    m = m.to(device).eval()
    return m, w, p


def gen_inputs(seed, B, L, snr_db):
    """tests/golden/make_golden.py::gen_inputs (numpy legacy RandomState: stream-stable)."""
    rs = np.random.mtrand.RandomState(seed)
    u = rs.randint(0, 2, size=(B, L, 1)).astype(np.float32)
    sigma = 10 ** (-snr_db / 20.0)
    noise = (sigma * rs.standard_normal((B, L, 3))).astype(np.float32)
    return u, noise


def torch_rnn_modules(weights, n_iter, H, F=5, scale=1.0):
    """torch.nn.GRU / Linear modules (CPU: the reference's own operators, decoders.py:43-58) loaded from fixture-style keys
    'dec.dec{1,2}_rnns.<i>.module.<name>' / 'dec.dec{1,2}_outputs.<i>.module.<name>'."""
    rnns, outs = ([], []), ([], [])
    for i in range(n_iter):
        for s_ in (0, 1):
            gru = torch.nn.GRU(2 + F, H, num_layers=2, bias=True, batch_first=True, bidirectional=True)
            pre = "dec.dec%d_rnns.%d.module." % (s_ + 1, i)
            gru.load_state_dict({k[len(pre):]: torch.from_numpy(np.asarray(v) * np.float32(scale)) for k, v in weights.items() if k.startswith(pre)})
            pre = "dec.dec%d_outputs.%d.module." % (s_ + 1, i)
            w = {k[len(pre):]: torch.from_numpy(np.asarray(v) * np.float32(scale)) for k, v in weights.items() if k.startswith(pre)}
            lin = torch.nn.Linear(2 * H, w["weight"].shape[0])
            lin.load_state_dict(w)
            rnns[s_].append(gru)
            outs[s_].append(lin)
    return rnns[0], rnns[1], outs[0], outs[1]
