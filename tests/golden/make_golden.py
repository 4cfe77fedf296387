"""Generate the golden fixtures under tests/golden/ by executing the UNMODIFIED
reference (/root/reference) on CPU in the build container.

    python tests/golden/make_golden.py [--sweep-blocks 10000]

The reference cannot travel to the GPU box, so its outputs are committed here:

  weights_c1.npz   state_dict of models/dta_cont_cnn2_cnn5_enctrain2_dectrainneg15_2.pt (enc2/dec5)
  weights_c3.npz   state_dict of models/enc5_dec5_cont_1dBenc.pt                        (enc5/dec5)
  kat_c1_b4.npz    SURVEY.md Appendix C known-answer run (torch.manual_seed(0), B=4)
  io_c1_b8.npz     B=8 @ 0 dB, numpy-RandomState inputs, with every dec Linear output (hooks)
  io_c3_b6.npz     same for the enc5/dec5 checkpoint, B=6 @ 1 dB
  weights_c1s.npz / io_c1s_b8.npz   the binarised-code checkpoint dta_steq2_... (train_channel_mode block_norm_ste), B=8 @ 2 dB
  rnn_*.npz        DEC_LargeRNN (bi-GRU decoder) runs with seeded default-init weights: weights, input, output
  perm.npz         interleaver goldens (p, inverse, gather of arange through the reference modules)
  grad_c1_b6.npz   one training step (forward, clamp, BCE, backward) of the reference with checkpoint c1, B=6 @ -1.5 dB:
                   loss, per-parameter gradient norms / first values, six gradients in full
  flags_c1_b8.npz  -precompute_norm_stats (three successive batches, running scalars) and -is_variable_block_len (block length 40)
  dense_*.npz      DEC_LargeCNN with DenseSameShapeConv1d stacks (-encoder TurboAE_rate3_cnn_dense), seeded default init x2
  ber_c1.json      12-point BER/BLER sweep (reference trainer.py:157-178 loop restated with seeded
                   numpy inputs, batch 500) -- per-point bit/block error counts

All random inputs come from numpy's legacy RandomState (stream-stable by
numpy's compatibility policy) except the Appendix-C KAT, which stores the torch
tensors themselves.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import compat  # noqa: E402

C1_ARGS = ["-encoder", "TurboAE_rate3_cnn", "-decoder", "TurboAE_rate3_cnn", "-enc_num_unit", "100",
           "-enc_num_layer", "2", "-enc_kernel_size", "5", "-dec_num_layer", "5", "-dec_num_unit", "100",
           "-dec_kernel_size", "5", "-channel", "awgn", "-num_train_dec", "5", "-num_train_enc", "1",
           "-code_rate_k", "1", "-code_rate_n", "3", "-block_len", "100", "--no-cuda"]
C3_ARGS = [a if a != "2" or C1_ARGS[i - 1] != "-enc_num_layer" else "5" for i, a in enumerate(C1_ARGS)]
C1S_ARGS = C1_ARGS + ["-train_channel_mode", "block_norm_ste", "-test_channel_mode", "block_norm_ste"]      # README.md:84-88
CKPT = {"c1": "models/dta_cont_cnn2_cnn5_enctrain2_dectrainneg15_2.pt", "c3": "models/enc5_dec5_cont_1dBenc.pt",
        "c1s": "models/dta_steq2_cnn2_cnn5_enctrain2_dectrainneg15_2.pt"}


def build_reference_model(cfg, batch_size):
    compat.install()
    args = compat.reference_args({"c1": C1_ARGS, "c3": C3_ARGS, "c1s": C1S_ARGS}[cfg] + ["-batch_size", str(batch_size)])
    from numpy import arange
    from numpy.random import mtrand
    from encoders import ENC_interCNN          # reference main.py:35-36
    from decoders import DEC_LargeCNN          # reference main.py:75-76
    from channel_ae import Channel_AE
    p_array = mtrand.RandomState(0).permutation(arange(args.block_len))   # main.py:123-127
    enc, dec = ENC_interCNN(args, p_array), DEC_LargeCNN(args, p_array)
    enc.set_parallel(); dec.set_parallel()      # main.py:157-159 (gives the '.module.' keys)
    model = Channel_AE(args, enc, dec)
    sd = torch.load(os.path.join(compat.REFERENCE_ROOT, CKPT[cfg]), weights_only=True)
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model, args, p_array


def gen_inputs(seed, B, L, snr_db):
    rs = np.random.mtrand.RandomState(seed)
    u = rs.randint(0, 2, size=(B, L, 1)).astype(np.float32)
    sigma = 10 ** (-snr_db / 20.0)
    noise = (sigma * rs.standard_normal((B, L, 3))).astype(np.float32)
    return u, noise


def dump_weights(cfg):
    model, _, _ = build_reference_model(cfg, 4)
    sd = {k: v.detach().cpu().numpy().astype(np.float32) for k, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, "weights_%s.npz" % cfg), **sd)
    print("weights_%s.npz: %d tensors, %d params" % (cfg, len(sd), sum(v.size for v in sd.values())))


def dump_io(cfg, B, seed, snr_db, name):
    model, args, p = build_reference_model(cfg, B)
    u, noise = gen_inputs(seed, B, args.block_len, snr_db)
    lin_out = []
    hooks = []
    for i in range(args.num_iteration):
        for mod in (model.dec.dec1_outputs[i], model.dec.dec2_outputs[i]):
            hooks.append(mod.register_forward_hook(lambda m, a, o: lin_out.append(o.detach().numpy().copy())))
    with torch.no_grad():
        y, codes = model(torch.from_numpy(u), torch.from_numpy(noise))
        x_tx = None
    for h in hooks:
        h.remove()
    out = dict(u=u, noise=noise, codes=codes.numpy(), received=(codes + torch.from_numpy(noise)).numpy(),
               y=y.numpy(), p=np.asarray(p, dtype=np.int64), snr_db=np.float32(snr_db))
    assert len(lin_out) == 2 * args.num_iteration
    for j, a in enumerate(lin_out):
        out["lin_%02d" % j] = a
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "BER", float(np.mean(np.round(y.numpy()) != u)))


def dump_kat():
    model, args, p = build_reference_model("c1", 4)
    torch.manual_seed(0)
    X = torch.randint(0, 2, (4, 100, 1), dtype=torch.float)
    noise = torch.randn(4, 100, 3)
    with torch.no_grad():
        y, codes = model(X, noise)
    np.savez_compressed(os.path.join(HERE, "kat_c1_b4.npz"), X=X.numpy(), noise=noise.numpy(),
                        codes=codes.numpy(), y=y.numpy(), p=np.asarray(p, dtype=np.int64))
    print("KAT bit errors", int((torch.round(y) != X).sum()), "of 400; y[0,:4]", y[0, :4, 0].tolist())


def dump_perm():
    compat.install()
    from interleavers import Interleaver, DeInterleaver
    from numpy import arange
    from numpy.random import mtrand
    out = {}
    for L, seed in ((100, 0), (100, 7), (10, 0), (1000, 0), (1, 0)):
        p = mtrand.RandomState(seed).permutation(arange(L))
        x = torch.arange(3 * L * 5, dtype=torch.float32).view(3, L, 5)
        out["p_%d_%d" % (L, seed)] = np.asarray(p, dtype=np.int64)
        out["fwd_%d_%d" % (L, seed)] = Interleaver(None, p)(x).contiguous().numpy()
        out["inv_%d_%d" % (L, seed)] = DeInterleaver(None, p)(x).contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, "perm.npz"), **out)
    print("perm.npz written")


def dump_rnn(name, B, L, H, n_iter, seed):
    """DEC_LargeRNN (reference decoders.py:16-149): no checkpoint is shipped, so weights are torch.manual_seed-ed default init."""
    compat.install()
    args = compat.reference_args(["-encoder", "TurboAE_rate3_cnn", "-decoder", "TurboAE_rate3_rnn", "-dec_rnn", "gru",
                                  "-dec_num_unit", str(H), "-num_iteration", str(n_iter), "-block_len", str(L),
                                  "-batch_size", str(B), "-code_rate_k", "1", "-code_rate_n", "3", "--no-cuda"])
    from numpy import arange
    from numpy.random import mtrand
    from decoders import DEC_LargeRNN
    p_array = mtrand.RandomState(0).permutation(arange(L))
    torch.manual_seed(seed)
    dec = DEC_LargeRNN(args, p_array)
    dec.set_parallel()
    dec.eval()
    rs = np.random.mtrand.RandomState(seed)
    received = (rs.randint(0, 2, size=(B, L, 3)) * 2.0 - 1.0 + 0.8 * rs.standard_normal((B, L, 3))).astype(np.float32)
    with torch.no_grad():
        y = dec(torch.from_numpy(received)).numpy()
    out = {"dec." + k: v.detach().numpy().astype(np.float32) for k, v in dec.state_dict().items()}
    out.update(received=received, y=y, p=np.asarray(p_array, dtype=np.int64), cfg=np.array([B, L, H, n_iter], dtype=np.int64))
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "params", sum(v.size for k, v in out.items() if k.startswith("dec.")), "y range", float(y.min()), float(y.max()))


def dump_ber(blocks, batch=500):
    model, args, p = build_reference_model("c1", batch)
    snrs = [-1.5 + 0.5 * i for i in range(12)]          # trainer.py:157-158 with the default flags
    res = {"batch": batch, "blocks": blocks, "snrs": snrs, "bit_errors": [], "block_errors": [], "first_batch_bit_errors": [],
           "seed_rule": "seed = 100000 + 1000*snr_index + batch_index; gen_inputs() in make_golden.py"}
    for si, snr in enumerate(snrs):
        be = ble = 0
        for bi in range(blocks // batch):
            u, noise = gen_inputs(100000 + 1000 * si + bi, batch, args.block_len, snr)
            with torch.no_grad():
                y, _ = model(torch.from_numpy(u), torch.from_numpy(noise))
            wrong = np.round(y.numpy()) != u
            be += int(wrong.sum()); ble += int(wrong.reshape(batch, -1).any(axis=1).sum())
            if bi == 0:
                res["first_batch_bit_errors"].append(int(wrong.sum()))
        res["bit_errors"].append(be); res["block_errors"].append(ble)
        print("snr %+.1f dB  BER %.6f  BLER %.5f" % (snr, be / (blocks * 100.0), ble / float(blocks)), flush=True)
    json.dump(res, open(os.path.join(HERE, "ber_c1.json"), "w"), indent=1)


def dump_grad(cfg, B, seed, snr_db, name):
    """One trainer.train step of the UNMODIFIED reference on CPU (reference trainer.py:53-74: forward through Channel_AE, clamp,
    BCE (loss.py:32-35), loss.backward()) with the shipped checkpoint and seeded inputs: loss, every parameter gradient's
    L2 norm and first 16 values, and a few gradients in full.  Pins the training arithmetic of the oracle (row f1)."""
    import torch.nn.functional as F
    model, args, p_array = build_reference_model(cfg, B)
    model.train()
    u, noise = gen_inputs(seed, B, args.block_len, snr_db)
    out, codes = model(torch.from_numpy(u), torch.from_numpy(noise))
    loss = F.binary_cross_entropy(torch.clamp(out, 0.0, 1.0), torch.from_numpy(u))
    loss.backward()
    d = {"u": u, "noise": noise, "p": np.asarray(p_array), "loss": np.float64(float(loss)), "snr_db": np.float64(snr_db)}
    full = ("enc.enc_cnn_1.module.cnns.0.weight", "enc.enc_linear_3.module.weight", "dec.dec1_cnns.0.module.cnns.0.weight",
            "dec.dec2_cnns.3.module.cnns.2.bias", "dec.dec1_outputs.5.module.weight", "dec.dec2_outputs.5.module.weight")
    names = []
    for k, v in model.named_parameters():
        g = v.grad.detach().numpy().astype(np.float32)
        names.append(k)
        d["norm/" + k] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
        d["head/" + k] = g.reshape(-1)[:16].copy()
        if k in full:
            d["full/" + k] = g
    np.savez_compressed(os.path.join(HERE, name), **d)
    print("%s: loss %.6f, %d parameter gradients" % (name, float(loss), len(names)))


def dump_dense(name, B=4, L=30, units=20, n_layer=3, n_iter=2, seed=5):
    """DEC_LargeCNN built from DenseSameShapeConv1d (reference decoders.py:173-176: every -encoder other than TurboAE_rate3_cnn;
    cnn_utils.py:49-82): no checkpoint is shipped, so weights are torch.manual_seed-ed default init x2; weights, input, output."""
    compat.install()
    args = compat.reference_args(["-encoder", "TurboAE_rate3_cnn_dense", "-decoder", "TurboAE_rate3_cnn", "-dec_num_unit", str(units),
                                  "-dec_num_layer", str(n_layer), "-dec_kernel_size", "5", "-num_iteration", str(n_iter), "-block_len", str(L),
                                  "-batch_size", str(B), "-code_rate_k", "1", "-code_rate_n", "3", "--no-cuda"])
    from numpy import arange
    from numpy.random import mtrand
    from decoders import DEC_LargeCNN
    p = mtrand.RandomState(0).permutation(arange(L))
    torch.manual_seed(seed)
    dec = DEC_LargeCNN(args, p)
    dec.set_parallel()
    with torch.no_grad():
        for q in dec.parameters():
            q.mul_(2.0)
    dec.eval()
    rs = np.random.mtrand.RandomState(seed)
    rec = (rs.randint(0, 2, size=(B, L, 3)) * 2.0 - 1.0 + 0.7 * rs.standard_normal((B, L, 3))).astype(np.float32)
    with torch.no_grad():
        y = dec(torch.from_numpy(rec)).numpy().astype(np.float32)
    out = {"dec." + k: v.detach().numpy().astype(np.float32) for k, v in dec.state_dict().items()}
    out.update(received=rec, y=y, p=np.asarray(p, np.int64), cfg=np.array([B, L, units, n_layer, n_iter], np.int64))
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, y.shape, float(y.min()), float(y.max()), [tuple(v.shape) for k, v in out.items() if "cnns.0.module.cnns" in k and "weight" in k][:3])


def dump_dense_enc(name, B=4, L=30, units=20, n_layer=3, seed=6):
    """ENC_interCNN built from DenseSameShapeConv1d (reference encoders.py:313-327: -encoder TurboAE_rate3_cnn_dense): no checkpoint
    is shipped, so weights are torch.manual_seed-ed default init x2; weights, bits, codes."""
    compat.install()
    args = compat.reference_args(["-encoder", "TurboAE_rate3_cnn_dense", "-decoder", "TurboAE_rate3_cnn", "-enc_num_unit", str(units),
                                  "-enc_num_layer", str(n_layer), "-enc_kernel_size", "5", "-block_len", str(L), "-batch_size", str(B),
                                  "-code_rate_k", "1", "-code_rate_n", "3", "--no-cuda"])
    from numpy import arange
    from numpy.random import mtrand
    from encoders import ENC_interCNN
    p = mtrand.RandomState(0).permutation(arange(L))
    torch.manual_seed(seed)
    enc = ENC_interCNN(args, p)
    enc.set_parallel()
    with torch.no_grad():
        for q in enc.parameters():
            q.mul_(2.0)
    enc.eval()
    u = np.random.mtrand.RandomState(seed).randint(0, 2, size=(B, L, 1)).astype(np.float32)
    with torch.no_grad():
        codes = enc(torch.from_numpy(u)).numpy().astype(np.float32)
    out = {"enc." + k: v.detach().numpy().astype(np.float32) for k, v in enc.state_dict().items()}
    out.update(u=u, codes=codes, p=np.asarray(p, np.int64), cfg=np.array([B, L, units, n_layer], np.int64))
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, codes.shape, float(codes.mean()), float(codes.std()))


def dump_flags(name):
    """Two non-default branches of the hot path, executed on the UNMODIFIED reference with checkpoint c1:
    -precompute_norm_stats (encoders.py:110-114): codes of three successive batches + the running scalars after each;
    -is_variable_block_len (encoders.py:353-360, decoders.py:208-215): a batch of block length 40 through enc + dec."""
    B = 8
    model, args, p = build_reference_model("c1", B)
    out = {}
    args.precompute_norm_stats = True
    model.enc.reset_precomp()
    with torch.no_grad():
        for i, seed in enumerate((11, 12, 13)):
            u, _ = gen_inputs(seed, B, args.block_len, 0.0)
            out["run_codes_%d" % i] = model.enc(torch.from_numpy(u)).numpy().astype(np.float32)
            out["run_scalars_%d" % i] = np.array([float(model.enc.mean_scalar), float(model.enc.std_scalar), model.enc.num_test_block], np.float64)
    args.precompute_norm_stats = False
    args.is_variable_block_len = True
    L = 40
    u, noise = gen_inputs(77, B, L, 1.0)
    with torch.no_grad():
        codes = model.enc(torch.from_numpy(u))
        y = model.dec(codes + torch.from_numpy(noise))
    out.update(var_u=u, var_noise=noise, var_codes=codes.numpy().astype(np.float32), var_y=y.numpy().astype(np.float32),
               var_p=np.asarray(model.dec.interleaver.p_array.numpy(), np.int64))
    args.is_variable_block_len = False
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep-blocks", type=int, default=10000)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    todo = a.only.split(",") if a.only else ["weights", "kat", "io", "perm", "rnn", "ber", "grad", "flags", "dense"]
    if "weights" in todo:
        dump_weights("c1"); dump_weights("c3"); dump_weights("c1s")
    if "kat" in todo:
        dump_kat()
    if "io" in todo:
        dump_io("c1", 8, 4242, 0.0, "io_c1_b8.npz"); dump_io("c3", 6, 777, 1.0, "io_c3_b6.npz")
        dump_io("c1s", 8, 999, 2.0, "io_c1s_b8.npz")
    if "perm" in todo:
        dump_perm()
    if "rnn" in todo:
        dump_rnn("rnn_h32_i2_l40_b5.npz", 5, 40, 32, 2, 11); dump_rnn("rnn_h100_i1_l100_b3.npz", 3, 100, 100, 1, 12)
    if "ber" in todo:
        dump_ber(a.sweep_blocks)
    if "grad" in todo:
        dump_grad("c1", 6, 2718, -1.5, "grad_c1_b6.npz")
    if "flags" in todo:
        dump_flags("flags_c1_b8.npz")
    if "dense" in todo:
        dump_dense("dense_u20_l3_i2_b4.npz")
        dump_dense_enc("dense_enc_u20_l3_b4.npz")
