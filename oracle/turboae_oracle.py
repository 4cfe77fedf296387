"""CPU restatement (numpy, float32) of the TurboAE hot path.

TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA path and the
timed CPU baseline of ``bench.py``; the product (``turboae_b200``) never
imports it and has no CPU fallback.

Parity status: PINNED.  The reference ships no golden vectors or tests for this
path, so the oracle is pinned against outputs of the
reference itself, executed unmodified in the build container by
``tests/golden/make_golden.py`` (fixtures committed under ``tests/golden/``;
``tests/test_oracle.py`` re-checks every one of them on CPU).

Each function cites the reference lines (under /root/reference) it restates.
Weights are passed as a dict ``name -> float32 ndarray`` using the reference's
``state_dict`` key names (``enc.enc_cnn_1.module.cnns.0.weight`` ...); the
``.module.`` infix produced by ``set_parallel()`` (reference encoders.py:343-349,
decoders.py:194-199) is optional.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- #
# permutation                                                                 #
# --------------------------------------------------------------------------- #
def make_perm(block_len: int, seed: int = 0) -> np.ndarray:
    """reference main.py:123-127 and commpy/channelcoding/interleavers.py:77-82:
    ``RandomState(seed).permutation(arange(block_len))`` (legacy numpy stream)."""
    return np.random.mtrand.RandomState(seed).permutation(np.arange(block_len)).astype(np.int64)


def inverse_perm(p: np.ndarray) -> np.ndarray:
    """reference interleavers.py:29-33: ``rp[p[i]] = i``."""
    p = np.asarray(p).reshape(-1)
    rp = np.zeros_like(p)
    for i in range(len(p)):
        rp[p[i]] = i
    return rp


def interleave(x: np.ndarray, p: np.ndarray) -> np.ndarray:
    """reference interleavers.py:15-21: ``out[b,i,f] = in[b,p[i],f]``."""
    return np.ascontiguousarray(x[:, np.asarray(p).reshape(-1), :])


def deinterleave(x: np.ndarray, p: np.ndarray) -> np.ndarray:
    """reference interleavers.py:43-48: ``out[b,i,f] = in[b,rp[i],f]``."""
    return np.ascontiguousarray(x[:, inverse_perm(p), :])


# --------------------------------------------------------------------------- #
# building blocks                                                             #
# --------------------------------------------------------------------------- #
def elu(z: np.ndarray) -> np.ndarray:
    """F.elu, alpha = 1 (reference cnn_utils.py:24-25, 43)."""
    z = z.astype(F32, copy=False)
    return np.where(z > 0, z, np.expm1(np.minimum(z, F32(0)))).astype(F32)


def conv1d_same(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """One ``Conv1d(k, stride 1, padding k//2)`` on a channel-last tensor.

    reference cnn_utils.py:15-22 (construction) and :36-46 (the transposes make
    the module act on ``(B, L, C)``):
    ``y[b,l,o] = bias[o] + sum_c sum_t W[o,c,t] * x[b,l+t-k//2,c]`` with zero
    padding (cross-correlation, as torch.nn.Conv1d).
    """
    B, L, cin = x.shape
    cout, cin_w, k = w.shape
    assert cin == cin_w
    pad = k // 2
    xp = np.zeros((B, L + 2 * pad, cin), dtype=F32)
    xp[:, pad:pad + L, :] = x
    # im2col: (B, L, k, cin)
    cols = np.stack([xp[:, t:t + L, :] for t in range(k)], axis=2).reshape(B * L, k * cin)
    wm = np.ascontiguousarray(np.transpose(w, (2, 1, 0)).reshape(k * cin, cout), dtype=F32)
    y = cols @ wm + b.astype(F32)[None, :]
    return y.reshape(B, L, cout).astype(F32, copy=False)


def same_shape_conv1d(x: np.ndarray, layers) -> np.ndarray:
    """reference cnn_utils.py:36-46: ``x <- ELU(conv_j(x))`` for every layer."""
    h = x.astype(F32, copy=False)
    for (w, b) in layers:
        h = elu(conv1d_same(h, w, b))
    return h


def dense_same_shape_conv1d(x: np.ndarray, layers) -> np.ndarray:
    """DenseSameShapeConv1d.forward (reference cnn_utils.py:67-82): layer idx sees [input, out_0, ..., out_{idx-1}]."""
    this_input, out = x, None
    for idx, (w, b) in enumerate(layers):
        if idx > 0:
            this_input = np.concatenate([this_input, out], axis=2)
        out = elu(conv1d_same(this_input, w, b))
    return out


def linear(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """torch.nn.Linear on the last axis: ``y = x @ W^T + b``."""
    B, L, c = x.shape
    y = x.reshape(B * L, c) @ np.ascontiguousarray(w.T, dtype=F32) + b.astype(F32)[None, :]
    return y.reshape(B, L, -1).astype(F32, copy=False)


def sigmoid(z: np.ndarray) -> np.ndarray:
    z = z.astype(F32, copy=False)
    out = np.empty_like(z)
    pos = z >= 0
    out[pos] = F32(1) / (F32(1) + np.exp(-z[pos]))
    ez = np.exp(z[~pos])
    out[~pos] = ez / (F32(1) + ez)
    return out


# --------------------------------------------------------------------------- #
# weight access                                                               #
# --------------------------------------------------------------------------- #
def _get(weights, key):
    if key in weights:
        return np.asarray(weights[key], dtype=F32)
    # tolerate presence/absence of the DataParallel ".module." infix
    alt = key.replace(".module.", ".")
    if alt in weights:
        return np.asarray(weights[alt], dtype=F32)
    raise KeyError(key)


def _stack(weights, prefix, n_layer):
    return [(_get(weights, "%s.module.cnns.%d.weight" % (prefix, j)),
             _get(weights, "%s.module.cnns.%d.bias" % (prefix, j))) for j in range(n_layer)]


def count_layers(weights, prefix) -> int:
    n = 0
    while ("%s.module.cnns.%d.weight" % (prefix, n) in weights
           or "%s.cnns.%d.weight" % (prefix, n) in weights):
        n += 1
    return n


# --------------------------------------------------------------------------- #
# encoder                                                                     #
# --------------------------------------------------------------------------- #
def enc_forward_unnormalised(u: np.ndarray, weights, p: np.ndarray, prefix: str = "enc", dense: bool = False) -> np.ndarray:
    """reference encoders.py:362-373: the three branches and the concat.

    ``u`` is ``(B, L, 1)`` in {0,1}; returns ``x_tx`` ``(B, L, 3)``.  ``dense``: the branches are DenseSameShapeConv1d stacks
    (reference encoders.py:322-330, -encoder TurboAE_rate3_cnn_dense).
    """
    x = (F32(2.0) * u.astype(F32) - F32(1.0))
    outs = []
    for i, inp in ((1, x), (2, x), (3, interleave(x, p))):
        n = count_layers(weights, "%s.enc_cnn_%d" % (prefix, i))
        h = (dense_same_shape_conv1d if dense else same_shape_conv1d)(inp, _stack(weights, "%s.enc_cnn_%d" % (prefix, i), n))
        y = linear(h, _get(weights, "%s.enc_linear_%d.module.weight" % (prefix, i)),
                   _get(weights, "%s.enc_linear_%d.module.bias" % (prefix, i)))
        outs.append(elu(y))                      # enc_act == 'elu' (encoders.py:86-100)
    return np.concatenate(outs, axis=2)


def power_constraint(x_tx: np.ndarray):
    """reference encoders.py:102-116 (default branch): ``(x - mean) / std`` over
    ALL elements of the batch, ``torch.std`` = unbiased (N-1)."""
    xd = x_tx.astype(np.float64)
    n = xd.size
    mean = xd.sum() / n
    var = ((xd - mean) ** 2).sum() / (n - 1)
    mean32 = F32(mean)
    std32 = F32(np.sqrt(var))
    return ((x_tx - mean32) * F32(1.0) / std32).astype(F32), mean32, std32


def power_constraint_running(x_tx: np.ndarray, state):
    """reference encoders.py:107-116 with -precompute_norm_stats (:110-114): running averages of the batch mean and
    unbiased std over the calls so far.  `state` = [mean_scalar, std_scalar, num_test_block] (float32, float32, float),
    updated in place like the module attributes (:76-84 start them at 0, 1, 0)."""
    this_mean = np.float32(x_tx.astype(np.float64).mean())
    this_std = np.float32(x_tx.astype(np.float64).std(ddof=1))
    state[2] += 1.0
    k = np.float32(state[2])
    state[0] = np.float32((np.float32(state[0]) * (k - np.float32(1.0)) + this_mean) / k)
    state[1] = np.float32((np.float32(state[1]) * (k - np.float32(1.0)) + this_std) / k)
    return ((x_tx - state[0]) / state[1]).astype(np.float32)


def ste_quantize(x: np.ndarray, value_limit: float = 1.0, quantize_level: float = 2) -> np.ndarray:
    """STEQuantize.forward, reference encoders.py:20-37."""
    xc = np.clip(x.astype(F32), F32(-value_limit), F32(value_limit))
    if quantize_level == 2:
        return np.sign(xc).astype(F32)
    q, rng = F32(quantize_level), F32(2.0 * value_limit)
    return (np.round((xc + F32(value_limit)) * ((q - F32(1.0)) / rng)) * rng / (q - F32(1.0)) - F32(value_limit)).astype(F32)


def enc_forward(u: np.ndarray, weights, p: np.ndarray, prefix: str = "enc", ste: bool = False, value_limit: float = 1.0,
                quantize_level: float = 2, dense: bool = False) -> np.ndarray:
    """reference encoders.py:351-377 -> codes ``(B, L, 3)``; ``ste`` = train_channel_mode 'block_norm_ste' (:118-120)."""
    codes, _, _ = power_constraint(enc_forward_unnormalised(u, weights, p, prefix, dense=dense))
    return ste_quantize(codes, value_limit, quantize_level) if ste else codes


# --------------------------------------------------------------------------- #
# decoder                                                                     #
# --------------------------------------------------------------------------- #
def dec_forward(received: np.ndarray, weights, p: np.ndarray, num_iteration: int = 6,
                num_iter_ft: int = 5, extrinsic: bool = True, prefix: str = "dec",
                trace: list | None = None, dense: bool = False) -> np.ndarray:
    """reference decoders.py:219-269 -> posteriors ``(B, L, 1)`` in (0,1).

    If ``trace`` is a list, the output of every ``dec{1,2}_outputs[idx]`` Linear
    (``x_plr`` BEFORE the extrinsic subtraction -- what a forward hook on the
    reference's Linear modules sees) is appended: 2*num_iteration arrays, used
    to localise a kernel bug to one stack.
    """
    r = received.astype(F32, copy=False)
    B, L, _ = r.shape
    r_sys = r[:, :, 0:1]
    r_par1 = r[:, :, 1:2]
    r_par2 = r[:, :, 2:3]
    r_sys_int = interleave(r_sys, p)                                  # decoders.py:222
    prior = np.zeros((B, L, num_iter_ft), dtype=F32)                  # decoders.py:227
    n_layer = count_layers(weights, "%s.dec1_cnns.0" % prefix)
    same_shape_conv1d_ = dense_same_shape_conv1d if dense else same_shape_conv1d      # decoders.py:173-176
    x_plr = None
    for idx in range(num_iteration):
        last = idx == num_iteration - 1
        x_in = np.concatenate([r_sys, r_par1, prior], axis=2)         # decoders.py:230 / 252
        h = same_shape_conv1d_(x_in, _stack(weights, "%s.dec1_cnns.%d" % (prefix, idx), n_layer))
        x_plr = linear(h, _get(weights, "%s.dec1_outputs.%d.module.weight" % (prefix, idx)),
                       _get(weights, "%s.dec1_outputs.%d.module.bias" % (prefix, idx)))
        if trace is not None:
            trace.append(x_plr.copy())
        if extrinsic:
            x_plr = x_plr - prior                                     # decoders.py:235-236 / 257-258
        x_plr_int = interleave(x_plr, p)                              # decoders.py:238 / 260
        x_in = np.concatenate([r_sys_int, r_par2, x_plr_int], axis=2)  # decoders.py:240 / 262
        h = same_shape_conv1d_(x_in, _stack(weights, "%s.dec2_cnns.%d" % (prefix, idx), n_layer))
        x_plr = linear(h, _get(weights, "%s.dec2_outputs.%d.module.weight" % (prefix, idx)),
                       _get(weights, "%s.dec2_outputs.%d.module.bias" % (prefix, idx)))
        if trace is not None:
            trace.append(x_plr.copy())
        if not last:
            if extrinsic:
                x_plr = x_plr - x_plr_int                             # decoders.py:246-247
            prior = deinterleave(x_plr, p)                            # decoders.py:249
    return sigmoid(deinterleave(x_plr, p))                            # decoders.py:267


# --------------------------------------------------------------------------- #
# system model + metrics                                                      #
# --------------------------------------------------------------------------- #
def channel_ae_forward(u, noise, weights, p, num_iteration=6):
    """reference channel_ae.py:38-73 (AWGN branch): returns ``(x_dec, codes)``."""
    codes = enc_forward(u, weights, p)
    received = (codes + noise.astype(F32)).astype(F32)
    return dec_forward(received, weights, p, num_iteration), codes


def errors_ber(y_true: np.ndarray, y_pred: np.ndarray) -> float:
    """reference utils.py:6-18: mean of ``round(y_true) != round(y_pred)``."""
    t = np.round(y_true.reshape(y_true.shape[0], -1))
    q = np.round(y_pred.reshape(y_pred.shape[0], -1))
    return float(np.mean(t != q))


def errors_bler(y_true: np.ndarray, y_pred: np.ndarray) -> float:
    """reference utils.py:49-66: fraction of blocks with at least one bit error."""
    t = np.round(y_true.reshape(y_true.shape[0], -1))
    q = np.round(y_pred.reshape(y_pred.shape[0], -1))
    return float(np.mean(np.any(t != q, axis=1)))


def snr_db2sigma(snr_db: float) -> float:
    """reference utils.py:69-70."""
    return 10 ** (-snr_db * 1.0 / 20)


# algorithmic work per codeword ---------------------
def decoder_flops_per_codeword(L=100, n_iter=6, n_ft=5, units=100, n_layer=5, k=5) -> int:
    macs = L * (2 * n_iter * ((2 + n_ft) * units * k + (n_layer - 1) * units * units * k)
                + (2 * n_iter - 1) * units * n_ft + units * 1)
    return 2 * macs


def encoder_flops_per_codeword(L=100, units=100, n_layer=2, k=5) -> int:
    macs = 3 * L * (1 * units * k + (n_layer - 1) * units * units * k) + 3 * L * units
    return 2 * macs


# --------------------------------------------------------------------------- #
# on-device channel                            #
# --------------------------------------------------------------------------- #
def philox4x32_10(counter: np.ndarray, seed: int) -> np.ndarray:
    """Philox4x32-10 (Salmon et al., SC'11; the generator behind curand / torch CUDA RNG) for 64-bit counters
    ``counter`` (uint64 array) and a 64-bit key: returns uint32 array of shape (len(counter), 4).  This is the
    published algorithm restated -- turboae_b200's ``tae_awgn_f32`` must reproduce it bit for bit."""
    c = np.asarray(counter, dtype=np.uint64)
    c0 = (c & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    c1 = (c >> np.uint64(32)).astype(np.uint64)
    c2 = np.zeros_like(c0)
    c3 = np.zeros_like(c0)
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & MASK
        k1 = (k1 + np.uint64(0xBB67AE85)) & MASK
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def awgn_noise(n: int, seed: int, offset: int = 0) -> np.ndarray:
    """The N(0,1) stream of ``tae_awgn_f32``: element i = Box-Muller of words (2*(i%4//2), +1) of Philox counter
    offset + i//4 -- cos branch for even i, sin branch for odd i."""
    n4 = (n + 3) // 4
    x = philox4x32_10(np.arange(offset, offset + n4, dtype=np.uint64), seed)
    u = (x.astype(np.float32) + F32(0.5)) * F32(2.3283064365386963e-10)
    z = np.empty((n4, 4), dtype=F32)
    for j in (0, 2):
        r = np.sqrt(F32(-2.0) * np.log(u[:, j])).astype(F32)
        ang = (F32(2.0) * u[:, j + 1]).astype(np.float64) * np.pi
        z[:, j] = r * np.cos(ang).astype(F32)
        z[:, j + 1] = r * np.sin(ang).astype(F32)
    return z.reshape(-1)[:n]


def awgn(codes: np.ndarray, sigma: float, seed: int, offset: int = 0) -> np.ndarray:
    """reference channel_ae.py:41-42: received = codes + fwd_noise, fwd_noise = sigma * N(0,1) (channels.py:31-35)."""
    z = awgn_noise(codes.size, seed, offset).reshape(codes.shape)
    return (codes.astype(F32) + F32(sigma) * z).astype(F32)


def error_counts(y_true: np.ndarray, y_pred: np.ndarray):
    """(bit errors, block errors): the numerators of reference utils.py:6-18 (errors_ber) and :49-66 (errors_bler)."""
    t = np.round(y_true.reshape(y_true.shape[0], -1))
    q = np.round(y_pred.reshape(y_pred.shape[0], -1))
    wrong = t != q
    return int(wrong.sum()), int(wrong.any(axis=1).sum())


# --------------------------------------------------------------------------- #
# DEC_LargeRNN                                 #
# --------------------------------------------------------------------------- #
def gru_direction(x: np.ndarray, w_ih, w_hh, b_ih, b_hh, reverse: bool = False) -> np.ndarray:
    """One direction of one torch.nn.GRU layer, batch_first (the reference calls torch.nn.GRU, decoders.py:43-52; this is
    PyTorch's documented cell):  r = s(W_ir x + b_ir + W_hr h + b_hr), z = s(W_iz x + b_iz + W_hz h + b_hz),
    n = tanh(W_in x + b_in + r * (W_hn h + b_hn)), h' = (1 - z) * n + z * h;  gate order in the weights is r, z, n."""
    B, L, _ = x.shape
    H = w_hh.shape[1]
    xp = (x.reshape(B * L, -1).astype(F32) @ w_ih.T.astype(F32) + b_ih.astype(F32)).reshape(B, L, 3 * H)
    h = np.zeros((B, H), dtype=F32)
    out = np.zeros((B, L, H), dtype=F32)
    steps = range(L - 1, -1, -1) if reverse else range(L)
    for t in steps:
        hp = h @ w_hh.T.astype(F32) + b_hh.astype(F32)
        r = sigmoid(xp[:, t, :H] + hp[:, :H])
        z = sigmoid(xp[:, t, H:2 * H] + hp[:, H:2 * H])
        n = np.tanh(xp[:, t, 2 * H:] + r * hp[:, 2 * H:]).astype(F32)
        h = ((F32(1.0) - z) * n + z * h).astype(F32)
        out[:, t, :] = h
    return out


def gru_stack(x: np.ndarray, weights, prefix: str, num_layers: int = 2) -> np.ndarray:
    """torch.nn.GRU(num_layers=2, bidirectional=True, batch_first=True) forward: (B, L, in) -> (B, L, 2H)."""
    h = x
    for layer in range(num_layers):
        outs = []
        for suffix, rev in (("", False), ("_reverse", True)):
            k = "l%d%s" % (layer, suffix)
            outs.append(gru_direction(h, _get(weights, "%s.module.weight_ih_%s" % (prefix, k)), _get(weights, "%s.module.weight_hh_%s" % (prefix, k)),
                                      _get(weights, "%s.module.bias_ih_%s" % (prefix, k)), _get(weights, "%s.module.bias_hh_%s" % (prefix, k)), rev))
        h = np.concatenate(outs, axis=2)
    return h


def dec_rnn_forward(received: np.ndarray, weights, p: np.ndarray, num_iteration: int = 6, num_iter_ft: int = 5,
                    extrinsic: bool = True, prefix: str = "dec") -> np.ndarray:
    """DEC_LargeRNN.forward, reference decoders.py:86-149 (dropout 0, dec_act 'linear'): the turbo schedule of dec_forward
    with bi-GRU stacks instead of conv stacks."""
    r = received.astype(F32, copy=False)
    B, L, _ = r.shape
    r_sys, r_par1, r_par2 = r[:, :, 0:1], r[:, :, 1:2], r[:, :, 2:3]
    r_sys_int = interleave(r_sys, p)
    prior = np.zeros((B, L, num_iter_ft), dtype=F32)
    x_plr = None
    for idx in range(num_iteration):
        last = idx == num_iteration - 1
        h = gru_stack(np.concatenate([r_sys, r_par1, prior], axis=2), weights, "%s.dec1_rnns.%d" % (prefix, idx))
        x_plr = linear(h, _get(weights, "%s.dec1_outputs.%d.module.weight" % (prefix, idx)), _get(weights, "%s.dec1_outputs.%d.module.bias" % (prefix, idx)))
        if extrinsic:
            x_plr = x_plr - prior
        x_plr_int = interleave(x_plr, p)
        h = gru_stack(np.concatenate([r_sys_int, r_par2, x_plr_int], axis=2), weights, "%s.dec2_rnns.%d" % (prefix, idx))
        x_plr = linear(h, _get(weights, "%s.dec2_outputs.%d.module.weight" % (prefix, idx)), _get(weights, "%s.dec2_outputs.%d.module.bias" % (prefix, idx)))
        if not last:
            if extrinsic:
                x_plr = x_plr - x_plr_int
            prior = deinterleave(x_plr, p)
    return sigmoid(deinterleave(x_plr, p))
