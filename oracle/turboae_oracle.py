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
def enc_forward_unnormalised(u: np.ndarray, weights, p: np.ndarray, prefix: str = "enc") -> np.ndarray:
    """reference encoders.py:362-373: the three branches and the concat.

    ``u`` is ``(B, L, 1)`` in {0,1}; returns ``x_tx`` ``(B, L, 3)``.
    """
    x = (F32(2.0) * u.astype(F32) - F32(1.0))
    outs = []
    for i, inp in ((1, x), (2, x), (3, interleave(x, p))):
        n = count_layers(weights, "%s.enc_cnn_%d" % (prefix, i))
        h = same_shape_conv1d(inp, _stack(weights, "%s.enc_cnn_%d" % (prefix, i), n))
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


def enc_forward(u: np.ndarray, weights, p: np.ndarray, prefix: str = "enc") -> np.ndarray:
    """reference encoders.py:351-377 -> codes ``(B, L, 3)``."""
    codes, _, _ = power_constraint(enc_forward_unnormalised(u, weights, p, prefix))
    return codes


# --------------------------------------------------------------------------- #
# decoder                                                                     #
# --------------------------------------------------------------------------- #
def dec_forward(received: np.ndarray, weights, p: np.ndarray, num_iteration: int = 6,
                num_iter_ft: int = 5, extrinsic: bool = True, prefix: str = "dec",
                trace: list | None = None) -> np.ndarray:
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
    x_plr = None
    for idx in range(num_iteration):
        last = idx == num_iteration - 1
        x_in = np.concatenate([r_sys, r_par1, prior], axis=2)         # decoders.py:230 / 252
        h = same_shape_conv1d(x_in, _stack(weights, "%s.dec1_cnns.%d" % (prefix, idx), n_layer))
        x_plr = linear(h, _get(weights, "%s.dec1_outputs.%d.module.weight" % (prefix, idx)),
                       _get(weights, "%s.dec1_outputs.%d.module.bias" % (prefix, idx)))
        if trace is not None:
            trace.append(x_plr.copy())
        if extrinsic:
            x_plr = x_plr - prior                                     # decoders.py:235-236 / 257-258
        x_plr_int = interleave(x_plr, p)                              # decoders.py:238 / 260
        x_in = np.concatenate([r_sys_int, r_par2, x_plr_int], axis=2)  # decoders.py:240 / 262
        h = same_shape_conv1d(x_in, _stack(weights, "%s.dec2_cnns.%d" % (prefix, idx), n_layer))
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
