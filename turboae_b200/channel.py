"""On-device channel and metrics around the hot path.

``awgn`` replaces ``generate_noise`` + the add in ``Channel_AE.forward`` (reference channels.py:21-35,
channel_ae.py:41-42) and ``error_counts`` the numerators of ``errors_ber`` / ``errors_bler`` (reference utils.py:6-18,
49-66); ``ber_sweep`` is the SNR loop of ``trainer.test`` (reference trainer.py:157-178, 215-232) with bits, noise,
encode, decode and counting all on the device -- at > 2 M codewords/s the CPU randn and the 60 MB upload per batch of
the reference loop would otherwise dominate.  The noise stream (Philox4x32-10 + Box-Muller) is reproducible from
(seed, offset); the reference is unseeded, so it is compatible only with itself and with oracle/turboae_oracle.py."""
from __future__ import annotations

import torch

from . import _lib


def snr_db2sigma(snr_db: float) -> float:
    """reference utils.py:69-70."""
    return 10 ** (-snr_db * 1.0 / 20)


def awgn(codes: torch.Tensor, sigma: float, seed: int, offset: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
    """received = codes + sigma * N(0,1); element i of the flattened tensor uses Philox counter offset + i // 4."""
    _lib.require_cuda(codes, "awgn input")
    if codes.dtype != torch.float32 or not codes.is_contiguous():
        raise _lib.TaeError("awgn expects a contiguous float32 tensor")
    out = torch.empty_like(codes) if out is None else out
    with torch.cuda.device(codes.device):
        _lib.check(_lib.load().tae_awgn_f32(_lib.ptr(codes), _lib.ptr(out), codes.numel(), float(sigma), int(seed) & (2 ** 64 - 1),
                                            int(offset), _lib.stream_ptr(codes.device)))
    return out


class DeviceNoise:
    """``generate_noise`` of the reference (channels.py:7-35) for the AWGN channel, drawn on the device instead of the CPU:
    the reference's trainer draws ``torch.randn`` for a whole batch on the host and uploads it every step (trainer.py:53-62,
    167-171), which at this package's rates is most of a step.  Same distributions: test time ``sigma * N(0, 1)`` with
    ``sigma = 10^(-snr/20)``; training (``test_sigma == 'default'``) a per-element sigma uniform between the sigmas of
    ``snr_low`` and ``snr_high`` times ``N(0, 1)`` (channels.py:21-24).  The normal draws are this package's Philox stream
    (``tae_awgn_f32``; key = ``seed``, the counter advances with every call), so the noise is NOT the sequence torch's CPU generator
    would have produced -- the reference is unseeded, and a seeded comparison against its own classes needs the host channel.
    Other channels (bec, bsc, t-dist, radar, ge*) fall through to the reference's function."""

    def __init__(self, reference_generate_noise, device, seed: int = 0):
        self.reference = reference_generate_noise
        self.device = torch.device(device)
        self.seed = int(seed)
        self.offset = 0

    def normal(self, shape) -> torch.Tensor:
        z = torch.zeros(tuple(shape), dtype=torch.float32, device=self.device)
        out = awgn(z, 1.0, self.seed, self.offset, out=z)
        self.offset += (z.numel() + 3) // 4
        return out

    def __call__(self, noise_shape, args, test_sigma="default", snr_low=0.0, snr_high=0.0, mode="encoder"):
        if getattr(args, "channel", "awgn") != "awgn":
            return self.reference(noise_shape, args, test_sigma=test_sigma, snr_low=snr_low, snr_high=snr_high, mode=mode)
        noise = self.normal(noise_shape)
        if isinstance(test_sigma, str):                                    # 'default': the training mixture (channels.py:21-24)
            lo, hi = snr_db2sigma(snr_low), snr_db2sigma(snr_high)
            if lo == hi:
                return noise.mul_(hi)
            return noise.mul_(torch.rand(tuple(noise_shape), device=self.device).mul_(lo - hi).add_(hi))
        return noise.mul_(snr_db2sigma(test_sigma))                        # channels.py:30-35


def error_counts(y_true: torch.Tensor, y_pred: torch.Tensor, counts: torch.Tensor | None = None) -> torch.Tensor:
    """Adds (bit errors, block errors) of this batch to the device int64[2] tensor ``counts`` (created if None)."""
    _lib.require_cuda(y_pred, "error_counts input")
    if y_true.shape != y_pred.shape or y_true.dtype != torch.float32 or y_pred.dtype != torch.float32:
        raise _lib.TaeError("error_counts expects two float32 tensors of the same shape")
    B = y_true.shape[0]
    L = y_true.numel() // max(B, 1)
    if counts is None:
        counts = torch.zeros(2, dtype=torch.int64, device=y_pred.device)
    with torch.cuda.device(y_pred.device):
        _lib.check(_lib.load().tae_error_count_f32(_lib.ptr(y_true.contiguous()), _lib.ptr(y_pred.contiguous()), B, L,
                                                   _lib.ptr(counts), _lib.stream_ptr(y_pred.device)))
    return counts


@torch.no_grad()
def ber_sweep(enc, dec, snrs, num_block: int, batch_size: int, block_len: int = 100, seed: int = 0, device=None):
    """The test loop of the reference (trainer.py:157-178): for every SNR, num_block // batch_size batches of random bits
    through encode -> AWGN -> decode; returns (ber list, bler list, raw counts).  Everything stays on the device; one
    device->host read per SNR point."""
    device = device or next(dec.parameters()).device
    n_batch = num_block // batch_size
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    bers, blers, raw = [], [], []
    offset = 0
    for snr in snrs:
        sigma = snr_db2sigma(snr)
        counts = torch.zeros(2, dtype=torch.int64, device=device)
        for _ in range(n_batch):
            u = torch.randint(0, 2, (batch_size, block_len, 1), device=device, generator=gen).float()    # trainer.py:167
            codes = enc(u)
            received = awgn(codes, sigma, seed, offset)                                                   # trainer.py:169, channel_ae.py:41
            offset += (codes.numel() + 3) // 4
            error_counts(u, dec(received), counts)                                                        # trainer.py:176-177
        c = counts.tolist()
        raw.append(c)
        bers.append(c[0] / float(n_batch * batch_size * block_len))
        blers.append(c[1] / float(n_batch * batch_size))
    return bers, blers, raw
