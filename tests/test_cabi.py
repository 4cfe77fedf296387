"""CPU-side checks: the C-ABI library loads and exports every symbol include/turboae_b200.h declares, and the
Python mirror of the reference interface (module tree, checkpoint keys, interleaver host logic, error
behaviour without a GPU) is right.  No compute call is made here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import turboae_b200 as T
from turboae_b200 import _lib
from helpers import ROOT, Codec, load_npz, make_args
from oracle import turboae_oracle as O


def _header_functions():
    src = open(os.path.join(ROOT, "include", "turboae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tae_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), "libturboae_b200.so does not export %s" % n
    assert sorted(_lib.PUBLIC_SYMBOLS) == names          # the ctypes table covers the whole header, nothing more
    assert lib.tae_version() >= 100


def test_host_side_size_queries_need_no_gpu():
    lib = _lib.load()
    cfg = _lib.TaeDecConfig(100, 6, 5, 5, 100, 5, 1)
    assert lib.tae_dec_param_count(ctypes.byref(cfg)) == 2_453_656
    assert lib.tae_dec_packed_bytes(ctypes.byref(cfg)) > 0
    enc = _lib.TaeEncConfig(100, 2, 100, 5)
    assert lib.tae_enc_param_count(ctypes.byref(enc)) == 152_403                # row a4
    enc5 = _lib.TaeEncConfig(100, 5, 100, 5)
    assert lib.tae_enc_param_count(ctypes.byref(enc5)) == 603_303
    bad = _lib.TaeDecConfig(100, 6, 5, 5, 100, 4, 1)                            # even kernel size
    assert lib.tae_dec_param_count(ctypes.byref(bad)) == 0
    assert b"kernel_size" in lib.tae_last_error()


def test_split_operand_path_size_queries_and_validation_need_no_gpu():
    """TAE_PRECISION_F16X3 (tae_x3.cu): image sizes are arithmetic (2 + 22 (layers - 1) slots of 10 752 bytes per stack),
    unsupported shapes are refused before any CUDA call."""
    lib = _lib.load()
    cfg = _lib.TaeDecConfig(100, 6, 5, 5, 100, 5, 1)
    assert lib.tae_dec_packed_bytes_x3(ctypes.byref(cfg)) == 12 * (2 + 22 * 4) * 10752
    assert lib.tae_dec_workspace_bytes(ctypes.byref(cfg), 50000, _lib.PRECISION_F16X3) == 256
    enc = _lib.TaeEncConfig(100, 2, 100, 5)
    assert lib.tae_enc_packed_bytes_x3(ctypes.byref(enc)) == 3 * (2 + 22) * 10752
    assert lib.tae_enc_packed_bytes_x3(ctypes.byref(_lib.TaeEncConfig(510, 5, 104, 5))) == 3 * (2 + 22 * 4) * 10752
    for bad, word in ((_lib.TaeEncConfig(600, 2, 100, 5), b"block_len"), (_lib.TaeEncConfig(100, 2, 128, 5), b"num_unit"),
                      (_lib.TaeEncConfig(100, 2, 100, 3), b"kernel_size"), (_lib.TaeEncConfig(100, 9, 100, 5), b"num_layer")):
        assert lib.tae_enc_packed_bytes_x3(ctypes.byref(bad)) == 0 and word in lib.tae_last_error(), word
    assert lib.tae_dec_packed_bytes_x3(ctypes.byref(_lib.TaeDecConfig(100, 6, 6, 5, 100, 5, 1))) == 0      # num_iter_ft > 5
    assert lib.tae_enc_forward_f16x3(ctypes.byref(enc), None, None, None, None, None, None, None, 3, None, 0, None) == -1
    assert lib.tae_enc_forward_f16x3(ctypes.byref(enc), None, None, None, None, None, None, None, 0, None, 0, None) == 0   # empty batch
    assert lib.tae_dec_forward(ctypes.byref(cfg), None, None, None, None, None, None, None, 3, _lib.PRECISION_F16X3, None, 0, None) == -1
    # the module resolves 'auto' without a device: the split-operand kernel where it covers the shape, else the fp32 kernels
    e = T.ENC_interCNN(make_args(no_cuda=True), O.make_perm(100, 0))
    assert e.precision == "auto" and e.resolved_precision(100) == "f16x3" and e.resolved_precision(510) == "f16x3" and e.resolved_precision(600) == "fp32"
    e.precision = "nonsense"
    with pytest.raises(_lib.TaeError):
        e.resolved_precision(100)


def test_error_codes_on_bad_arguments():
    lib = _lib.load()
    assert lib.tae_interleave_f32(None, None, None, 4, 0, 1, None) == -1        # TAE_EINVAL, before any CUDA call
    assert b"bad shape" in lib.tae_last_error()
    assert lib.tae_interleave_f32(None, None, None, 0, 10, 1, None) == 0        # empty batch is a no-op
    cfg = _lib.TaeDecConfig(100, 6, 5, 5, 100, 5, 1)
    assert lib.tae_dec_forward(ctypes.byref(cfg), None, None, None, None, None, None, None, 3, 0, None, 0, None) == -1
    with pytest.raises(_lib.TaeError):
        _lib.check(-2)


@pytest.mark.parametrize("cfg,n_enc_layer", [("c1", 2), ("c3", 5)])
def test_module_tree_matches_shipped_checkpoint_keys(cfg, n_enc_layer):
    w = load_npz("weights_%s.npz" % cfg)
    args = make_args(enc_num_layer=n_enc_layer, no_cuda=True)
    m = Codec(args, O.make_perm(100, 0))
    sd = m.state_dict()
    assert sorted(sd.keys()) == sorted(w.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == w[k].shape, k
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    # canonical flat order == library's parameter count
    lib = _lib.load()
    assert sum(p.numel() for p in m.dec.ordered_parameters()) == lib.tae_dec_param_count(ctypes.byref(m.dec.config(100)))
    assert sum(p.numel() for p in m.enc.ordered_parameters()) == lib.tae_enc_param_count(ctypes.byref(m.enc.config(100)))
    assert len({id(p) for p in m.dec.ordered_parameters()}) == len(list(m.dec.parameters()))
    # without set_parallel the keys lose the '.module.' level, like the reference (encoders.py:343-349)
    plain = T.DEC_LargeCNN(args, O.make_perm(100, 0))
    assert "dec1_cnns.0.cnns.0.weight" in plain.state_dict()


def test_rnn_decoder_module_tree_matches_reference_keys():
    g = load_npz("rnn_h32_i2_l40_b5.npz")
    B, L, H, n_iter = g["cfg"].tolist()
    m = T.DEC_LargeRNN(make_args(no_cuda=True, num_iteration=n_iter, dec_num_unit=H, block_len=L), g["p"])
    m.set_parallel()
    sd = m.state_dict()
    ref_keys = sorted(k[4:] for k in g if k.startswith("dec."))
    assert sorted(sd.keys()) == ref_keys
    for k, v in sd.items():
        assert tuple(v.shape) == g["dec." + k].shape, k
    with torch.no_grad(), pytest.raises(_lib.TaeError):
        m(torch.zeros(B, L, 3))


def test_interleaver_host_logic():
    p = O.make_perm(100, 0)
    args = make_args(no_cuda=True)
    il, dl = T.Interleaver(args, p), T.DeInterleaver(args, p)
    assert il.p_array.dtype == torch.int64 and il.p_array.tolist() == p.tolist()
    assert dl.reverse_p_array.tolist() == O.inverse_perm(p).tolist()
    before = il._p
    il.set_parray(p.copy())                       # Channel_AE.forward re-sets the same permutation every call
    assert il._p is before
    il.set_parray(O.make_perm(100, 7))
    assert il._p is not before
    with pytest.raises(_lib.TaeError):
        il.set_parray(np.array([0, 0, 1]))


def test_no_cpu_fallback():
    args = make_args(no_cuda=True)
    p = O.make_perm(100, 0)
    enc, dec = T.ENC_interCNN(args, p), T.DEC_LargeCNN(args, p)
    with torch.no_grad():
        with pytest.raises(_lib.TaeError):
            dec(torch.zeros(2, 100, 3))
        with pytest.raises(_lib.TaeError):
            enc(torch.zeros(2, 100, 1))
        with pytest.raises(_lib.TaeError):
            T.Interleaver(args, p)(torch.zeros(2, 100, 1))


def test_unsupported_configurations_raise():
    p = O.make_perm(100, 0)
    # the dense decoder variant (decoders.py:173-176) is built from DenseSameShapeConv1d with the reference's parameter shapes ...
    d = T.DEC_LargeCNN(make_args(encoder="TurboAE_rate3_cnn_dense", dec_num_unit=20, dec_num_layer=3), p)
    assert d.dense and tuple(d.dec1_cnns[0].cnns[2].weight.shape) == (20, 7 + 2 * 20, 5)
    # ... and so is ENC_interCNN (encoders.py:322-330; main.py:34 routes both -encoder values to this class)
    e = T.ENC_interCNN(make_args(encoder="TurboAE_rate3_cnn_dense", enc_num_unit=20, enc_num_layer=3), p)
    assert e.dense and e.train_precision == "fp32" and tuple(e.enc_cnn_3.cnns[2].weight.shape) == (20, 1 + 2 * 20, 5)
