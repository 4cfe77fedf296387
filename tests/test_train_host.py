"""CPU tests (`-m "not gpu"`) of the host-side logic behind the tensor-core training and GRU paths: weight-gradient job
construction, flat-gradient buffer reuse, the C ABI's argument validation and geometry helpers (no kernel is launched)."""
import ctypes as C
from types import SimpleNamespace

import pytest
import torch

from helpers import make_args
from turboae_b200 import _lib, train_tc


def _jobs(units, n_layer, cin0, fouts, groups, splits=None):
    n_stacks = len(fouts)
    img = lambda n: torch.zeros(max(n, 1), dtype=torch.uint8)
    cb = _lib.IMG_CHUNK_BYTES
    stash_y = img(n_stacks * n_layer * groups * 13 * cb // 1024)       # only the base addresses matter here
    stash_g, stash_x, stash_d = img(64), img(64), img(64)
    offsets, off = [], 0
    for st in range(n_stacks):
        layers = []
        for j in range(n_layer):
            cin = cin0 if j == 0 else units
            layers.append((off, off + units * cin * 5))
            off += units * cin * 5 + units
        offsets.append((layers, off))
        off += fouts[st] * units + fouts[st]
    gflat = torch.zeros(off)
    jobs = train_tc.wgrad_jobs(n_layer, units, cin0, fouts, groups, stash_y, stash_x, stash_g, stash_d, gflat, offsets, splits=splits)
    return jobs, offsets, gflat


@pytest.mark.parametrize("units,n_layer,cin0,fouts,groups", [(100, 5, 7, [5] * 11 + [1], 200), (64, 3, 5, [3, 3], 7), (60, 2, 1, [1, 1, 1], 40),
                                                              (30, 4, 7, [5, 1], 1), (8, 2, 4, [2], 3), (96, 5, 7, [5, 1], 1000)])
def test_wgrad_jobs_cover_every_parameter_exactly_once(units, n_layer, cin0, fouts, groups):
    jobs, offsets, gflat = _jobs(units, n_layer, cin0, fouts, groups)
    base = gflat.data_ptr()
    seen = {}                      # (flat index of the gradient element of output channel 0, group) -> count
    bias_jobs = {}
    last_cost = None
    last_job = None
    for j in jobs:
        # constraints of tae_wgrad_bf16 (include/turboae_b200.h)
        assert j.n_cols % 16 == 0 and 8 * j.b_nc <= j.n_cols <= 8 * (j.b_nc + 1) and j.taps * j.n_cols <= 512
        assert 1 <= j.b_nc <= 13 and 0 <= j.b_c0 and j.b_c0 + j.b_nc <= j.b_chunks and 1 <= j.taps <= 5
        assert 0 <= 2 - j.taps // 2 + j.tap_shift and 2 - j.taps // 2 + j.tap_shift + j.taps - 1 <= 4      # row offsets within the halo
        assert 0 <= j.n_valid <= 8 * j.b_nc and 1 <= j.m_valid <= 104 and 0 <= j.g0 < j.g1 <= groups
        if j.bias_grad:
            assert 8 * j.b_nc < j.n_cols                                  # the spare column that carries the bias gradient exists
            for g in range(j.g0, j.g1):
                bias_jobs[(j.bias_grad, g)] = bias_jobs.get((j.bias_grad, g), 0) + 1
        goff = (j.grad - base) // 4
        for n in range(j.n_valid):
            for t in range(j.taps):
                for g in range(j.g0, j.g1):
                    k = (goff + (j.n0 + n) * j.s_n + t * j.s_t, g)
                    seen[k] = seen.get(k, 0) + 1
        cost = train_tc._job_cost_us(j.b_chunks, j.n_cols, j.taps) * (j.g1 - j.g0)
        # longest first; the jobs of one layer (same A image, same group range) follow their longest member directly
        follower = last_job is not None and (j.a_img, j.g0, j.g1) == (last_job.a_img, last_job.g0, last_job.g1)
        assert last_cost is None or cost <= last_cost + 1e-9
        if not follower:
            last_cost = cost
        last_job = j
    assert all(v == 1 for v in seen.values()) and all(v == 1 for v in bias_jobs.values())
    for st, (layers, lin_w_off) in enumerate(offsets):
        for jl, (w_off, b_off) in enumerate(layers):
            cin = cin0 if jl == 0 else units
            for e in range(cin * 5):                                      # every (input channel, tap) of output channel 0
                for g in (0, groups - 1):
                    assert seen.get((w_off + e, g)) == 1, (st, jl, e, g)
            assert bias_jobs.get((base + 4 * b_off, 0)) == 1 and bias_jobs.get((base + 4 * b_off, groups - 1)) == 1
        for f in range(fouts[st]):
            assert seen.get((lin_w_off + f * units, 0)) == 1


def test_wgrad_jobs_forced_splits_partition_the_groups():
    jobs, _, _ = _jobs(100, 5, 7, [5, 1], 10, splits=3)
    ranges = sorted({(j.g0, j.g1) for j in jobs})
    assert ranges == [(0, 3), (3, 6), (6, 10)]


def test_flat_gradient_buffer_is_reused_unless_a_grad_still_aliases_it():
    buf = SimpleNamespace(gflat=None, jobs="cached")
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5))]
    flat = torch.zeros(17)
    g1, fresh = train_tc._flat_grad(buf, flat, params)
    assert fresh and buf.jobs is None and g1.numel() == 17
    buf.jobs = "cached"
    g1.fill_(3.0)
    g2, fresh = train_tc._flat_grad(buf, flat, params)                    # zero_grad(set_to_none=True) case: same buffer, zeroed
    assert not fresh and g2.data_ptr() == g1.data_ptr() and float(g2.abs().sum()) == 0.0 and buf.jobs == "cached"
    params[0].grad = g2[:12].view(3, 4)                                    # autograd adopted a view as .grad: accumulation in progress
    g3, fresh = train_tc._flat_grad(buf, flat, params)
    assert fresh and g3.data_ptr() != g2.data_ptr() and buf.jobs is None   # a new buffer: the live gradient is not clobbered


def test_supported_configurations():
    assert train_tc.supported(make_args(), "dec") and train_tc.supported(make_args(), "enc")
    assert not train_tc.supported(make_args(dec_kernel_size=3), "dec")
    assert not train_tc.supported(make_args(dec_num_unit=128), "dec")
    assert not train_tc.supported(make_args(block_len=1000), "dec") and not train_tc.supported(make_args(block_len=1000), "enc")
    assert not train_tc.supported(make_args(enc_num_layer=1), "enc")


def test_cabi_geometry_and_validation_without_a_gpu():
    lib = _lib.load()
    assert lib.tae_train_groups(100, 1000) == 200 and lib.tae_train_groups(100, 1) == 1 and lib.tae_train_groups(100, 0) == 0
    assert lib.tae_train_groups(37, 29) == 3                               # 13 codewords of 37 + 2 rows per 512-row group
    assert lib.tae_train_groups(512, 3) == 3 and lib.tae_train_groups(513, 3) == 0
    assert [lib.tae_gru_rows_per_block(b) for b in (1, 2368, 4736, 5000, 9472, 18944, 10 ** 6)] == [32, 32, 32, 64, 64, 128, 128]
    assert lib.tae_gru_tile_bytes(300, 50, 26, 96) == 4 * 50 * 26 * 96 * 16   # 4 blocks: an even number (whole CTA pairs)
    assert lib.tae_gru_tile_bytes(300, 50, 26, 48) == 0                     # rows per block: 32, 64, 96 or 128
    # weight image of one layer-direction: (input k-steps + 7) x 3584 + input k-steps x 1792 + 7 x 1792 bytes per CTA
    assert lib.tae_gru_packed_bytes(100, 7, 7) == 2 * ((1 + 7) * 3584 + 1 * 1792 + 7 * 1792)
    assert lib.tae_gru_packed_bytes(100, 200, 100) == 2 * ((14 + 7) * 3584 + 14 * 1792 + 7 * 1792)
    assert lib.tae_gru_packed_bytes(33, 7, 7) == 0 and b"hidden size" in lib.tae_last_error()
    assert lib.tae_gru_packed_bytes(100, 300, 100) == 0                     # 39 input chunks > 26
    # malformed weight-gradient jobs are rejected before anything touches the device
    good = dict(a_img=8, b_img=8, grad=8, bias_grad=None, b_chunks=13, b_c0=0, b_nc=8, taps=5, n_cols=64, m_valid=100, n_valid=64, n0=0,
                s_m=500, s_n=5, s_t=1, g0=0, g1=4, tap_shift=0)
    for bad in (dict(n_cols=72), dict(b_nc=9), dict(b_c0=6), dict(taps=6), dict(tap_shift=1), dict(taps=3, tap_shift=-2), dict(b_nc=14, n_cols=112), dict(n_cols=112), dict(m_valid=105), dict(g1=-1), dict(grad=None)):
        job = _lib.TaeWgradJob(**{**good, **bad})
        arr = (_lib.TaeWgradJob * 1)(job)
        assert lib.tae_wgrad_bf16(arr, 1, None, C.c_void_p(8), 4096, None) == -1, bad
        assert b"malformed" in lib.tae_last_error()
    assert lib.tae_wgrad_bf16(None, 0, None, None, 0, None) == 0             # an empty job list is a no-op


def test_local_power_normalisation_without_process_group():
    from turboae_b200 import shard
    x = torch.randn(4, 10, 3, dtype=torch.float32, requires_grad=True)
    y = shard.PowerNorm.apply(x, shard.LOCAL)
    ref = (x - x.mean()) / x.std()
    assert torch.allclose(y, ref, atol=1e-6)
    g = torch.randn_like(x)
    y.backward(g)
    xr = x.detach().clone().requires_grad_(True)
    ((xr - xr.mean()) / xr.std()).backward(g)
    assert torch.allclose(x.grad, xr.grad, atol=1e-6)


def test_flat_cache_follows_writes_through_data():
    """`p.data.copy_()` (the reference's Lookahead, optimizers.py:29) bumps no version counter; the cache is told
    by the optimizer post-step hook / launch.py's Lookahead wrapper through invalidate_all()."""
    import torch
    import turboae_b200 as T
    from turboae_b200._flat import FlatCache
    p = [torch.nn.Parameter(torch.ones(4)), torch.nn.Parameter(torch.zeros(3))]
    c = FlatCache()
    f0 = c.get(p)
    c.derived["bf16"] = "image"
    assert c.get(p) is f0 and c.derived                     # unchanged parameters: cached
    p[0].data.copy_(torch.full((4,), 2.0))                  # invisible to p._version
    T.invalidate_all()
    f1 = c.get(p)
    # (the flat image is persistent: refreshed in place by one multi-tensor copy, same address; the derived images are dropped)
    assert f1.data_ptr() == f0.data_ptr() and float(f1[0]) == 2.0 and not c.derived
    # an optimizer step invalidates on its own (post-step hook registered at import)
    opt = torch.optim.SGD(p, lr=1.0)
    f2 = c.get(p)
    c.derived["bf16"] = "image"
    p[1].grad = torch.ones(3)
    opt.step()
    f3 = c.get(p)
    assert f3.data_ptr() == f2.data_ptr() and float(f3[-1]) == -1.0 and not c.derived
    # another parameter set (sizes / device): a new image
    q = [torch.nn.Parameter(torch.ones(5)), torch.nn.Parameter(torch.zeros(3))]
    f4 = c.get(q)
    assert f4.numel() == 8 and float(f4.sum()) == 5.0 and c.get(p).numel() == 7


def test_stash_is_not_shared_between_two_live_forwards():
    """two forwards of one module before backward must not share the activation stash."""
    import gc
    import torch
    from turboae_b200 import train_tc, _lib

    class Mod:
        pass
    m = Mod()
    b1 = train_tc._buffers(m, 2, 2, 1, 1, 10, 5, "cpu")
    tok1, gen1 = b1.claim()
    assert b1.busy()
    b2 = train_tc._buffers(m, 2, 2, 1, 1, 10, 5, "cpu")      # second forward while the first graph is alive
    assert b2 is not b1
    tok1.done = True                                        # backward of the first graph ran
    assert train_tc._buffers(m, 2, 2, 1, 1, 10, 5, "cpu") is b1
    tok3, gen3 = b1.claim()
    del tok3                                                # the graph was dropped without backward
    gc.collect()
    assert not b1.busy()

    class Ctx:
        pass
    ctx = Ctx()
    ctx.buf, ctx.gen = b1, gen1                             # stale generation (retain_graph + a later forward)
    with pytest.raises(_lib.TaeError):
        train_tc._check_stash(ctx)


def test_canonical_parameter_order_is_cached_and_follows_replaced_parameters():
    """_flat.OrderedParameters: the module tree is walked once; the list is dropped whenever Parameter objects can have been
    replaced (set_parallel, load_state_dict at any level incl. assign=True, .to() / dtype casts, deepcopy keeps its own)."""
    import copy
    import pickle
    import torch
    import turboae_b200 as T
    from helpers import make_args
    from oracle import turboae_oracle as O
    p = O.make_perm(100, 0)
    d, e = T.DEC_LargeCNN(make_args(), p), T.ENC_interCNN(make_args(), p)
    fresh = lambda m: [id(x) for x in m._walk_ordered_parameters()]
    l1 = d.ordered_parameters()
    assert d.ordered_parameters() is l1 and len(l1) == 144 and len(e.ordered_parameters()) == 18
    d.set_parallel()
    assert d.ordered_parameters() is not l1 and [id(x) for x in d.ordered_parameters()] == fresh(d)
    l2 = d.ordered_parameters()
    d.load_state_dict(d.state_dict())
    assert d.ordered_parameters() is not l2
    l3 = d.ordered_parameters()
    d.load_state_dict(d.state_dict(), assign=True)                      # replaces the Parameter objects
    assert [id(x) for x in d.ordered_parameters()] == fresh(d) and any(x is not y for x, y in zip(l3, d.ordered_parameters()))
    model = torch.nn.Module()                                           # Channel_AE-like parent: the load recurses into enc / dec
    model.enc, model.dec = e, d
    le = e.ordered_parameters()
    model.load_state_dict(model.state_dict(), assign=True)
    assert e.ordered_parameters() is not le and [id(x) for x in e.ordered_parameters()] == fresh(e)
    d2 = copy.deepcopy(d)
    assert [id(x) for x in d2.ordered_parameters()] == fresh(d2)
    assert all(x is not y for x, y in zip(d2.ordered_parameters(), d.ordered_parameters()))
    d.double()
    assert [id(x) for x in d.ordered_parameters()] == fresh(d)
    pickle.dumps(d)                                                     # the load hook is a module-level function


def test_backward_split_and_group_range_job_lists_partition_the_batch():
    """The two-launch backward (train_tc.backward_split) and the weight-gradient job lists of its two group ranges: together they
    cover every (parameter slab, group) exactly once, like the single list."""
    import torch
    from turboae_b200 import train_tc
    # one full wave + a partly filled one on 74 CTA pairs -> split after the full wave; otherwise one launch
    assert train_tc.backward_split(100, 148) == 74 and train_tc.backward_split(75, 148) == 74
    assert train_tc.backward_split(150, 148) == 148 and train_tc.backward_split(223, 148) == 222
    for units in (1, 50, 74, 126, 148, 200, 300, 5000):
        assert train_tc.backward_split(units, 148) == 0, units
    assert train_tc.backward_split(100, 0) == 0
    units, n_layer, cin0, fouts, groups = 100, 5, 7, [5, 1], 200
    cb = 8256
    stash_y = torch.zeros(1, dtype=torch.uint8)
    offsets, off = [], 0
    for st in range(2):
        layers = []
        for j in range(n_layer):
            cin = cin0 if j == 0 else units
            layers.append((off, off + units * cin * 5))
            off += units * cin * 5 + units
        offsets.append((layers, off))
        off += fouts[st] * units + fouts[st]
    gflat = torch.zeros(off)
    mk = lambda rng: train_tc.wgrad_jobs(n_layer, units, cin0, fouts, groups, stash_y, stash_y, stash_y, stash_y, gflat, offsets, group_range=rng)
    whole, head, tail = mk(None), mk((0, 148)), mk((148, 200))
    cover = lambda jobs: sorted((j.grad, j.n0, j.b_c0, j.taps, g) for j in jobs for g in range(j.g0, j.g1))
    assert cover(head + tail) == cover(whole) and len(set(cover(whole))) == len(cover(whole))
    assert all(j.g1 <= 148 for j in head) and all(j.g0 >= 148 for j in tail) and head and tail
    assert mk((5, 5)) == []
    with pytest.raises(ValueError):
        mk((10, 201))
