"""GPU suite (-m gpu): the CUDA path, called through the C ABI via the host-side mirror, against the oracle
on the same seeded inputs, against the golden fixtures minted from the reference, and -- at BASELINE's
full sizes -- through size-independent properties.

Tolerances (north star): cell indices, corner weights and the Splat-max grid are BIT-EXACT; everything
that sums floats is within rel 1e-5 (absolute floor 1e-5 * max|ref|).
"""
import ctypes

import numpy as np
import pytest
import torch

import cloud_transformers_b200 as ctb
from cloud_transformers_b200 import _lib, functional as CF
from oracle import ct_numpy as O
from tests.util import GOLDEN_CASES, assert_close, make_inputs

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
MODES = ["atomic", "tile", "deterministic"]


@pytest.fixture(autouse=True)
def _mode_reset():
    yield
    ctb.config.mode = "auto"
    ctb.config.fused = True
    ctb.config.use_plan = True


def t(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def n(x):
    return x.detach().cpu().numpy()


def run_block(keys, feat, pad, conv, go, gz, W, H, dim, fused=True, mode="auto", reduce="max"):
    """positions -> Splat -> (given conv grid) -> Slice, forward and both backwards, like an MHCT block."""
    ctb.config.mode = mode
    ctb.config.fused = fused
    dp = ctb.DifferentiablePositions(tensor_size=W, heads=H, dim=dim).to(DEV)
    sp = ctb.Splat(tensor_size=W, heads=H, dim=dim, reduce=reduce).to(DEV)
    sl = ctb.Slice(tensor_size=W, heads=H, dim=dim).to(DEV)
    k = t(keys).requires_grad_(True)
    f = t(feat).requires_grad_(True)
    p = t(pad)
    c = t(conv).requires_grad_(True) if conv is not None else None
    lc, idx = dp(k)
    z = sp(lc, idx, f, p)
    if c is None:
        c = torch.randn_like(z).requires_grad_(True)
    out = sl(lc, idx, c, p)
    res = dict(lc=n(lc), idx=n(idx), z=n(z), out=n(out), conv=n(c))
    if go is not None:
        (out * t(go)).sum().backward(retain_graph=True)
        res["gk_slice"] = n(k.grad)
        res["gconv"] = n(c.grad)
        k.grad = None
    if gz is not None:
        (z * t(gz)).sum().backward()
        res["gk_splat"] = n(k.grad)
        res["gfeat"] = n(f.grad)
    torch.cuda.synchronize()
    return res


def oracle_block(keys, feat, pad, conv, go, gz, W, H, dim, reduce="max"):
    lc, idx = O.positions_fwd(keys, W, H, dim)
    z, arg = O.splat_fwd(lc, idx, feat, W, H, dim, pad, reduce=reduce, return_arg=True)
    out = O.slice_fwd(lc, idx, conv, H, pad)
    gg, glc = O.slice_bwd(lc, idx, conv, go, H, pad)
    gf, glc2 = O.splat_bwd(lc, idx, feat, gz, arg, W, H, dim, pad, reduce=reduce)
    return dict(lc=lc, idx=idx, z=z, out=out, gconv=gg, gk_slice=O.positions_bwd(keys, glc, W, H, dim), gfeat=gf,
                gk_splat=O.positions_bwd(keys, glc2, W, H, dim), arg=arg)


def compare(res, ref, what, exact_z=True):
    assert np.array_equal(res["idx"], ref["idx"]), what + ": flattened_index must be bit-exact"
    assert np.array_equal(res["lc"], ref["lc"]), what + ": local_coordinate must be bit-exact"
    if exact_z:
        assert np.array_equal(res["z"], ref["z"]), what + ": Splat-max grid must be bit-exact"
    else:
        assert_close(res["z"], ref["z"], what + " z")
    for k in ("out", "gconv", "gk_slice", "gfeat", "gk_splat"):
        assert_close(res[k], ref[k], what + " " + k)


# --- golden fixtures minted from the reference -------------------------------------------------------
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_against_reference_golden(golden, name, fused, mode):
    dim, W, H, F, N, B = GOLDEN_CASES[name]
    g = {k.split("/", 1)[1]: golden[k] for k in golden.files if k.startswith(name + "/")}
    res = run_block(g["keys"], g["feat"], g.get("pad"), g["conv"], g["go"], g["gz"], W, H, dim, fused, mode)
    ref = dict(g)
    ref["idx"] = g["idx"].astype(np.int64)
    compare(res, ref, "%s fused=%s %s" % (name, fused, mode))


def test_adversarial_indices_bit_exact(golden):
    ak = golden["adv/keys"]
    for dim in (2, 3):
        for W in ((8, 16, 32, 64, 128, 256) if dim == 2 else (8, 16, 32, 64)):
            keys = np.tile(ak[None, None, :], (1, dim, 1)).astype(np.float32)
            dp = ctb.DifferentiablePositions(tensor_size=W, heads=1, dim=dim)
            lc, idx = dp(t(keys))
            assert np.array_equal(n(idx), golden["adv/idx_d%d_w%d" % (dim, W)].astype(np.int64))
            assert np.array_equal(n(lc), golden["adv/lc_d%d_w%d" % (dim, W)])


def test_indices_bit_exact_on_1e7_keys():
    rng = np.random.default_rng(0)
    for dim, W in [(2, 128), (3, 32)]:
        N = 1 << 20
        keys = np.tanh(rng.standard_normal((5 if dim == 2 else 4, dim, N)) * 1.5).astype(np.float32)
        dp = ctb.DifferentiablePositions(tensor_size=W, heads=1, dim=dim)
        lc, idx = dp(t(keys))
        lc_o, idx_o = O.positions_fwd(keys, W, 1, dim)
        assert np.array_equal(n(idx), idx_o) and np.array_equal(n(lc), lc_o)


# --- oracle on seeded inputs: every model shape class at reduced batch -------------------------------
SHAPES = [
    # dim, W, H, F, N, B
    (2, 128, 4, 4, 2048, 2), (3, 32, 4, 4, 2048, 2), (2, 64, 4, 16, 2048, 1), (3, 16, 4, 16, 2048, 1),
    (2, 16, 4, 16, 2048, 1), (3, 8, 4, 32, 2048, 1), (2, 64, 16, 4, 2048, 2), (3, 32, 2, 4, 4096, 1),
    (2, (24, 40), 3, 5, 777, 2), (3, (6, 9, 12), 2, 3, 1000, 2), (2, 8, 2, 1, 1, 1), (3, 4, 1, 2, 5, 3),
]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "d%d_w%s_h%d_f%d_n%d_b%d" % s)
def test_against_oracle(shape, mode):
    dim, W, H, F, N, B = shape
    keys, feat, pad = make_inputs(hash(shape) % 1000, B, H, dim, F, N, pad=(N % 2 == 1))
    C = int(np.prod(O._sizes(W, dim)))
    rng = np.random.default_rng(5)
    conv = rng.standard_normal((B, H * F) + tuple(O._sizes(W, dim))).astype(np.float32)
    go = rng.standard_normal((B, H * F, N)).astype(np.float32)
    gz = rng.standard_normal(conv.shape).astype(np.float32)
    ref = oracle_block(keys, feat, pad, conv, go, gz, W, H, dim)
    for fused in (True, False):
        res = run_block(keys, feat, pad, conv, go, gz, W, H, dim, fused, mode)
        compare(res, ref, "fused=%s %s" % (fused, mode))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("dist", ["uniform", "onecell"])
def test_collision_extremes(dist, mode):
    """uniform keys and the adversarial all-points-in-one-cell case."""
    for dim, W, H, F, N, B in [(2, 32, 2, 4, 1024, 1), (3, 8, 2, 8, 1024, 1)]:
        keys, feat, pad = make_inputs(11, B, H, dim, F, N, dist=dist)
        rng = np.random.default_rng(6)
        conv = rng.standard_normal((B, H * F) + (W,) * dim).astype(np.float32)
        go = rng.standard_normal((B, H * F, N)).astype(np.float32)
        gz = rng.standard_normal(conv.shape).astype(np.float32)
        ref = oracle_block(keys, feat, pad, conv, go, gz, W, H, dim)
        res = run_block(keys, feat, pad, conv, go, gz, W, H, dim, True, mode)
        compare(res, ref, "%s %s d%d" % (dist, mode, dim))


def test_edge_keys_and_zero_floor():
    """keys exactly +-1 / beyond, all-negative features (nothing beats the 0 floor), fully padded cloud."""
    dim, W, H, F, N, B = 2, 16, 2, 3, 64, 2
    keys, feat, _ = make_inputs(2, B, H, dim, F, N)
    keys[0, :, :8] = np.array([1.0, -1.0, 2.0, -3.0, 0.99999994, -0.99999994, 0.0, 1.0], dtype=np.float32)
    feat[0] = -np.abs(feat[0])
    pad = np.ones((B, N), dtype=np.float32)
    pad[1] = 0.0
    for mode in MODES:
        ctb.config.mode = mode
        dp = ctb.DifferentiablePositions(tensor_size=W, heads=H, dim=dim)
        sp = ctb.Splat(tensor_size=W, heads=H, dim=dim)
        k = t(keys).requires_grad_(True)
        f = t(feat).requires_grad_(True)
        lc, idx = dp(k)
        z = sp(lc, idx, f, t(pad))
        assert float(z.abs().max()) == 0.0          # batch 0 all negative, batch 1 fully padded
        assert int(idx.min()) >= 0 and int(idx.max()) < W * W
        z.sum().backward()
        assert float(f.grad.abs().max()) == 0.0 and float(k.grad.abs().max()) == 0.0


@pytest.mark.parametrize("mode", MODES)
def test_tie_breaking_first_winner(mode):
    """Duplicate points => exact ties; the gradient goes to exactly one winner, the lowest e = s*N + n
    (torch-scatter CPU rule)."""
    dim, W, H, F, N, B = 2, 8, 1, 2, 32, 1
    keys, feat, _ = make_inputs(4, B, H, dim, F, N)
    keys[:, :, 16:] = keys[:, :, :16]
    feat[:, :, 16:] = feat[:, :, :16]
    rng = np.random.default_rng(3)
    conv = rng.standard_normal((B, H * F, W, W)).astype(np.float32)
    go = rng.standard_normal((B, H * F, N)).astype(np.float32)
    gz = rng.standard_normal(conv.shape).astype(np.float32)
    ref = oracle_block(keys, feat, None, conv, go, gz, W, H, dim)
    res = run_block(keys, feat, None, conv, go, gz, W, H, dim, True, mode)
    compare(res, ref, "ties " + mode)
    assert np.abs(res["gfeat"][:, :, 16:]).max() == 0.0   # the duplicates (higher n) never win


@pytest.mark.parametrize("mode", MODES)
def test_reduce_sum_variant(mode):
    for dim, W, H, F, N, B in [(2, 16, 2, 4, 512, 2), (3, 8, 2, 4, 300, 1)]:
        keys, feat, pad = make_inputs(9, B, H, dim, F, N, pad=True)
        rng = np.random.default_rng(8)
        conv = rng.standard_normal((B, H * F) + (W,) * dim).astype(np.float32)
        go = rng.standard_normal((B, H * F, N)).astype(np.float32)
        gz = rng.standard_normal(conv.shape).astype(np.float32)
        ref = oracle_block(keys, feat, pad, conv, go, gz, W, H, dim, reduce="sum")
        res = run_block(keys, feat, pad, conv, go, gz, W, H, dim, True, mode, reduce="sum")
        compare(res, ref, "sum " + mode, exact_z=False)
        if mode == "deterministic":
            # fixed summation order (windows of the sorted entry list, folded in window order): bit-reproducible
            res2 = run_block(keys, feat, pad, conv, go, gz, W, H, dim, True, mode, reduce="sum")
            assert np.array_equal(res["z"], res2["z"]) and np.array_equal(res["gconv"], res2["gconv"])


def test_deterministic_mode_is_bit_reproducible_and_matches_atomic():
    dim, W, H, F, N, B = 3, 16, 4, 16, 2048, 2
    keys, feat, _ = make_inputs(21, B, H, dim, F, N)
    rng = np.random.default_rng(2)
    conv = rng.standard_normal((B, H * F, W, W, W)).astype(np.float32)
    go = rng.standard_normal((B, H * F, N)).astype(np.float32)
    gz = rng.standard_normal(conv.shape).astype(np.float32)
    runs = [run_block(keys, feat, None, conv, go, gz, W, H, dim, True, "deterministic") for _ in range(3)]
    for r in runs[1:]:
        for k in ("z", "out", "gconv", "gk_slice", "gfeat", "gk_splat"):
            assert np.array_equal(r[k], runs[0][k]), "deterministic mode must be bit-identical run to run: " + k
    for other in ("atomic", "tile"):
        at = run_block(keys, feat, None, conv, go, gz, W, H, dim, True, other)
        assert np.array_equal(at["z"], runs[0]["z"]) and np.array_equal(at["out"], runs[0]["out"])
        assert np.array_equal(at["gfeat"], runs[0]["gfeat"])
        assert_close(at["gconv"], runs[0]["gconv"], other + " vs deterministic grad grid")
    # Splat-max of the tile mode needs no sort to be reproducible (max + min-e are order independent)
    t1 = run_block(keys, feat, None, conv, go, gz, W, H, dim, True, "tile")
    t2 = run_block(keys, feat, None, conv, go, gz, W, H, dim, True, "tile")
    for k in ("z", "out", "gfeat", "gk_splat", "gk_slice"):
        assert np.array_equal(t1[k], t2[k]), k


def test_arg_identical_across_algorithms():
    """The winner index written by the atomic (max + min-e pass) and the sorted kernels is the same tensor."""
    for dim, W, H, F, N, B in [(2, 32, 4, 4, 2048, 2), (3, 8, 2, 8, 1024, 1)]:
        keys, feat, pad = make_inputs(13, B, H, dim, F, N, pad=True)
        keys[:, :, N // 2:] = keys[:, :, :N // 2]      # force ties
        feat[:, :, N // 2:] = feat[:, :, :N // 2]
        geom = CF.Geometry(O._sizes(W, dim), H, dim)
        sh = geom.shape(B, F, N)
        args = []
        feat_d, pad_d = t(feat), t(pad)
        for mode in (_lib.MODE_ATOMIC, _lib.MODE_DETERMINISTIC, _lib.MODE_TILE):
            h = CF.PositionsHandle(t(keys), geom)
            z = torch.empty((B, H * F, geom.C), device=DEV)
            arg = torch.empty((B, H * F, geom.C), dtype=torch.int32, device=DEV)
            plan = h.plan() if mode == _lib.MODE_DETERMINISTIC else None
            import ctypes
            CF._call("ctb_splat_fwd_keys", CF._ptr(h.keys_c()), CF._ptr(feat_d), CF._ptr(pad_d), CF._ptr(z),
                     CF._ptr(arg), ctypes.byref(sh), 0, mode, CF._ptr(plan), CF._stream(z))
            torch.cuda.synchronize()
            args.append((n(z), n(arg)))
        for other in args[1:]:
            assert np.array_equal(args[0][0], other[0]) and np.array_equal(args[0][1], other[1])
        lc, idx = O.positions_fwd(keys, W, H, dim)
        z_o, arg_o = O.splat_fwd(lc, idx, feat, W, H, dim, pad, return_arg=True)
        arg_o = np.where(arg_o == (1 << dim) * N, -1, arg_o).reshape(B, H * F, geom.C)
        assert np.array_equal(args[0][1], arg_o.astype(np.int32))


# --- full BASELINE sizes: size-independent properties -------------------------------------------------
FULL = [(2, 128, 16, 4, 2048, 32), (3, 32, 16, 4, 2048, 32), (2, 64, 16, 16, 2048, 32), (3, 16, 16, 16, 2048, 32),
        (2, 16, 16, 16, 2048, 32), (3, 8, 16, 32, 2048, 32), (3, 32, 16, 4, 4096, 8)]


@pytest.mark.parametrize("shape", FULL, ids=lambda s: "d%d_w%d_h%d_f%d_n%d_b%d" % s)
def test_full_size_properties(shape):
    dim, W, H, F, N, B = shape
    g = torch.Generator(device=DEV).manual_seed(1)
    keys = torch.tanh(torch.randn(B, H * dim, N, device=DEV, generator=g))
    feat = torch.randn(B, H * F, N, device=DEV, generator=g)
    outs = {}
    for mode in MODES:
        ctb.config.mode = mode
        dp = ctb.DifferentiablePositions(tensor_size=W, heads=H, dim=dim)
        sp = ctb.Splat(tensor_size=W, heads=H, dim=dim)
        sl = ctb.Slice(tensor_size=W, heads=H, dim=dim)
        k = keys.clone().requires_grad_(True)
        f = feat.clone().requires_grad_(True)
        lc, idx = dp(k)
        z = sp(lc, idx, f)
        out = sl(lc, idx, z)
        gsum = torch.autograd.grad(out.sum(), [k, f], retain_graph=True)
        # Slice is linear in the grid: slice(2 z) == 2 slice(z)
        out2 = sl(lc, idx, 2.0 * z)
        assert torch.allclose(out2, 2.0 * out, rtol=1e-5, atol=1e-6)
        # weights of a point sum to 1 => slicing a constant grid returns the constant
        ones = sl(lc, idx, torch.ones_like(z))
        assert float((ones - 1).abs().max()) < 1e-5
        # z >= 0 (zero floor) and every positive cell is attained by some point: z <= max positive feature
        assert float(z.min()) >= 0.0
        assert float(z.max()) <= float(feat.clamp(min=0).max()) + 1e-6
        # index range
        assert int(idx.min()) >= 0 and int(idx.max()) < W ** dim
        outs[mode] = (z, out, gsum)
    zd, od, gd = outs["deterministic"]
    for other in ("atomic", "tile"):
        za, oa, ga = outs[other]
        assert torch.equal(za, zd), "Splat-max grid: all three algorithms must agree bit for bit"
        assert torch.equal(oa, od)
        assert torch.allclose(ga[1], gd[1], rtol=1e-4, atol=1e-5)
        assert torch.allclose(ga[0], gd[0], rtol=1e-4, atol=1e-4 * float(ga[0].abs().max()))
    # cross-check one (b, h) unit of the full-size run against the oracle
    kb = n(keys[:1, :dim]); fb = n(feat[:1, :F])
    lc_o, idx_o = O.positions_fwd(kb, W, 1, dim)
    z_o = O.splat_fwd(lc_o, idx_o, fb, W, 1, dim)
    assert np.array_equal(n(zd[:1, :F]), z_o)


def test_occupancy_count():
    z = torch.randn(3, 8, 16, 16, device=DEV)
    z[z.abs() < 0.5] = 0
    assert int(CF.count_occupied(z)) == int((z.abs() > 1e-9).sum())


def test_non_default_stream_and_noncontiguous_inputs():
    dim, W, H, F, N, B = 2, 16, 2, 4, 256, 2
    keys, feat, _ = make_inputs(1, B, H, dim, F, N)
    ref = run_block(keys, feat, None, None, None, None, W, H, dim)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        dp = ctb.DifferentiablePositions(tensor_size=W, heads=H, dim=dim)
        sp = ctb.Splat(tensor_size=W, heads=H, dim=dim)
        k = t(np.ascontiguousarray(keys.transpose(0, 2, 1))).transpose(1, 2)   # non-contiguous view
        f = t(np.ascontiguousarray(feat.transpose(0, 2, 1))).transpose(1, 2)
        lc, idx = dp(k)
        z = sp(lc, idx, f)
    s.synchronize()
    assert np.array_equal(n(z), ref["z"])


def test_tile_sum_is_order_independent_and_accurate():
    """TILE-mode scatter-add accumulates fixed-point integers with native shared-memory atomics (integer addition
    is associative): the result is bit-identical from run to run AND invariant under any permutation of the
    points -- stronger than a fixed summation order -- and closer to the exact sum than float accumulation:
    |err| <= 2^-24 |exact| + N * 2^-39 * max|feature|."""
    # the last four rows walk the channel-lane kernel's variants: paired half-warps with idle channel lanes and a
    # ragged last chunk, odd last axis (unpaired), 8-lane groups, and the paired 2-D case
    for dim, W, H, F, N, B in [(3, 8, 2, 32, 2048, 2), (2, 64, 2, 16, 2048, 1), (3, 32, 2, 4, 2048, 1),
                               (2, 16, 2, 8, 5000, 1), (3, 8, 2, 12, 1000, 1), (3, (8, 8, 7), 2, 16, 777, 1),
                               (2, 16, 2, 5, 300, 2), (2, 16, 2, 16, 2048, 1)]:
        keys, feat, pad = make_inputs(31, B, H, dim, F, N, pad=True)
        geom = CF.Geometry(O._sizes(W, dim), H, dim)
        rng = np.random.default_rng(4)
        perm = rng.permutation(N)
        outs = []
        for k_, f_, p_ in ((keys, feat, pad), (keys[:, :, perm], feat[:, :, perm], pad[:, perm])):
            ctb.config.mode = "tile"
            ctb.config.use_plan = False      # the plan-free tile scatters (with a plan, dense grids sum in a fixed order)
            h = CF.PositionsHandle(t(k_), geom)
            with torch.no_grad():
                outs.append(n(CF.fused_splat(h, t(f_), t(p_), _lib.REDUCE_SUM)))
        assert np.array_equal(outs[0], outs[1]), "sum must not depend on the order of the points"
        lc, idx = O.positions_fwd(keys, W, H, dim)
        pre = O._pre_splat(lc, feat, pad, H).astype(np.float64)
        Bq, Hq, Fq, Sq, Nq = pre.shape
        C = int(np.prod(O._sizes(W, dim)))
        ex = np.zeros((Bq * Hq * Fq, C))
        rows = np.broadcast_to(np.arange(Bq * Hq * Fq)[:, None], (Bq * Hq * Fq, Sq * Nq))
        index = np.broadcast_to(idx.reshape(Bq, Hq, 1, Sq * Nq), (Bq, Hq, Fq, Sq * Nq)).reshape(Bq * Hq * Fq, Sq * Nq)
        np.add.at(ex, (rows, index), pre.reshape(Bq * Hq * Fq, Sq * Nq))
        ex = ex.reshape(outs[0].shape)
        err = np.abs(outs[0].astype(np.float64) - ex)
        bound = 2.0 ** -24 * np.abs(ex) + N * 2.0 ** -39 * float(np.abs(feat).max())
        assert (err <= bound).all(), float((err - bound).max())


def test_sorted_slice_forward_is_bit_identical_to_the_tile_gather_and_matches_the_oracle():
    """Dense grids: with a plan, Slice forward walks the points in cell-sorted order (ctb_sgather.cuh) -- same
    arithmetic per (point, channel), so the same bits as the index-order tile gather; layers/cloud_transform.py:204-211."""
    lib = _lib.load()
    for dim, W, H, F, N, B, pad_on in [(3, 8, 2, 32, 2048, 2, False), (2, 16, 2, 16, 2048, 1, True), (3, 16, 2, 16, 2048, 1, True),
                                        (3, 8, 2, 12, 1000, 1, True), (2, 16, 3, 5, 300, 2, False)]:
        keys, feat, pad = make_inputs(17, B, H, dim, F, N, pad=pad_on, dist="onecell" if N == 300 else "tanh")
        sizes = O._sizes(W, dim)
        geom = CF.Geometry(sizes, H, dim)
        rng = np.random.default_rng(5)
        conv = rng.standard_normal((B, H * F) + tuple(sizes)).astype(np.float32)
        sh = geom.shape(B, F, N, _lib.DTYPE_F32)
        assert lib.ctb_op_uses_plan(ctypes.byref(sh), _lib.OP_SLICE_FWD, 0, _lib.MODE_TILE) == 1, (dim, W, F, N)
        outs = []
        for use_plan in (True, False):
            ctb.config.mode = "tile"
            ctb.config.use_plan = use_plan
            h = CF.PositionsHandle(t(keys), geom)
            with torch.no_grad():
                if use_plan:
                    h.plan()                      # as after a plan-based Splat forward
                outs.append(n(CF.fused_slice(h, t(conv), t(pad))))
        assert np.array_equal(outs[0], outs[1])
        lc, idx = O.positions_fwd(keys, W, H, dim)
        assert_close(outs[0], O.slice_fwd(lc, idx, conv, H, pad), "sorted slice fwd %s" % ((dim, W, F, N),))


def test_tile_sum_limb_headroom_worst_case():
    """Every one of the 2048 points sits exactly on one grid node (corner weight 1) and carries the largest
    mantissa: the per-cell limb sums reach their design maximum.  The sum 2048 * (2 - 2^-23) is representable, so
    the result must be exact -- any overflow of a 32-bit limb would show."""
    big = np.float32(2.0) - np.float32(2.0 ** -23)
    for dim, W, F in [(3, 5, 16), (2, 65, 4), (2, 9, 16)]:
        N, H, B = 2048, 1, 1
        keys = np.full((B, H * dim, N), -0.5, np.float32)          # (k + 1) * (W - 1) / 2 is an integer
        for sign in (1.0, -1.0):
            feat = np.full((B, H * F, N), sign * big, np.float32)
            ctb.config.mode = "tile"
            h = CF.PositionsHandle(t(keys), CF.Geometry(O._sizes(W, dim), H, dim))
            with torch.no_grad():
                z = n(CF.fused_splat(h, t(feat), None, _lib.REDUCE_SUM))
            node = (W - 1) // 4
            want = np.zeros_like(z)
            want[(0, slice(None)) + (node,) * dim] = np.float32(sign) * (np.float32(4096.0) - np.float32(2.0 ** -12))
            assert np.array_equal(z, want), (dim, W, F, sign, float(np.abs(z - want).max()))


def test_empty_cloud_matches_reference_shapes():
    """N = 0 (and B = 0): the reference's torch ops return a zero grid / empty per-point tensors; so does the mirror,
    and gradients flow (as zeros) instead of raising."""
    for B, N in ((2, 0), (0, 5)):
        H, dim, W, F = 2, 3, 4, 3
        dp = ctb.DifferentiablePositions(tensor_size=W, heads=H, dim=dim)
        sp = ctb.Splat(tensor_size=W, heads=H, dim=dim)
        sl = ctb.Slice(tensor_size=W, heads=H, dim=dim)
        keys = torch.zeros(B, H * dim, N, device=DEV, requires_grad=True)
        feat = torch.zeros(B, H * F, N, device=DEV, requires_grad=True)
        lc, idx = dp(keys)
        assert lc.shape == (B, H, 8, N) and idx.shape == (B, H, 8, N) and idx.dtype == torch.int64
        z = sp(lc, idx, feat)
        assert z.shape == (B, H * F, W, W, W) and float(z.abs().sum()) == 0.0
        conv = torch.randn(B, H * F, W, W, W, device=DEV, requires_grad=True)
        out = sl(lc, idx, conv)
        assert out.shape == (B, H * F, N)
        (z.sum() + out.sum()).backward()
        assert feat.grad is not None and feat.grad.shape == feat.shape
        assert conv.grad is not None and float(conv.grad.abs().sum()) == 0.0


# --- A8: fused projection + tanh, and the MHCT block mirror ------------------------------------------
def _torch_keys(pcd, keys_res, shift, log_R, scales, res_scale, H, dim):
    """layers/utils.py:25-34 / :53-61 + multihead_ct.py:93-97 in plain torch."""
    from cloud_transformers_b200.so3 import so3_exponential_map
    B, _, N = pcd.shape
    p = pcd[:, None] + (res_scale if res_scale is not None else 1.0) * keys_res.reshape(B, H, 3, N)
    p = p + shift[None, :, :, None]
    q = torch.einsum('bhcp,hcn->bhnp', p, so3_exponential_map(log_R))[:, :, :dim]
    if scales is not None:
        q = q * scales[None, :, :, None]
    return torch.tanh(q.reshape(B, H * dim, N))


@pytest.mark.parametrize("dim,with_scales,with_rs", [(2, False, False), (3, True, False), (3, False, True), (2, True, True)])
def test_fused_projection_matches_torch(dim, with_scales, with_rs):
    from cloud_transformers_b200.so3 import so3_exponential_map
    B, H, N = 3, 5, 777
    g = torch.Generator(device=DEV).manual_seed(dim * 10 + with_scales)
    pcd = (torch.rand(B, 3, N, device=DEV, generator=g) * 2 - 1).requires_grad_(True)
    res = (0.3 * torch.randn(B, H * 3, N, device=DEV, generator=g)).requires_grad_(True)
    shift = (0.1 * torch.randn(H, 3, device=DEV, generator=g)).requires_grad_(True)
    log_R = torch.randn(H, 3, device=DEV, generator=g).requires_grad_(True)
    scales = (1 + 0.2 * torch.randn(H, dim, device=DEV, generator=g)).requires_grad_(True) if with_scales else None
    rs = torch.tensor(0.7, device=DEV, requires_grad=True) if with_rs else None
    gk = torch.randn(B, H * dim, N, device=DEV, generator=g)
    ref = _torch_keys(pcd, res, shift, log_R, scales, rs, H, dim)
    inputs = [t_ for t_ in (pcd, res, shift, log_R, scales, rs) if t_ is not None]
    g_ref = torch.autograd.grad((ref * gk).sum(), inputs)
    out = CF.project_keys(pcd, res, shift, so3_exponential_map(log_R), scales, rs, heads=H, dim=dim)
    g_out = torch.autograd.grad((out * gk).sum(), inputs)
    assert_close(n(out), n(ref), "keys", rtol=1e-5, atol_scale=2e-6)
    for a, b_, name in zip(g_out, g_ref, ("pcd", "res", "shift", "log_R", "scales", "res_scale")):
        assert_close(n(a), n(b_), "grad " + name, rtol=1e-4, atol_scale=1e-4)


@pytest.mark.parametrize("dim,W,F", [(2, 16, 8), (3, 8, 4)])
def test_mhct_block_mirror_matches_unfused_composition(dim, W, F):
    """MultiHead mirror (fused prologue + fused Splat / Slice) against the same block assembled from torch ops and the
    reference-API modules.  tanhf may differ from torch.tanh in the last ulp, which can move a point that sits on a
    cell boundary: the comparison allows a tiny fraction of outliers."""
    from cloud_transformers_b200 import mhct
    torch.manual_seed(0)
    B, H, N, M = 2, 4, 512, 32
    blk = mhct.MultiHead(model_dim=M, in_feature_dim=F, out_model_dim=M, tensor_size=W, tensor_dim=dim, heads=H,
                         scales=True).to(DEV)
    with torch.no_grad():
        blk.key_bn.weight.normal_(0, 0.3)      # zero-init would switch the learned offsets off
        blk.transform.scales.normal_(1, 0.1)
    x = torch.randn(B, M, N, device=DEV, requires_grad=True)
    pcd = (torch.rand(B, 3, N, device=DEV) * 2 - 1).requires_grad_(True)
    out, stats = blk(x, pcd)
    gout = torch.randn_like(out)
    params = [p for p in blk.parameters() if p.requires_grad]
    g_fused = torch.autograd.grad((out * gout).sum(), [x, pcd] + params, allow_unused=True)

    def unfused():
        kv = blk.keys_values_pred(x)
        keys_res = blk.key_bn(kv[:, :H * 3])
        values = blk.values_bn(kv[:, H * 3:])
        lattice = _torch_keys(pcd, keys_res, blk.transform.shift, blk.transform.log_R, blk.transform.scales, None, H, dim)
        ctb.config.fused = False
        lc, idx = blk.diff_poss(lattice)
        z = blk.splat(lc, idx, values)
        return blk.after(blk.slice(lc, idx, blk.conv(z)))

    ref = unfused()
    g_ref = torch.autograd.grad((ref * gout).sum(), [x, pcd] + params, allow_unused=True)
    ctb.config.fused = True
    bad = (out - ref).abs() > 1e-4 * ref.abs().max() + 1e-4 * ref.abs()
    assert float(bad.float().mean()) < 2e-3, float(bad.float().mean())
    names = ["x", "pcd"] + [k for k, p in blk.named_parameters() if p.requires_grad]
    rels = {}
    for name, a, b_ in zip(names, g_fused, g_ref):
        if a is None or b_ is None:
            assert a is None and b_ is None, name
            continue
        rels[name] = (float((a - b_).norm()), float(b_.norm()))
    top = max(v[1] for v in rels.values())
    # gradients that are zero in exact arithmetic (e.g. a bias in front of a BatchNorm) are pure rounding noise:
    # judge every tensor against its own norm plus a floor tied to the largest gradient
    bad = {k: v for k, v in rels.items() if v[0] > 2e-2 * v[1] + 1e-5 * top}
    assert not bad, bad
    assert float(stats[0]) > 0


# --- other BASELINE configs and the sweep extremes ------------------------------
BIG = [
    # dim, W, H, F, N, B
    (3, 32, 16, 4, 4096, 2),      # S3DIS 1x1 block, 3D 32^3
    (2, 128, 16, 4, 16384, 1),    # inpainting decoder, 2D 128^2
    (3, 32, 16, 4, 16384, 1),     # inpainting decoder, 3D 32^3
    (2, 256, 4, 4, 65536, 1),     # sweep: 2D 256^2, N = 64K  (tile kernels at their point-count limit)
    (3, 64, 4, 4, 32768, 1),      # sweep: 3D 64^3 (one channel plane = 1 MB > shared memory => row slabs)
    (2, 32, 64, 32, 1024, 1),     # sweep: 64 heads, F = 32
    (2, 256, 2, 4, 262144, 1),    # sweep max: N = 256K (beyond the tile kernels => L2-atomic kernels)
    (2, 16, 2, 16, 8192, 1),      # 2 units only: the gathers split a unit's points over 4 CTAs (channel-last tile)
    (3, 16, 2, 16, 16384, 1),     # same with plane-major TMA tiles and several channel groups
]


@pytest.mark.parametrize("shape", BIG, ids=lambda s: "d%d_w%d_h%d_f%d_n%d_b%d" % s)
def test_large_shapes_all_algorithms_agree(shape):
    dim, W, H, F, N, B = shape
    g = torch.Generator(device=DEV).manual_seed(7)
    keys = torch.tanh(torch.randn(B, H * dim, N, device=DEV, generator=g))
    feat = torch.randn(B, H * F, N, device=DEV, generator=g)
    res = {}
    for mode in ("atomic", "auto"):
        ctb.config.mode = mode
        dp = ctb.DifferentiablePositions(tensor_size=W, heads=H, dim=dim)
        sp = ctb.Splat(tensor_size=W, heads=H, dim=dim)
        sl = ctb.Slice(tensor_size=W, heads=H, dim=dim)
        k = keys.clone().requires_grad_(True)
        f = feat.clone().requires_grad_(True)
        lc, idx = dp(k)
        z = sp(lc, idx, f)
        conv = torch.sin(z * 3 + 0.1)
        out = sl(lc, idx, conv)
        gk, gf = torch.autograd.grad(out.square().sum(), [k, f])
        res[mode] = (z.detach(), out.detach(), gk, gf)
    za, oa, gka, gfa = res["atomic"]
    zt, ot, gkt, gft = res["auto"]
    assert torch.equal(za, zt), "Splat-max grid must agree bit for bit across algorithms"
    assert torch.allclose(oa, ot, rtol=1e-5, atol=1e-6)
    assert torch.allclose(gfa, gft, rtol=1e-4, atol=1e-4 * float(gfa.abs().max()))
    assert torch.allclose(gka, gkt, rtol=1e-4, atol=1e-4 * float(gka.abs().max()))
    # one unit against the oracle
    kb, fb = n(keys[:1, :dim]), n(feat[:1, :F])
    lc_o, idx_o = O.positions_fwd(kb, W, 1, dim)
    z_o = O.splat_fwd(lc_o, idx_o, fb, W, 1, dim)
    assert np.array_equal(n(zt[:1, :F]), z_o)
    assert_close(n(ot[:1, :F]), O.slice_fwd(lc_o, idx_o, np.sin(z_o * 3 + 0.1).astype(np.float32), 1), "slice")


# --- bf16 grid storage mode (north star: rel 1e-2) -----------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 128, 4, 4, 2048, 2), (3, 32, 4, 4, 2048, 1), (2, 64, 4, 16, 2048, 1),
                                   (3, 8, 4, 32, 2048, 1), (2, (24, 40), 3, 5, 777, 2), (2, (10, 12), 2, 3, 500, 1),
                                   (3, 16, 2, 16, 16384, 1)],
                         ids=lambda s: "d%d_w%s_h%d_f%d_n%d_b%d" % s)
def test_bf16_grid_storage_mode(shape):
    """Grids (z, convolved, grad_grid, grad_z) stored as bf16, arithmetic in fp32: against the fp32 oracle fed with the
    bf16-rounded grids, within rel 1e-2 (abs floor 1e-2 * max|ref|)."""
    dim, W, H, F, N, B = shape
    keys, feat, pad = make_inputs(17, B, H, dim, F, N, pad=True)
    sizes = tuple(O._sizes(W, dim))
    rng = np.random.default_rng(9)
    bf = lambda a: torch.from_numpy(a).to(torch.bfloat16).float().numpy()
    conv = bf(rng.standard_normal((B, H * F) + sizes).astype(np.float32))
    go = rng.standard_normal((B, H * F, N)).astype(np.float32)
    gz = bf(rng.standard_normal(conv.shape).astype(np.float32))
    ref = oracle_block(keys, feat, pad, conv, go, gz, W, H, dim)
    ctb.config.mode = "tile"
    dp = ctb.DifferentiablePositions(tensor_size=W, heads=H, dim=dim)
    sp = ctb.Splat(tensor_size=W, heads=H, dim=dim, out_dtype=torch.bfloat16)
    sl = ctb.Slice(tensor_size=W, heads=H, dim=dim)
    k = t(keys).requires_grad_(True)
    f = t(feat).requires_grad_(True)
    c = t(conv).to(torch.bfloat16).requires_grad_(True)
    lc, idx = dp(k)
    z = sp(lc, idx, f, t(pad))
    assert z.dtype == torch.bfloat16
    out = sl(lc, idx, c, t(pad))
    assert out.dtype == torch.float32
    assert_close(n(z.float()), ref["z"], "bf16 z", rtol=1e-2, atol_scale=1e-2)
    assert_close(n(out), ref["out"], "slice of bf16 grid", rtol=1e-2, atol_scale=1e-2)
    (out * t(go)).sum().backward(retain_graph=True)
    assert c.grad.dtype == torch.bfloat16
    assert_close(n(c.grad.float()), ref["gconv"], "bf16 grad grid", rtol=1e-2, atol_scale=1e-2)
    assert_close(n(k.grad), ref["gk_slice"], "grad keys via slice", rtol=1e-2, atol_scale=1e-2)
    k.grad = None
    (z.float() * t(gz)).sum().backward()
    assert_close(n(f.grad), ref["gfeat"], "grad features", rtol=1e-2, atol_scale=1e-2)
    assert_close(n(k.grad), ref["gk_splat"], "grad keys via splat", rtol=1e-2, atol_scale=1e-2)


def test_fused_path_is_dropped_when_keys_change_in_place():
    """lc / idx describe the positions at the time DifferentiablePositions ran (the reference's Splat / Slice only see
    those tensors): if the keys are modified in place afterwards, the fused path -- which would recompute the positions
    from the new keys -- must not be taken."""
    dim, W, H, F, N, B = 2, 16, 2, 4, 300, 2
    keys, feat, _ = make_inputs(3, B, H, dim, F, N)
    dp = ctb.DifferentiablePositions(tensor_size=W, heads=H, dim=dim).to(DEV)
    sp = ctb.Splat(tensor_size=W, heads=H, dim=dim).to(DEV)
    sl = ctb.Slice(tensor_size=W, heads=H, dim=dim).to(DEV)
    with torch.no_grad():
        k = t(keys)
        lc, idx = dp(k)
        z_before = sp(lc, idx, t(feat))
        k.mul_(-0.5)                                   # in-place change of the keys after the positions were taken
        z_after = sp(lc, idx, t(feat))
        out_after = sl(lc, idx, z_after)
    lc_o, idx_o = O.positions_fwd(keys, W, H, dim)
    z_o = O.splat_fwd(lc_o, idx_o, feat, W, H, dim)
    assert np.array_equal(n(z_before), z_o) and np.array_equal(n(z_after), z_o)
    assert_close(n(out_after), O.slice_fwd(lc_o, idx_o, z_o, H), "slice after in-place key change")


@pytest.mark.parametrize("shape", FULL + BIG, ids=lambda s: "d%d_w%d_h%d_f%d_n%d_b%d" % s)
def test_full_and_big_shapes_one_unit_against_oracle_with_gradients(shape):
    """Every BASELINE-size shape, default algorithm: forward AND all four gradients of one (b, h) unit against the
    oracle.  Units are independent (each owns its grid slab, cloud_transform.py:164-178), so the oracle only has to run
    on the slice of the inputs that unit sees; the unit is taken from the middle of the batch and the heads."""
    dim, W, H, F, N, B = shape
    g = torch.Generator(device=DEV).manual_seed(11)
    keys = torch.tanh(torch.randn(B, H * dim, N, device=DEV, generator=g))
    feat = torch.randn(B, H * F, N, device=DEV, generator=g)
    grid = (B, H * F) + (W,) * dim
    conv = torch.randn(grid, device=DEV, generator=g)
    go = torch.randn(B, H * F, N, device=DEV, generator=g)
    gz = torch.randn(grid, device=DEV, generator=g)
    ctb.config.mode = "auto"
    dp = ctb.DifferentiablePositions(tensor_size=W, heads=H, dim=dim)
    sp = ctb.Splat(tensor_size=W, heads=H, dim=dim)
    sl = ctb.Slice(tensor_size=W, heads=H, dim=dim)
    k = keys.clone().requires_grad_(True)
    f = feat.clone().requires_grad_(True)
    c = conv.clone().requires_grad_(True)
    lc, idx = dp(k)
    z = sp(lc, idx, f)
    out = sl(lc, idx, c)
    gk_slice, gconv = torch.autograd.grad((out * go).sum(), [k, c], retain_graph=True)
    gk_splat, gfeat = torch.autograd.grad((z * gz).sum(), [k, f])
    b, h = B // 2, H // 2
    ks = slice(h * dim, (h + 1) * dim)
    fs = slice(h * F, (h + 1) * F)
    ref = oracle_block(n(keys[b:b + 1, ks]), n(feat[b:b + 1, fs]), None, n(conv[b:b + 1, fs]), n(go[b:b + 1, fs]),
                       n(gz[b:b + 1, fs]), W, 1, dim)
    what = "unit (%d, %d) of %s" % (b, h, (shape,))
    assert np.array_equal(n(idx[b:b + 1, h:h + 1]), ref["idx"]) and np.array_equal(n(lc[b:b + 1, h:h + 1]), ref["lc"])
    assert np.array_equal(n(z[b:b + 1, fs]), ref["z"]), what + ": Splat-max grid must be bit-exact"
    assert_close(n(out[b:b + 1, fs]), ref["out"], what + " out")
    assert_close(n(gconv[b:b + 1, fs]), ref["gconv"], what + " grad grid")
    assert_close(n(gfeat[b:b + 1, fs]), ref["gfeat"], what + " grad features")
    assert_close(n(gk_slice[b:b + 1, ks]), ref["gk_slice"], what + " grad keys (Slice)")
    assert_close(n(gk_splat[b:b + 1, ks]), ref["gk_splat"], what + " grad keys (Splat)")
