"""CPU suite, part 1: the oracle against the golden fixtures minted from the reference, against the real
reference (build container only), and the host build of the kernels' position arithmetic against the
oracle (bit-exact)."""
import ctypes
import os

import numpy as np
import pytest

from oracle import ct_numpy as O
from oracle import reference_loader as RL
from tests.util import GOLDEN_CASES, assert_close, make_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_matches_golden(golden, name):
    dim, W, H, F, N, B = GOLDEN_CASES[name]
    g = {k.split("/", 1)[1]: golden[k] for k in golden.files if k.startswith(name + "/")}
    pad = g.get("pad")
    lc, idx = O.positions_fwd(g["keys"], W, H, dim)
    assert np.array_equal(idx, g["idx"].astype(np.int64)), "flattened_index must be bit-exact"
    assert np.array_equal(lc, g["lc"]), "local_coordinate must be bit-exact"
    z, arg = O.splat_fwd(lc, idx, g["feat"], W, H, dim, pad, return_arg=True)
    assert np.array_equal(z, g["z"]), "Splat max is order independent => bit-exact"
    out = O.slice_fwd(lc, idx, g["conv"], H, pad)
    assert_close(out, g["out"], "slice fwd")
    gg, glc = O.slice_bwd(lc, idx, g["conv"], g["go"], H, pad)
    assert_close(gg, g["gconv"], "grad conv")
    assert_close(O.positions_bwd(g["keys"], glc, W, H, dim), g["gk_slice"], "grad keys via slice")
    gf, glc2 = O.splat_bwd(lc, idx, g["feat"], g["gz"], arg, W, H, dim, pad)
    assert_close(gf, g["gfeat"], "grad features")
    assert_close(O.positions_bwd(g["keys"], glc2, W, H, dim), g["gk_splat"], "grad keys via splat")


def test_oracle_adversarial_indices(golden):
    ak = golden["adv/keys"]
    for dim in (2, 3):
        for W in ((8, 16, 32, 64, 128, 256) if dim == 2 else (8, 16, 32, 64)):
            keys = np.tile(ak[None, None, :], (1, dim, 1))
            lc, idx = O.positions_fwd(keys, W, 1, dim)
            assert np.array_equal(idx, golden["adv/idx_d%d_w%d" % (dim, W)].astype(np.int64))
            assert np.array_equal(lc, golden["adv/lc_d%d_w%d" % (dim, W)])
            assert idx.min() >= 0 and idx.max() < W ** dim


def test_scatter_rule_loop_vs_vectorised():
    rng = np.random.default_rng(0)
    src = rng.standard_normal((5, 200)).astype(np.float32)
    src[:, ::7] = src[:, 3:4]          # inject exact ties
    src[1, :] = -np.abs(src[1, :])     # a row where nothing beats the zero floor
    index = rng.integers(0, 9, (5, 200))
    a, b = O.scatter_max_first(src, index, 9)
    c, d = O.scatter_max_loop(src, index, 9)
    assert np.array_equal(a, c) and np.array_equal(b, d)
    assert (b[1] == 200).all() and (a[1] == 0).all()


@pytest.mark.skipif(not RL.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("dim,W,H,F,N,B", [(2, 16, 4, 4, 300, 2), (3, 8, 3, 5, 257, 2), (3, (4, 6, 5), 2, 2, 64, 2)])
def test_oracle_matches_reference_live(dim, W, H, F, N, B):
    import torch
    from oracle import ct_torch as T
    ct, _, _ = RL.load_reference_layers()
    keys, feat, pad = make_inputs(7, B, H, dim, F, N, pad=True)
    dp = ct.DifferentiablePositions(tensor_size=W, heads=H, dim=dim)
    sp = ct.Splat(tensor_size=W, heads=H, dim=dim)
    sl = ct.Slice(tensor_size=W, heads=H, dim=dim)
    k = torch.from_numpy(keys).requires_grad_(True)
    f = torch.from_numpy(feat).requires_grad_(True)
    p = torch.from_numpy(pad)
    lc, idx = dp(k)
    z = sp(lc, idx, f, p)
    out = sl(lc, idx, z, p)
    lc_o, idx_o = O.positions_fwd(keys, W, H, dim)
    assert np.array_equal(idx_o, idx.numpy()) and np.array_equal(lc_o, lc.detach().numpy())
    z_o = O.splat_fwd(lc_o, idx_o, feat, W, H, dim, pad)
    assert np.array_equal(z_o, z.detach().numpy())
    assert_close(O.slice_fwd(lc_o, idx_o, z_o, H, pad), out.detach().numpy(), "slice")
    # the torch port used as the CPU baseline issues the same ops
    lc_t, idx_t = T.positions(torch.from_numpy(keys), W, H, dim)
    assert torch.equal(idx_t, idx) and torch.equal(lc_t, lc.detach())
    assert torch.equal(T.splat(lc_t, idx_t, torch.from_numpy(feat), W, H, dim, p), z.detach())


def _host_positions(keys, W, H, dim, grad_lc=None):
    import __graft_entry__ as ge
    lib = ctypes.CDLL(ge.HOST_LIB)
    B, _, N = keys.shape
    S = 1 << dim
    sizes = (ctypes.c_int32 * 3)(*(list(O._sizes(W, dim)) + [1] * (3 - dim)))
    lc = np.empty((B, H, S, N), dtype=np.float32)
    idx = np.empty((B, H, S, N), dtype=np.int64)
    gk = np.empty_like(keys) if grad_lc is not None else None
    keys = np.ascontiguousarray(keys)
    st = lib.ctb_host_positions(
        keys.ctypes.data_as(ctypes.c_void_p), lc.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p),
        gk.ctypes.data_as(ctypes.c_void_p) if gk is not None else None,
        np.ascontiguousarray(grad_lc).ctypes.data_as(ctypes.c_void_p) if grad_lc is not None else None,
        ctypes.c_int(B * H), ctypes.c_int(N), ctypes.c_int(dim), sizes)
    assert st == 0
    return lc, idx, gk


@pytest.mark.parametrize("dim,W", [(2, 128), (2, 64), (2, 16), (3, 32), (3, 16), (3, 8), (2, (6, 10)), (3, (4, 6, 5)),
                                   (2, 256), (3, 64)])
def test_kernel_position_arithmetic_bit_exact_on_host(dim, W, golden):
    """The exact arithmetic the CUDA kernels use (ctb_positions.cuh compiled for the host) against the
    oracle: >= 1e6 random keys plus the adversarial sweep, bit for bit."""
    rng = np.random.default_rng(dim * 1000 + hash(str(W)) % 997)
    n = 400000
    keys = np.concatenate([np.tanh(rng.standard_normal(n) * 1.5), rng.uniform(-1, 1, n),
                           golden["adv/keys"].astype(np.float64)]).astype(np.float32)
    N = keys.size
    k = np.stack([np.roll(keys, 17 * a) for a in range(dim)])[None]     # [1, dim, N]
    lc_o, idx_o = O.positions_fwd(k, W, 1, dim)
    lc_h, idx_h, _ = _host_positions(k, W, 1, dim)
    assert np.array_equal(idx_h, idx_o)
    assert np.array_equal(lc_h, lc_o)
    assert idx_h.min() >= 0 and idx_h.max() < int(np.prod(O._sizes(W, dim)))


def test_kernel_position_backward_on_host():
    for dim, W in [(2, 16), (3, 8)]:
        keys, _, _ = make_inputs(3, 2, 3, dim, 1, 500)
        keys[0, 0, :3] = [1.0, -1.0, 0.99999994]
        rng = np.random.default_rng(1)
        glc = rng.standard_normal((2, 3, 1 << dim, 500)).astype(np.float32)
        _, _, gk = _host_positions(keys, W, 3, dim, glc)
        assert_close(gk, O.positions_bwd(keys, glc, W, 3, dim), "positions bwd")
        assert gk[0, 0, 0] == 0 and gk[0, 0, 1] == 0      # clamped keys get no gradient


def test_positions_gradient_is_unscaled_finite_difference():
    """grad_keys equals d lc / d x WITHOUT the (W-1)/2 factor (GradientBalancing, cloud_transform.py:17-23)."""
    dim, W, H = 2, 16, 1
    keys = np.array([[[0.3], [-0.42]]], dtype=np.float32)
    glc = np.array([[[[1.0], [0.0], [0.0], [0.0]]]], dtype=np.float32)
    gk = O.positions_bwd(keys, glc, W, H, dim)
    h = 1e-3
    lc_p, _ = O.positions_fwd(keys + np.array([[[h], [0]]], dtype=np.float32), W, H, dim)
    lc_m, _ = O.positions_fwd(keys - np.array([[[h], [0]]], dtype=np.float32), W, H, dim)
    fd = (lc_p[0, 0, 0, 0] - lc_m[0, 0, 0, 0]) / (2 * h)
    scale = (W - 1) * 0.5
    assert abs(gk[0, 0, 0] * scale - fd) < 1e-2 * abs(fd)
