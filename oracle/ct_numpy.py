"""CPU oracle for the Splat / Slice hot path -- TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's algorithm for the path
    layers/cloud_transform.py  (DifferentiablePositions :72-121, Splat :131-180, Slice :190-227)
    layers/utils.py            (bilinear_coords :158-186, trilinear_coords :100-155)
plus hand-derived backward formulas matching what torch autograd produces for that code
(GradientBalancing identity backward, cloud_transform.py:21-23; clamp mask, :91).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker or the timed CPU baseline.  The product path
(cloud_transformers_b200/) never imports it and fails loudly when the CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the reference's own code run in the build container
(oracle/reference_loader.py + tests/golden/make_golden.py -> tests/golden/*.npz).  The third-party
arithmetic of torch_scatter.scatter_max (not vendored in the reference; install_deps.sh:6, torch-scatter
2.0.x for torch 1.6) is restated from its published CPU semantics: out is not re-initialised, update rule
`if (src > out) { out = src; arg = e }` in ascending e, arg sentinel = src.size(dim).

All float arithmetic is float32 with one rounding per operation, in the reference's operation order.
"""
import numpy as np

F32 = np.float32
EPS = 1e-7  # cloud_transform.py:59


def _sizes(tensor_size, dim):
    """cloud_transform.py:41-46 -- int -> dim*[int], tuple kept."""
    if isinstance(tensor_size, int):
        return [tensor_size] * dim
    assert len(tensor_size) == dim
    return [int(t) for t in tensor_size]


def clamp_bounds():
    """cloud_transform.py:91 -- python doubles -1+eps / 1-eps converted to float32 by torch.clamp."""
    return F32(-1 + EPS), F32(1 - EPS)


def scaled_coords(keys, tensor_size, heads, dim):
    """cloud_transform.py:89-94 -- reshape, clamp, (k + 1) * ((W - 1) * 0.5).

    keys: float32 [B, heads*dim, N]  ->  x float32 [B*heads, dim, N]
    """
    keys = np.asarray(keys, dtype=F32)
    B, HD, N = keys.shape
    assert HD == heads * dim
    W = _sizes(tensor_size, dim)
    lo, hi = clamp_bounds()
    k = keys.reshape(B * heads, dim, N)
    k = np.minimum(np.maximum(k, lo), hi)          # NaN propagates like torch.clamp
    tensor_mod = np.asarray(W, dtype=F32)[None, :, None]
    scale = (tensor_mod - F32(1)) * F32(0.5)
    return (k + F32(1.0)) * scale


def positions_fwd(keys, tensor_size, heads, dim):
    """DifferentiablePositions.forward, cloud_transform.py:72-121.

    returns local_coordinate float32 [B, H, S, N], flattened_index int64 [B, H, S, N]
    Corner order: bit0 of s -> +1 on axis 0 (slowest grid axis), bit1 -> axis 1, bit2 -> axis 2
    (layers/utils.py:103-110, :161-164).
    """
    keys = np.asarray(keys, dtype=F32)
    B, _, N = keys.shape
    W = _sizes(tensor_size, dim)
    S = 1 << dim
    x = scaled_coords(keys, tensor_size, heads, dim)           # [BH, d, N]
    fl = np.floor(x)
    hi_w = (fl + F32(1)) - x                                   # weight of the "floor" corner per axis
    lo_w = x - fl                                              # weight of the "+1" corner per axis
    lc = np.empty((B * heads, S, N), dtype=F32)
    idx = np.empty((B * heads, S, N), dtype=np.int64)
    fli = fl.astype(np.int64)
    for s in range(S):
        w = None
        flat = np.zeros((B * heads, N), dtype=np.int64)
        for a in range(dim):
            bit = (s >> a) & 1
            wa = lo_w[:, a] if bit else hi_w[:, a]
            w = wa if w is None else (w * wa)                  # left-assoc product, utils.py:144-151/179-182
            flat = flat * W[a] + (fli[:, a] + bit)             # x*W1*W2 + y*W2 + z, cloud_transform.py:113-119
        lc[:, s] = w
        idx[:, s] = flat
    return lc.reshape(B, heads, S, N), idx.reshape(B, heads, S, N)


def positions_bwd(keys, grad_lc, tensor_size, heads, dim):
    """Autograd of positions_fwd w.r.t. keys.

    d lc_s / d x_a = (+1 if bit_a(s) else -1) * prod_{a' != a} w_{a'}; floor has zero gradient;
    GradientBalancing.backward is the identity (cloud_transform.py:21-23) so there is NO (W-1)/2 factor;
    clamp passes gradient where lo <= key <= hi (:91).
    returns grad_keys float32 [B, heads*dim, N]
    """
    keys = np.asarray(keys, dtype=F32)
    grad_lc = np.asarray(grad_lc, dtype=F32)
    B, _, N = keys.shape
    S = 1 << dim
    x = scaled_coords(keys, tensor_size, heads, dim).astype(np.float64)
    fl = np.floor(x)
    hi_w = (fl + 1.0) - x
    lo_w = x - fl
    g = grad_lc.reshape(B * heads, S, N).astype(np.float64)
    gx = np.zeros_like(x)
    for s in range(S):
        for a in range(dim):
            sign = 1.0 if (s >> a) & 1 else -1.0
            other = np.ones((B * heads, N))
            for a2 in range(dim):
                if a2 != a:
                    other = other * (lo_w[:, a2] if (s >> a2) & 1 else hi_w[:, a2])
            gx[:, a] += sign * g[:, s] * other
    lo, hi = clamp_bounds()
    k = keys.reshape(B * heads, dim, N)
    mask = (k >= lo) & (k <= hi)
    return (gx * mask).astype(F32).reshape(B, heads * dim, N)


def _pre_splat(lc, features, pad, heads):
    """cloud_transform.py:156-161 -- (features * pad) * lc -> [B, H, F, S, N] float32."""
    lc = np.asarray(lc, dtype=F32)
    features = np.asarray(features, dtype=F32)
    B, H, S, N = lc.shape
    assert H == heads and features.shape[1] % heads == 0
    Fd = features.shape[1] // heads
    f = features.reshape(B, heads, Fd, N)
    if pad is not None:
        f = f * np.asarray(pad, dtype=F32)[:, None, None, :]
    return f[:, :, :, None, :] * lc[:, :, None, :, :]


def scatter_max_first(src, index, C):
    """torch_scatter.scatter_max(out=zeros, dim=-1) CPU semantics (not vendored; see module docstring).

    src   float32 [R, E], index int64 [R, E] (already broadcast), out zero-initialised [R, C].
    returns (out float32 [R, C], arg int64 [R, C]) with arg == E where nothing beat the 0 floor.
    Rule: ascending e, `if src > out: out = src; arg = e`  ==  arg = min{e : src_e == max and max > 0}.
    """
    R, E = src.shape
    rows = np.broadcast_to(np.arange(R)[:, None], (R, E))
    out = np.zeros((R, C), dtype=F32)
    np.maximum.at(out, (rows, index), np.where(np.isnan(src), F32(0), src))
    cand = (src == out[rows, index]) & (src > 0)
    arg = np.full((R, C), E, dtype=np.int64)
    e = np.broadcast_to(np.arange(E, dtype=np.int64)[None, :], (R, E))
    np.minimum.at(arg, (rows[cand], index[cand]), e[cand])
    return out, arg


def scatter_max_loop(src, index, C):
    """Literal loop form of the same rule, for small cases (cross-check of scatter_max_first)."""
    R, E = src.shape
    out = np.zeros((R, C), dtype=F32)
    arg = np.full((R, C), E, dtype=np.int64)
    for r in range(R):
        for e in range(E):
            c = index[r, e]
            if src[r, e] > out[r, c]:
                out[r, c] = src[r, e]
                arg[r, c] = e
    return out, arg


def splat_fwd(lc, idx, features, tensor_size, heads, dim, pad=None, reduce="max", return_arg=False):
    """Splat.forward, cloud_transform.py:131-180.

    reduce="max" is the reference (scatter_max onto zeros => implicit 0 floor);
    reduce="sum" is the north-star scatter-add variant (index_add restatement, float32 accumulate in
    ascending e order per cell).
    returns z float32 [B, H*F, *tensor_size]  (+ arg int64 [B, H, F, C], e = s*N + n, sentinel S*N)
    """
    W = _sizes(tensor_size, dim)
    C = int(np.prod(W))
    pre = _pre_splat(lc, features, pad, heads)                 # [B,H,F,S,N]
    B, H, Fd, S, N = pre.shape
    src = pre.reshape(B * H * Fd, S * N)
    index = np.broadcast_to(np.asarray(idx).reshape(B, H, 1, S * N), (B, H, Fd, S * N)).reshape(B * H * Fd, S * N)
    if reduce == "max":
        out, arg = scatter_max_first(src, index, C)
    else:
        out = np.zeros((B * H * Fd, C), dtype=F32)
        rows = np.broadcast_to(np.arange(B * H * Fd)[:, None], src.shape)
        np.add.at(out, (rows, index), src)
        arg = None
    z = out.reshape(B, H * Fd, *W)
    if return_arg:
        return z, (None if arg is None else arg.reshape(B, H, Fd, C))
    return z


def splat_bwd(lc, idx, features, grad_z, arg, tensor_size, heads, dim, pad=None, reduce="max"):
    """Backward of splat_fwd: ScatterMax backward routes each cell's gradient to its single
    arg winner, then MulBackward (cloud_transform.py:159-161).

    returns grad_features [B, H*F, N], grad_lc [B, H, S, N]   (float64 accumulate, cast to float32)
    """
    lc = np.asarray(lc, dtype=F32)
    features = np.asarray(features, dtype=F32)
    B, H, S, N = lc.shape
    Fd = features.shape[1] // heads
    W = _sizes(tensor_size, dim)
    C = int(np.prod(W))
    gz = np.asarray(grad_z, dtype=np.float64).reshape(B, H, Fd, C)
    index = np.broadcast_to(np.asarray(idx).reshape(B, H, 1, S * N), (B, H, Fd, S * N))
    g_at = np.take_along_axis(gz, index, axis=3)               # grad_z at every (s,n)'s cell
    if reduce == "max":
        a_at = np.take_along_axis(np.asarray(arg).reshape(B, H, Fd, C), index, axis=3)
        e = np.arange(S * N, dtype=np.int64)[None, None, None, :]
        g_pre = np.where(a_at == e, g_at, 0.0)
    else:
        g_pre = g_at
    g_pre = g_pre.reshape(B, H, Fd, S, N)
    f = features.reshape(B, H, Fd, N).astype(np.float64)
    p = np.ones((B, N)) if pad is None else np.asarray(pad, dtype=np.float64)
    fm = f * p[:, None, None, :]
    grad_feat = (g_pre * lc[:, :, None].astype(np.float64)).sum(3) * p[:, None, None, :]
    grad_lc = (g_pre * fm[:, :, :, None, :]).sum(2)
    return grad_feat.reshape(B, H * Fd, N).astype(F32), grad_lc.astype(F32)


def slice_fwd(lc, idx, grid, heads, pad=None):
    """Slice.forward, cloud_transform.py:190-227.  Sum over corners sequential in s, float32."""
    lc = np.asarray(lc, dtype=F32)
    B, H, S, N = lc.shape
    grid = np.asarray(grid, dtype=F32)
    Fd = grid.shape[1] // heads
    g = grid.reshape(B, H, Fd, -1)
    index = np.broadcast_to(np.asarray(idx).reshape(B, H, 1, S * N), (B, H, Fd, S * N))
    gathered = np.take_along_axis(g, index, axis=3).reshape(B, H, Fd, S, N)
    prod = gathered * lc[:, :, None]
    out = prod[:, :, :, 0].copy()
    for s in range(1, S):
        out = out + prod[:, :, :, s]
    out = out.reshape(B, H * Fd, N)
    if pad is not None:
        out = out * np.asarray(pad, dtype=F32)[:, None, :]
    return out


def slice_bwd(lc, idx, grid, grad_out, heads, pad=None):
    """Backward of slice_fwd: gather backward = scatter-add into the grid.

    returns grad_grid (same shape as grid), grad_lc [B, H, S, N]  (float64 accumulate, cast to float32)
    """
    lc = np.asarray(lc, dtype=F32)
    B, H, S, N = lc.shape
    grid = np.asarray(grid, dtype=F32)
    Fd = grid.shape[1] // heads
    C = int(np.prod(grid.shape[2:]))
    g = grid.reshape(B, H, Fd, C).astype(np.float64)
    go = np.asarray(grad_out, dtype=np.float64).reshape(B, H, Fd, N)
    if pad is not None:
        go = go * np.asarray(pad, dtype=np.float64)[:, None, None, :]
    index = np.broadcast_to(np.asarray(idx).reshape(B, H, 1, S * N), (B, H, Fd, S * N))
    gathered = np.take_along_axis(g, index, axis=3).reshape(B, H, Fd, S, N)
    grad_lc = (gathered * go[:, :, :, None, :]).sum(2)
    contrib = (lc[:, :, None].astype(np.float64) * go[:, :, :, None, :]).reshape(B * H * Fd, S * N)
    gg = np.zeros((B * H * Fd, C))
    rows = np.broadcast_to(np.arange(B * H * Fd)[:, None], contrib.shape)
    np.add.at(gg, (rows, index.reshape(B * H * Fd, S * N)), contrib)
    return gg.reshape(grid.shape).astype(F32), grad_lc.astype(F32)


def so3_exponential_map(log_rot, eps=1e-4):
    """pytorch3d.transforms.so3.so3_exponential_map (not vendored; install_deps.sh:10, ~v0.2.5):
    Rodrigues with the squared angle clamped to >= eps.  log_rot [H, 3] -> R [H, 3, 3] (float64 math)."""
    v = np.asarray(log_rot, dtype=np.float64)
    nrms = (v * v).sum(1)
    theta = np.sqrt(np.maximum(nrms, eps))
    fac1 = np.sin(theta) / theta
    fac2 = (1.0 - np.cos(theta)) / (theta * theta)
    K = np.zeros((v.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -v[:, 2], v[:, 1]
    K[:, 1, 0], K[:, 1, 2] = v[:, 2], -v[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -v[:, 1], v[:, 0]
    R = fac1[:, None, None] * K + fac2[:, None, None] * (K @ K) + np.eye(3)[None]
    return R


def algorithmic_bytes(N, d, F, W, e=4, reduce="max"):
    """minimal HBM bytes per (batch, head) unit for Splat fwd + Slice fwd +
    Slice bwd + Splat bwd."""
    C = int(np.prod(_sizes(W, d)))
    grid_passes = 6 if reduce == "max" else 5
    return N * (24 * d + 5 * e * F) + grid_passes * e * F * C
