"""Torch-CPU port of the reference's Splat / Slice composition -- TEST INFRASTRUCTURE ONLY.

This is the "reference's CPU path" that bench.py times beside the GPU numbers (`cpu_baseline`, kind
"port") and under `--impl reference`: the reference itself is a composition of torch ops plus
torch_scatter.scatter_max, it cannot travel to the GPU box (/root/reference is absent there), so this
port issues the same sequence of torch ops -- materialised pre_splat, zero-filled grid, scatter-max,
expanded int64 index, gather, multiply, sum -- with autograd providing the backward exactly as it does
for the reference (layers/cloud_transform.py:72-121, :131-180, :190-227; layers/utils.py:100-186).
scatter_max is the first-winner shim from oracle/reference_loader.py (torch-scatter is not installed).

Validated against the real reference in tests/test_oracle.py (build container only).
Never imported by the product path.
"""
import torch

from .reference_loader import _ScatterMaxFirst

EPS = 1e-7


class _Balance(torch.autograd.Function):
    """cloud_transform.py:12-23 -- forward multiplies by scale, backward is the identity."""

    @staticmethod
    def forward(ctx, x, scale):
        return x * scale

    @staticmethod
    def backward(ctx, g):
        return g, None


def _sizes(tensor_size, dim):
    return [tensor_size] * dim if isinstance(tensor_size, int) else [int(t) for t in tensor_size]


def positions(keys, tensor_size, heads, dim):
    """cloud_transform.py:72-121 + utils.py:100-186 -> (lc [B,H,S,N] f32, idx [B,H,S,N] i64)."""
    B, _, N = keys.shape
    W = _sizes(tensor_size, dim)
    S = 1 << dim
    mod = torch.tensor(W, dtype=torch.float32, device=keys.device)[None, :, None]
    k = keys.reshape(B * heads, dim, N).clamp(-1 + EPS, 1 - EPS)
    x = _Balance.apply(k + 1.0, (mod - 1) * 0.5)
    fl = x.floor()
    up = (fl + 1) - x
    dn = x - fl
    fli = fl.long()
    ws, flats = [], []
    for s in range(S):
        w, flat = None, 0
        for a in range(dim):
            bit = (s >> a) & 1
            wa = dn[:, a] if bit else up[:, a]
            w = wa if w is None else w * wa
            flat = flat * W[a] + (fli[:, a] + bit)
        ws.append(w)
        flats.append(flat)
    lc = torch.stack(ws, dim=1).reshape(B, heads, S, N)
    idx = torch.stack(flats, dim=1).reshape(B, heads, S, N)
    return lc, idx


def splat(lc, idx, features, tensor_size, heads, dim, pad=None):
    """cloud_transform.py:131-180."""
    W = _sizes(tensor_size, dim)
    B, N = features.size(0), features.size(-1)
    Fd = features.size(1) // heads
    f = features.reshape(B, heads, Fd, N)
    if pad is not None:
        f = f * pad[:, None, None, :]
    pre = f[:, :, :, None] * lc[:, :, None]
    C = 1
    for w in W:
        C *= w
    z0 = torch.zeros((B, heads, Fd, C), device=pre.device)
    z, _ = _ScatterMaxFirst.apply(pre.reshape(B, heads, Fd, -1), idx.reshape(B, heads, 1, -1), z0)
    return z.reshape(B, heads * Fd, *W)


def slice_(lc, idx, grid, heads, dim, pad=None):
    """cloud_transform.py:190-227."""
    B, H, S, N = lc.shape
    Fd = grid.size(1) // heads
    ind = idx[:, :, None].expand(-1, -1, Fd, -1, -1).reshape(B, heads, Fd, -1)
    g = torch.gather(grid.reshape(B, heads, Fd, -1), 3, ind).reshape(B, heads, Fd, S, N)
    out = (g * lc[:, :, None]).sum(dim=3).reshape(B, heads * Fd, N)
    if pad is not None:
        out = out * pad[:, None, :]
    return out


def hot_path_fwd_bwd(keys, features, tensor_size, heads, dim, grad_out=None, pad=None):
    """One pass of the hot path: positions + Splat fwd, Slice fwd (on the splatted grid, the conv is
    outside the metric), then backward of both.  Returns (z, out, grad_keys, grad_features)."""
    keys = keys.detach().requires_grad_(True)
    features = features.detach().requires_grad_(True)
    lc, idx = positions(keys, tensor_size, heads, dim)
    z = splat(lc, idx, features, tensor_size, heads, dim, pad)
    out = slice_(lc, idx, z, heads, dim, pad)
    if grad_out is None:
        grad_out = torch.ones_like(out)
    out.backward(grad_out)
    return z.detach(), out.detach(), keys.grad, features.grad
