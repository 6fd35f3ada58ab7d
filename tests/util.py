"""Shared helpers for the test suite: seeded synthetic inputs and tolerance rules."""
import numpy as np

GOLDEN_CASES = {
    # name: (dim, tensor_size, heads, F, N, B)
    "c2d_w8": (2, 8, 2, 3, 40, 2),
    "c2d_w16_pad": (2, 16, 3, 4, 64, 2),
    "c3d_w4": (3, 4, 2, 2, 33, 2),
    "c3d_w8_pad": (3, 8, 2, 4, 48, 1),
    "c2d_rect": (2, (6, 10), 2, 2, 37, 1),
    "c3d_rect": (3, (4, 6, 5), 2, 3, 29, 2),
}

# fp32 tolerance of the north star: rel 1e-5, with an absolute floor tied to the magnitude of the tensor
RTOL = 1e-5


def assert_close(a, b, what="", rtol=RTOL, atol_scale=1e-5):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    atol = atol_scale * max(1e-30, float(np.abs(b).max()))
    err = np.abs(a - b)
    bad = err > atol + rtol * np.abs(b)
    assert not bad.any(), "%s: %d / %d outside tol, max abs err %.3e (max |ref| %.3e)" % (
        what, int(bad.sum()), bad.size, float(err.max()), float(np.abs(b).max()))


def make_inputs(seed, B, H, dim, F, N, pad=False, dist="tanh"):
    rng = np.random.default_rng(seed)
    if dist == "tanh":
        keys = np.tanh(rng.standard_normal((B, H * dim, N)) * 1.2)
    elif dist == "uniform":
        keys = rng.uniform(-1, 1, (B, H * dim, N))
    elif dist == "onecell":
        keys = np.full((B, H * dim, N), 0.1234) + rng.uniform(0, 1e-4, (B, H * dim, N))
    else:
        raise ValueError(dist)
    feat = rng.standard_normal((B, H * F, N))
    p = (rng.uniform(size=(B, N)) > 0.2).astype(np.float32) if pad else None
    return keys.astype(np.float32), feat.astype(np.float32), p
