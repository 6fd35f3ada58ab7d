"""CPU restatement of the reference's Chamfer nearest-neighbour kernels -- TEST INFRASTRUCTURE ONLY.

Follows chamfer_extension/chamfer.cu:12-134 (NmDistanceKernel: d = dx*dx + dy*dy + dz*dz in float32, first strict
minimum in ascending k) and :155-174 (NmDistanceGradKernel).  The reference ships no tests for it
and its extension cannot be built without a GPU toolchain run, so parity is pinned against this restatement: indices
bit-exact on inputs without float32 near-ties, distances / gradients within rel 1e-5 (the reference's nvcc build may or
may not contract the sum of squares into FMAs)."""
import numpy as np


def nn_distance(a, b):
    """a [B,n,3], b [B,m,3] float32 -> dist [B,n] float32, idx [B,n] int32 (first minimum)"""
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    d = a32[:, :, None, :] - b32[:, None, :, :]                       # float32 differences like the kernel
    dist = (d.astype(np.float64) ** 2).sum(-1)                        # exact sum of the float32 squares
    idx = dist.argmin(-1).astype(np.int32)                            # numpy argmin = first minimum
    return np.take_along_axis(dist, idx[..., None].astype(np.int64), -1)[..., 0].astype(np.float32), idx


def forward(xyz1, xyz2):
    d1, i1 = nn_distance(xyz1, xyz2)
    d2, i2 = nn_distance(xyz2, xyz1)
    return d1, d2, i1, i2


def backward(xyz1, xyz2, gd1, gd2, idx1, idx2):
    g1 = np.zeros(xyz1.shape, dtype=np.float64)
    g2 = np.zeros(xyz2.shape, dtype=np.float64)
    B = xyz1.shape[0]
    for b in range(B):
        diff = xyz1[b].astype(np.float64) - xyz2[b][idx1[b]].astype(np.float64)
        v = 2.0 * gd1[b][:, None].astype(np.float64) * diff
        g1[b] += v
        np.add.at(g2[b], idx1[b], -v)
        diff = xyz2[b].astype(np.float64) - xyz1[b][idx2[b]].astype(np.float64)
        v = 2.0 * gd2[b][:, None].astype(np.float64) * diff
        g2[b] += v
        np.add.at(g1[b], idx2[b], -v)
    return g1.astype(np.float32), g2.astype(np.float32)
