"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's auction EMD, emd_linear/emd_cuda.cu, one cloud at a
time in float32 numpy.  Only tests/ may import this.

Follows the reference step by step -- unassigned list (:30-93), Bid (:95-173), GetMax (:175-188), Assign (:190-210),
CalcDist (:212-221) -- with the two choices the reference leaves to thread timing made deterministic, as the B200
kernel makes them: among bidders for one target the HIGHEST increment wins and exact ties go to the smallest source
index (the reference lets any bidder within 1e-6 of the maximum win by a write race, :181-185), and a source's best
target is the FIRST maximum in ascending target index.  Arithmetic: squared distance with every product and sum rounded
to float32 (the reference's nvcc contracts them into FMAs: <= 1 ulp apart), value = float32(3.0 - float64(sqrt) -
float64(price)) as the double literal in :131 makes it.

Parity: pinned on the GPU against the reference's own compiled kernels (oracle/_ref/cuda_ext/ref_emd.so, built by
oracle/build_ref_cuda.py from /root/reference/emd_linear) in tests/test_losses_ref_gpu.py."""
import numpy as np

F32 = np.float32


def emd_auction(xyz1, xyz2, eps, iters):
    """xyz1, xyz2: [n, 3] float32.  Returns (dist [n] float32, assignment [n] int32)."""
    x1 = np.asarray(xyz1, dtype=F32)
    x2 = np.asarray(xyz2, dtype=F32)
    n = x1.shape[0]
    eps = F32(eps)
    price = np.zeros(n, dtype=F32)
    asg = np.full(n, -1, dtype=np.int64)
    inv = np.full(n, -1, dtype=np.int64)
    for it in range(iters):
        last = it == iters - 1
        una = np.nonzero(asg == -1)[0]
        if una.size == 0:
            break
        best_i = np.empty(una.size, dtype=np.int64)
        best = np.empty(una.size, dtype=F32)
        better = np.empty(una.size, dtype=F32)
        for c0 in range(0, una.size, 512):                             # (blocks of sources: memory)
            u = una[c0:c0 + 512]
            d = x2[None, :, :] - x1[u, None, :]                        # [u, n, 3]
            sq = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
            val = (3.0 - np.sqrt(sq).astype(np.float64) - price[None, :].astype(np.float64)).astype(F32)
            bi = val.argmax(axis=1)                                    # first maximum
            rows = np.arange(u.size)
            best_i[c0:c0 + 512] = bi
            best[c0:c0 + 512] = val[rows, bi]
            val[rows, bi] = -np.inf
            better[c0:c0 + 512] = np.maximum(val.max(axis=1), F32(-1e9)) if n > 1 else F32(-1e9)
        inc = ((best - better).astype(F32) + eps).astype(F32)
        if last:
            asg[una] = best_i
            break
        # per target: highest increment, then smallest source index
        order = np.lexsort((una, -inc.astype(np.float64), best_i))
        tgt_sorted = best_i[order]
        first = np.ones(order.size, dtype=bool)
        first[1:] = tgt_sorted[1:] != tgt_sorted[:-1]
        for o in order[first]:
            j, t = una[o], best_i[o]
            if inv[t] != -1:
                asg[inv[t]] = -1
            inv[t] = j
            asg[j] = t
            price[t] = F32(price[t] + inc[o])
    diff = x1 - x2[asg]
    dist = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
    return dist.astype(F32), asg.astype(np.int32)


def emd_grad(xyz1, xyz2, grad_dist, assignment):
    """emd_cuda.cu:277-300: grad_xyz1[j] = 2 g[j] (p1[j] - p2[a[j]]); no gradient for xyz2."""
    x1 = np.asarray(xyz1, dtype=F32)
    x2 = np.asarray(xyz2, dtype=F32)
    g = (np.asarray(grad_dist, dtype=F32) * F32(2.0)).astype(F32)
    return (g[:, None] * (x1 - x2[assignment])).astype(F32)
