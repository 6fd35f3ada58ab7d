"""GPU suite: the auction EMD (csrc/ctb_emd.cuh through ctb_emd_fwd / ctb_emd_bwd and the emd_module mirror) against
the oracle, bit for bit: same assignment, same squared distances (integer / index work and a fixed evaluation order)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import emd_numpy as E

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def clouds(seed, B, n, clustered=False):
    rng = np.random.default_rng(seed)
    a = rng.random((B, n, 3)).astype(np.float32)
    b = rng.random((B, n, 3)).astype(np.float32)
    if clustered:                       # many near-equal values: reconstruction vs ground truth of the same surface
        b = np.clip(a[:, rng.permutation(n)] + rng.normal(0, 0.01, (B, n, 3)).astype(np.float32), 0, 1).astype(np.float32)
    return a, b


@pytest.mark.parametrize("B,n,eps,iters,clustered", [
    (2, 1024, 0.005, 50, False), (1, 2048, 0.005, 50, True), (3, 300, 0.01, 7, False), (1, 1000, 0.004, 400, True),
    (2, 64, 0.005, 1, False), (1, 37, 1e-4, 3000, False), (1, 4096, 0.005, 20, True),
    (2, 9000, 0.005, 12, True)])      # above 4096 points: the state moves from shared memory to the global workspace
def test_auction_matches_the_oracle_bit_for_bit(B, n, eps, iters, clustered):
    from cloud_transformers_b200.emd import emdModule
    a, b = clouds(n + iters, B, n, clustered)
    dist, asg = emdModule()(torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV), eps, iters)
    dist, asg = dist.cpu().numpy(), asg.cpu().numpy()
    for i in range(B):
        d0, a0 = E.emd_auction(a[i], b[i], eps, iters)
        assert np.array_equal(asg[i], a0), "cloud %d: %d assignments differ" % (i, int((asg[i] != a0).sum()))
        assert np.array_equal(dist[i], d0)


def test_gradient_and_no_gradient_for_the_target_cloud():
    from cloud_transformers_b200.emd import emdFunction
    a, b = clouds(9, 2, 512)
    ta = torch.from_numpy(a).to(DEV).requires_grad_(True)
    tb = torch.from_numpy(b).to(DEV).requires_grad_(True)
    dist, asg = emdFunction.apply(ta, tb, 0.005, 50)
    g = torch.randn_like(dist)
    (dist * g).sum().backward()
    for i in range(2):
        assert np.array_equal(ta.grad[i].cpu().numpy(), E.emd_grad(a[i], b[i], g[i].cpu().numpy(), asg[i].cpu().numpy()))
    assert float(tb.grad.abs().max()) == 0.0                    # emd_module.py:73-80
    # the loss of train_inpainter.py:187-189
    loss = torch.sqrt(dist).mean(1).mean()
    assert torch.isfinite(loss)


def test_limits_and_errors():
    from cloud_transformers_b200 import _lib
    from cloud_transformers_b200.emd import emdModule
    lib = _lib.load()
    assert lib.ctb_emd_max_points() == 32768
    assert lib.ctb_emd_workspace_bytes(4, 4096) == 0 and lib.ctb_emd_workspace_bytes(4, 16384) >= 4 * 28 * 16384
    for n in (4096, 8192, 16384):               # the largest shared-memory cloud; a workspace cloud; the inpainting decoder's size
        a = torch.rand(2, n, 3, device=DEV)
        dist, asg = emdModule()(a, a.flip(1).contiguous(), 0.005, 50)
        assert int(asg.min()) >= 0 and int(asg.max()) < n
        assert (asg == torch.arange(n - 1, -1, -1, device=DEV, dtype=torch.int32)).all()      # the clouds are permutations
        assert float(dist.max()) == 0.0
    with pytest.raises(_lib.CtbError):
        emdModule()(torch.rand(1, 33000, 3, device=DEV), torch.rand(1, 33000, 3, device=DEV), 0.005, 5)
    with pytest.raises(Exception):
        emdModule()(torch.rand(1, 16, 3), torch.rand(1, 16, 3), 0.005, 5)          # CPU tensors: no fallback


def test_dropin_module_resolves_to_this_library():
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        sys.modules.pop("emd_linear", None)
        sys.modules.pop("emd_linear.emd_module", None)
        import emd_linear.emd_module as emd
        from cloud_transformers_b200.emd import emdModule
        assert emd.emdModule is emdModule
        d, a = emd.emdModule()(torch.rand(2, 128, 3, device=DEV), torch.rand(2, 128, 3, device=DEV), 0.005, 50)
        assert d.shape == (2, 128) and a.dtype == torch.int32
    finally:
        sys.path.remove(os.path.join(ROOT, "dropin"))
