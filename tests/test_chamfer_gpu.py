"""Chamfer nearest-neighbour distances (row N3) on the GPU against the numpy restatement of the reference's kernels
(oracle/chamfer_numpy.py: chamfer_extension/chamfer.cu:12-174).  Indices bit-exact, distances and gradients rel 1e-5."""
import numpy as np
import pytest
import torch

from cloud_transformers_b200 import chamfer as C
from oracle import chamfer_numpy as O
from tests.util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def clouds(seed, B, n, m):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-1, 1, (B, n, 3)).astype(np.float32), rng.uniform(-1, 1, (B, m, 3)).astype(np.float32))


@pytest.mark.parametrize("B,n,m", [(2, 100, 77), (1, 1024, 1024), (3, 2048, 513), (2, 5000, 16384), (1, 1, 9)])
def test_chamfer_forward_backward_against_oracle(B, n, m):
    a, b = clouds(B * 1000 + n, B, n, m)
    ta, tb = torch.from_numpy(a).to(DEV).requires_grad_(True), torch.from_numpy(b).to(DEV).requires_grad_(True)
    d1, d2, i1, i2 = C.ChamferFunction.apply(ta, tb)
    od1, od2, oi1, oi2 = O.forward(a, b)
    assert np.array_equal(i1.cpu().numpy(), oi1) and np.array_equal(i2.cpu().numpy(), oi2)
    assert_close(d1.detach().cpu().numpy(), od1, "dist1")
    assert_close(d2.detach().cpu().numpy(), od2, "dist2")
    rng = np.random.default_rng(1)
    g1, g2 = rng.standard_normal(od1.shape).astype(np.float32), rng.standard_normal(od2.shape).astype(np.float32)
    (d1 * torch.from_numpy(g1).to(DEV)).sum().backward(retain_graph=True)
    (d2 * torch.from_numpy(g2).to(DEV)).sum().backward()
    og1, og2 = O.backward(a, b, g1, g2, oi1, oi2)
    assert_close(ta.grad.cpu().numpy(), og1, "grad xyz1")
    assert_close(tb.grad.cpu().numpy(), og2, "grad xyz2")


def test_chamfer_ties_take_the_first_minimum():
    """duplicated targets: the reference keeps the first index (strict < in ascending k, chamfer.cu:36-40), whatever
    the split of the target range over CTAs"""
    a, b = clouds(5, 2, 3000, 4096)
    b[:, 2048:] = b[:, :2048]
    d1, d2, i1, i2 = C.ChamferFunction.apply(torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV))
    assert int(i1.max()) < 2048
    od1, od2, oi1, oi2 = O.forward(a, b)
    assert np.array_equal(i1.cpu().numpy(), oi1)


def test_loss_functions_match_reference_formulas():
    a, b = clouds(9, 2, 512, 700)
    pa = torch.from_numpy(a).to(DEV).permute(0, 2, 1)[:, :, None].contiguous()       # [B,3,1,N], dist_chamfer.py:67
    pb = torch.from_numpy(b).to(DEV).permute(0, 2, 1)[:, :, None].contiguous()
    od1, od2, _, _ = O.forward(a, b)
    assert abs(float(C.loss_chamfer(pa, pb)) - (od1.mean() + od2.mean())) < 1e-5
    assert abs(float(C.loss_chamfer_adj(pa, pb)) - (np.sqrt(od1).mean() + np.sqrt(od2).mean()) / 2) < 1e-5
    a2, b2 = a.copy(), b.copy()
    a2[..., 2] = 0
    b2[..., 2] = 0
    e1, e2, _, _ = O.forward(a2, b2)
    assert abs(float(C.loss_chamder_2d(pa[:, :2], pb[:, :2])) - (e1.mean() + e2.mean())) < 1e-5
