"""GPU suite: the two loss kernels of row N3 against the REFERENCE'S OWN CUDA kernels, compiled for sm_100a from
/root/reference/{emd_linear,chamfer_extension} by oracle/build_ref_cuda.py into oracle/_ref/cuda_ext/ (the .so files
travel to the GPU box; the sources are never copied).  Skipped where the prebuilt files are absent.

Chamfer: distances agree to 1 ulp-level tolerance (the reference's nvcc contracts the squared distance into FMAs),
indices wherever the two nearest candidates are not within that tolerance.
EMD: the reference resolves bids within 1e-6 of the maximum by a write race, so two runs of the reference itself can
differ; compared are the quantities the training scripts use -- the loss sqrt(dist).mean() -- and the quality of the
assignment (distinct targets), plus exact equality where the algorithm leaves no choice (a single iteration)."""
import numpy as np
import pytest
import torch

from oracle import build_ref_cuda as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref(name):
    mod = R.load(name)
    if mod is None:
        pytest.skip("oracle/_ref/cuda_ext/%s.so not built (python oracle/build_ref_cuda.py where /root/reference exists)" % name)
    return mod


def ref_emd_forward(mod, xyz1, xyz2, eps, iters):
    """the call of emd_linear/emd_module.py:30-63"""
    B, n, _ = xyz1.shape
    z = lambda *s, **k: torch.zeros(*s, device=DEV, **k)
    dist = z(B, n)
    assignment = z(B, n, dtype=torch.int32) - 1
    assignment_inv = z(B, n, dtype=torch.int32) - 1
    price, bid, bid_inc, max_inc = z(B, n), z(B, n, dtype=torch.int32), z(B, n), z(B, n)
    unass_idx, max_idx = z(B * n, dtype=torch.int32), z(B * n, dtype=torch.int32)
    unass_cnt, unass_cnt_sum, cnt_tmp = z(512, dtype=torch.int32), z(512, dtype=torch.int32), z(512, dtype=torch.int32)
    mod.forward(xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_inc, max_inc, unass_idx, unass_cnt,
                unass_cnt_sum, cnt_tmp, max_idx, eps, iters)
    torch.cuda.synchronize()
    return dist, assignment


def test_emd_against_the_reference_kernels():
    from cloud_transformers_b200.emd import emdModule
    mod = _ref("ref_emd")
    g = torch.Generator(device=DEV).manual_seed(1)
    for B, n, eps, iters in [(4, 1024, 0.005, 50), (2, 2048, 0.005, 50), (2, 2048, 0.004, 3000), (2, 16384, 0.005, 50)]:
        a = torch.rand(B, n, 3, device=DEV, generator=g)
        b = (a[:, torch.randperm(n, device=DEV, generator=g)] + 0.02 * torch.randn(B, n, 3, device=DEV, generator=g)).clamp(0, 1)
        d_ref, a_ref = ref_emd_forward(mod, a, b.contiguous(), eps, iters)
        d, asg = emdModule()(a, b.contiguous(), eps, iters)
        loss_ref, loss = torch.sqrt(d_ref).mean(1), torch.sqrt(d).mean(1)
        assert torch.allclose(loss, loss_ref, rtol=2e-2), (loss.tolist(), loss_ref.tolist())
        for i in range(B):
            u_ref, u = a_ref[i].unique().numel(), asg[i].unique().numel()
            assert u >= u_ref - max(4, n // 100), (u, u_ref)
        # the distances belong to the assignment, as CalcDist computes them
        picked = torch.gather(b, 1, asg.long()[..., None].expand(-1, -1, 3))
        assert torch.allclose(d, ((a - picked) ** 2).sum(-1), rtol=1e-5, atol=1e-9)
    # one iteration: every source takes its best target at zero prices -- no race in the reference either
    a, b = torch.rand(2, 1024, 3, device=DEV, generator=g), torch.rand(2, 1024, 3, device=DEV, generator=g)
    d_ref, a_ref = ref_emd_forward(mod, a, b, 0.005, 1)
    d, asg = emdModule()(a, b, 0.005, 1)
    same = (asg == a_ref).float().mean().item()
    assert same >= 0.999, same                       # (FMA contraction can swap two targets that tie to the last bit)
    assert torch.allclose(d, d_ref, rtol=1e-5, atol=1e-9)


def test_emd_backward_against_the_reference_kernel():
    from cloud_transformers_b200 import _lib
    from cloud_transformers_b200.functional import _call, _ptr, _stream
    mod = _ref("ref_emd")
    g = torch.Generator(device=DEV).manual_seed(2)
    a, b = torch.rand(2, 1024, 3, device=DEV, generator=g), torch.rand(2, 1024, 3, device=DEV, generator=g)
    gd = torch.randn(2, 1024, device=DEV, generator=g)
    asg = torch.randint(0, 1024, (2, 1024), device=DEV, generator=g, dtype=torch.int32)
    ref = torch.zeros_like(a)
    mod.backward(a, b, ref, gd, asg)
    ours = torch.empty_like(a)
    _call("ctb_emd_bwd", _ptr(a), _ptr(b), _ptr(gd), _ptr(asg), _ptr(ours), 2, 1024, _stream(a))
    torch.cuda.synchronize()
    assert torch.allclose(ours, ref, rtol=1e-6, atol=1e-8)


def test_chamfer_against_the_reference_kernels():
    from cloud_transformers_b200.chamfer import ChamferFunction
    mod = _ref("ref_chamfer")
    g = torch.Generator(device=DEV).manual_seed(3)
    for B, n, m in [(4, 2048, 2048), (2, 1000, 3000), (1, 8192, 2048)]:
        a, b = torch.rand(B, n, 3, device=DEV, generator=g), torch.rand(B, m, 3, device=DEV, generator=g)
        r1, r2 = torch.zeros(B, n, device=DEV), torch.zeros(B, m, device=DEV)
        i1, i2 = torch.zeros(B, n, dtype=torch.int32, device=DEV), torch.zeros(B, m, dtype=torch.int32, device=DEV)
        mod.forward(a, b, r1, r2, i1, i2)                      # the call of chamfer_extension/dist_chamfer.py:26-27
        torch.cuda.synchronize()
        d1, d2, j1, j2 = ChamferFunction.apply(a, b)
        assert torch.allclose(d1, r1, rtol=1e-5, atol=1e-9) and torch.allclose(d2, r2, rtol=1e-5, atol=1e-9)
        assert (j1 == i1).float().mean().item() >= 0.999 and (j2 == i2).float().mean().item() >= 0.999
        g1, g2 = torch.randn(B, n, device=DEV, generator=g), torch.randn(B, m, device=DEV, generator=g)
        ra, rb = torch.zeros_like(a), torch.zeros_like(b)
        mod.backward(a, b, ra, rb, g1, g2, i1, i2)             # dist_chamfer.py:50-52
        torch.cuda.synchronize()
        from cloud_transformers_b200.functional import _call, _ptr, _stream
        oa, ob = torch.empty_like(a), torch.empty_like(b)
        _call("ctb_chamfer_bwd", _ptr(a), _ptr(b), _ptr(g1), _ptr(g2), _ptr(i1), _ptr(i2), _ptr(oa), _ptr(ob), B, n, m, _stream(a))
        torch.cuda.synchronize()
        assert torch.allclose(oa, ra, rtol=1e-5, atol=1e-7) and torch.allclose(ob, rb, rtol=1e-4, atol=1e-6)
