"""Completion-loss kernels (row N3): this library against the reference's own CUDA kernels (oracle/_ref/cuda_ext, test
infrastructure -- which is why this script lives under tests/) at the shapes the training scripts use.
    python tests/bench_losses_vs_ref.py          # not collected by pytest"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloud_transformers_b200.chamfer import ChamferFunction  # noqa: E402
from cloud_transformers_b200.emd import emdModule  # noqa: E402
from oracle import build_ref_cuda as R  # noqa: E402
from tests.test_losses_ref_gpu import ref_emd_forward  # noqa: E402

DEV = "cuda:0"


def timeit(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g = torch.Generator(device=DEV).manual_seed(0)
ref_emd, ref_ch = R.load("ref_emd"), R.load("ref_chamfer")
print("EMD: B x n, eps, iters | ours ms | reference kernels ms | loss ours / reference")
for B, n, eps, iters, reps in [(32, 2048, 0.005, 50, 10), (8, 2048, 0.005, 50, 10), (32, 2048, 0.004, 3000, 2), (8, 8192, 0.005, 50, 3), (2, 16384, 0.005, 50, 3), (2, 16384, 0.004, 3000, 1)]:
    a = torch.rand(B, n, 3, device=DEV, generator=g)
    b = (a[:, torch.randperm(n, device=DEV, generator=g)] + 0.02 * torch.randn(B, n, 3, device=DEV, generator=g)).clamp(0, 1).contiguous()
    t = timeit(lambda: emdModule()(a, b, eps, iters), reps)
    lo = torch.sqrt(emdModule()(a, b, eps, iters)[0]).mean().item()
    if ref_emd is not None:
        tr = timeit(lambda: ref_emd_forward(ref_emd, a, b, eps, iters), reps)
        lr = torch.sqrt(ref_emd_forward(ref_emd, a, b, eps, iters)[0]).mean().item()
    else:
        tr, lr = float("nan"), float("nan")
    print("%3d x %4d, %.3f, %4d | %8.3f | %8.3f | %.5f / %.5f" % (B, n, eps, iters, t, tr, lo, lr), flush=True)
print("Chamfer forward: B x n x m | ours ms | reference kernels ms")
for B, n, m in [(32, 2048, 2048), (8, 8192, 8192), (32, 2048, 16384)]:
    a, b = torch.rand(B, n, 3, device=DEV, generator=g), torch.rand(B, m, 3, device=DEV, generator=g)
    t = timeit(lambda: ChamferFunction.apply(a, b), 10)
    if ref_ch is not None:
        r1, r2 = torch.zeros(B, n, device=DEV), torch.zeros(B, m, device=DEV)
        i1, i2 = torch.zeros(B, n, dtype=torch.int32, device=DEV), torch.zeros(B, m, dtype=torch.int32, device=DEV)
        tr = timeit(lambda: ref_ch.forward(a, b, r1, r2, i1, i2), 10)
    else:
        tr = float("nan")
    print("%3d x %5d x %5d | %8.3f | %8.3f" % (B, n, m, t, tr), flush=True)
