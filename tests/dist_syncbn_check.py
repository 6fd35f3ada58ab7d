"""Two-or-more-GPU check of cloud_transformers_b200.syncbn against torch.nn.SyncBatchNorm (NCCL), run under torchrun:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_syncbn_check.py

Every rank feeds a different batch through the same layers wrapped both ways; outputs, input gradients, parameter
gradients and running statistics must agree to rel 5e-5 of the tensor maximum (two fp32 implementations with different
summation orders, four normalisations deep) over several steps (also under CUDA-graph replay).  Launched by
tests/test_syncbn_gpu.py when the box has >= 2 GPUs."""
import copy
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloud_transformers_b200 import syncbn  # noqa: E402


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.c1 = nn.Conv1d(6, 48, 1, bias=False)
        self.b1 = nn.BatchNorm1d(48)
        self.c2 = nn.Conv1d(48, 20, 1, bias=False)
        self.b2 = nn.BatchNorm1d(20)
        self.b3 = nn.BatchNorm3d(5)
        self.b2d = nn.BatchNorm2d(4)
        self.lin = nn.Linear(20, 33)
        self.b4 = nn.BatchNorm1d(33)

    def forward(self, x):
        y = torch.relu_(self.b1(self.c1(x)))
        y = self.b2(self.c2(y))                                   # [B, 20, L]
        B, _, L = y.shape
        v = self.b3(y.reshape(B, 5, 2, 2, L)).reshape(B, 20, L)
        w = self.b2d(y.reshape(B, 4, 5, L)).reshape(B, 20, L)
        z = self.b4(self.lin((v + w).mean(-1)))       # BatchNorm1d on [B, C]
        return z, v


def close(a, b, what, rel=5e-5, floor=1e-12):
    scale = max(float(b.abs().max()), floor)
    err = float((a - b).abs().max()) / scale
    assert err <= rel, "%s: %.3e > %.1e" % (what, err, rel)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    # TF32 convolutions round their inputs to 10 bits: a 2e-7 difference between the two normalisations flips roundings
    # and shows up as 1e-4 downstream, which is not what this check is about
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    base = Net().to(dev)
    ref = nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(base))
    ours = syncbn.convert_sync_batchnorm(copy.deepcopy(base))
    assert sum(isinstance(m, syncbn.CtbSyncBatchNorm) for m in ours.modules()) == 5
    assert sorted(ref.state_dict().keys()) == sorted(ours.state_dict().keys())
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for step in range(4):
        x = torch.randn(8, 6, 300, device=dev, generator=g) * (1 + rank) + rank
        gw = torch.randn(8, 33, device=dev, generator=g)
        outs = []
        for net in (ref, ours):
            xi = x.clone().requires_grad_(True)
            for p in net.parameters():
                p.grad = None
            z, v = net(xi)
            ((z * gw).sum() + v.square().mean()).backward()
            outs.append((z.detach(), v.detach(), xi.grad, {k: p.grad for k, p in net.named_parameters()},
                         {k: b.clone() for k, b in net.named_buffers()}))
        (z0, v0, gx0, gp0, bf0), (z1, v1, gx1, gp1, bf1) = outs
        close(z1, z0, "step %d output" % step)
        close(v1, v0, "step %d 3-D branch" % step)
        close(gx1, gx0, "step %d input gradient" % step, rel=1e-4)
        for k in gp0:
            close(gp1[k], gp0[k], "step %d grad %s" % (step, k), rel=2e-4)
        for k in bf0:
            if bf0[k].dtype.is_floating_point:
                close(bf1[k], bf0[k], "step %d buffer %s" % (step, k), floor=1e-2)   # b3's running mean is ~0
            else:
                assert torch.equal(bf1[k], bf0[k]), k
    # a layer big enough for several statistics CTAs per channel (chunk partials folded by the last one), an odd L
    for shape in ((8, 16, 64, 64), (8, 3, 7, 9, 11)):
        big = nn.Sequential(nn.BatchNorm2d(shape[1]) if len(shape) == 4 else nn.BatchNorm3d(shape[1])).to(dev)
        with torch.no_grad():
            big[0].weight.uniform_(0.5, 1.5), big[0].bias.uniform_(-1, 1)
        nets = (nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(big)), syncbn.convert_sync_batchnorm(copy.deepcopy(big)))
        xb = torch.randn(*shape, device=dev, generator=g) * (1 + rank) + 3 * rank
        gb = torch.randn(*shape, device=dev, generator=g)
        res = []
        for net in nets:
            xi = xb.clone().requires_grad_(True)
            yb = net(xi)
            (yb * gb).sum().backward()
            res.append((yb.detach(), xi.grad, net[0].weight.grad, net[0].bias.grad, net[0].running_var.clone()))
        for a, b, what in zip(res[1], res[0], ("output", "input gradient", "grad weight", "grad bias", "running_var")):
            close(a, b, "%s of %s" % (what, shape), rel=2e-5)
    # other input dtypes: converted on the way in and out (fp32 arithmetic), gradient in the input's dtype
    hb = syncbn.convert_sync_batchnorm(nn.Sequential(nn.BatchNorm1d(12)).to(dev))
    xh = (torch.randn(8, 12, 64, device=dev, generator=g) * (1 + rank)).half().requires_grad_(True)
    yh = hb(xh)
    assert yh.dtype == torch.float16
    yh.float().square().sum().backward()
    assert xh.grad.dtype == torch.float16 and torch.isfinite(xh.grad.float()).all()
    with torch.no_grad():
        close(yh.float(), hb(xh.detach().float()), "half input", rel=2e-3)
    # the layers replay inside a CUDA graph (device-side epochs, fixed pointers).  A fresh copy whose first backward
    # runs on the warm-up stream: gradient accumulators created on the legacy default stream cannot be captured
    ours = syncbn.convert_sync_batchnorm(copy.deepcopy(base))
    xs = torch.randn(8, 6, 300, device=dev, generator=g)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            z, _ = ours(xs)
            z.sum().backward()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    for p in ours.parameters():
        p.grad = None
    with torch.cuda.graph(graph):
        zg, _ = ours(xs)
        zg.sum().backward()
    for _ in range(5):
        xs.copy_(torch.randn(8, 6, 300, device=dev, generator=g))
        graph.replay()
    torch.cuda.synchronize()
    for p in ref.parameters():
        p.grad = None
    ref.load_state_dict(ours.state_dict())
    ours.eval(), ref.eval()
    close(ours(xs)[0], ref(xs)[0], "eval mode after graph replays")
    dist.barrier()
    if rank == 0:
        print("syncbn check ok (world %d)" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
