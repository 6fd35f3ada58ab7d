"""Kernel breakdown of one training step of the reference's ScanObjectNN classifier on the B200 kernels (torch.profiler).
Single process:   python tools/profile_train.py
DDP + SyncBN:     python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/profile_train.py
Prints ms per step (CUDA events), the share of the ctb:: kernels, NCCL kernels and the top kernels (rank 0)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.benchmark = "--no-benchmark" not in sys.argv
torch.manual_seed(42)
model, _ = bench.load_reference_model_through_dropin("model_zoo/scanobject/classifier.py")
model = model.to(dev)
if world > 1:
    model = torch.nn.parallel.DistributedDataParallel(torch.nn.SyncBatchNorm.convert_sync_batchnorm(model),
                                                      device_ids=[local], output_device=local)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
gen = torch.Generator(device=dev).manual_seed(rank)
pcd = bench.surface_clouds(gen, 32, 2048, dev)[:, :, None]
y = torch.randint(0, 15, (32,), device=dev)
mask = (torch.rand(32, 2048, device=dev) < 0.7).float()
ce, bce = torch.nn.CrossEntropyLoss(), torch.nn.BCEWithLogitsLoss()


def step():
    class_pred, mask_pred, _ = model(pcd)
    loss = 0.5 * ce(class_pred, y) + 0.5 * bce(mask_pred[:, 0, 0], mask)
    loss.backward()
    opt.step()
    opt.zero_grad()


for _ in range(5):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
if rank == 0:
    rows = [r for r in prof.key_averages() if r.device_time_total > 0 and r.device_type == torch.autograd.DeviceType.CUDA]
    rows = sorted(rows, key=lambda r: -r.device_time_total)
    tot = sum(r.device_time_total for r in rows)
    ours = sum(r.device_time_total for r in rows if "ctb::" in r.key)
    nccl = [r for r in rows if "nccl" in r.key.lower()]
    print("world %d: %.1f ms per step (CUDA events), cudnn.benchmark %s" % (world, ms, torch.backends.cudnn.benchmark))
    print("device-busy %.1f ms in %d kernel launches; ctb kernels %.1f ms (%.1f%%); nccl %.1f ms in %d launches" % (
        tot / 1e3, sum(r.count for r in rows), ours / 1e3, 100 * ours / tot,
        sum(r.device_time_total for r in nccl) / 1e3, sum(r.count for r in nccl)))
    for r in rows[:28]:
        print("%8.2f ms %5.1f%% x%-4d %s" % (r.device_time_total / 1e3, 100 * r.device_time_total / tot, r.count, r.key[:100]))
if world > 1:
    dist.destroy_process_group()
