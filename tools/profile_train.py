"""Top CUDA kernels of one MHCT training step (torch.profiler) -- how much of the step is the Splat/Slice path."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cloud_transformers_b200.mhct import ScanObjectTrunk  # noqa: E402

torch.backends.cudnn.benchmark = "--no-benchmark" not in sys.argv
dev = torch.device("cuda:0")
model = ScanObjectTrunk().to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
gen = torch.Generator(device=dev).manual_seed(0)
pcd = bench.surface_clouds(gen, 32, 2048, dev)
y = torch.randint(0, 15, (32,), device=dev)


def step():
    logits, _ = model(pcd)
    loss = torch.nn.functional.cross_entropy(logits, y)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
print("ms per step", e0.elapsed_time(e1) / 5, "cudnn.benchmark", torch.backends.cudnn.benchmark)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
tot = sum(r.device_time_total for r in rows)
ours = sum(r.device_time_total for r in rows if "ctb::" in r.key)
print("total device ms %.1f, ctb kernels %.1f ms (%.1f%%)" % (tot / 1e3, ours / 1e3, 100 * ours / tot))
for r in rows[:22]:
    print("%8.2f ms %5.1f%% x%-4d %s" % (r.device_time_total / 1e3, 100 * r.device_time_total / tot, r.count, r.key[:90]))
