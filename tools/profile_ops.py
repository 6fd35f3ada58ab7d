"""Run every hot-path op of the six ScanObjectNN shape classes a few times (for ncu launch lists / captures).
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(tile_|gather_|scatter_|plan_|splat_|slice_)' \
        --csv --log-file gpurun_out/launches.csv python tools/profile_ops.py --mode auto --reps 2
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cloud_transformers_b200.hotpath import HotPath  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="auto")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--classes", default="a2d,a3d,b2d,b3d,c2d,c3d")
ap.add_argument("--batch", type=int, default=bench.B_PER_GPU)
args = ap.parse_args()
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(42)
for name, dim, W, F in bench.CLASSES:
    if name not in args.classes.split(","):
        continue
    data = bench.make_class_inputs(gen, dim, W, F, args.batch, dev)
    hp = HotPath(W, bench.H, dim, args.batch, F, bench.N_PTS, dev, mode=args.mode)
    torch.cuda.synchronize()
    for _ in range(args.reps):
        hp.fwd_bwd(*data)
    torch.cuda.synchronize()
print("done")
