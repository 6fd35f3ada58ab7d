"""Per-op timing of the six ScanObjectNN shape classes (CUDA events, L2 flushed before every launch) plus a
cross-check of the plan-based kernels against the plan-free tile kernels on the same inputs.
    python tools/bench_ops.py [--classes a2d,b3d] [--reps 5] [--check]
Prints one line per (class, op): ms, algorithmic GB/s, fraction of the measured HBM peak."""
import argparse
import ctypes
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cloud_transformers_b200 import _lib  # noqa: E402
from cloud_transformers_b200.hotpath import HotPath, algorithmic_bytes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--classes", default="a2d,a3d,b2d,b3d,c2d,c3d")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--mode", default="auto")
ap.add_argument("--check", action="store_true")
ap.add_argument("--batch", type=int, default=bench.B_PER_GPU)
ap.add_argument("--pairs", action="store_true", help="also time the 2-D / 3-D class pairs on two streams")
args = ap.parse_args()
dev = torch.device("cuda:0")
peak, _ = bench.peak_hbm()
gen = torch.Generator(device=dev).manual_seed(42)
tot = 0.0
for name, dim, W, F in bench.CLASSES:
    if name not in args.classes.split(","):
        continue
    keys, feat, conv, go, gz = bench.make_class_inputs(gen, dim, W, F, args.batch, dev)
    hp = HotPath(W, bench.H, dim, args.batch, F, bench.N_PTS, dev, mode=args.mode)
    ab = algorithmic_bytes(bench.N_PTS, dim, F, W ** dim)
    calls = {"plan": (lambda: hp.build_plan(keys)) if hp.plan is not None else None,
             "splat_fwd": lambda: hp.splat_fwd_only(keys, feat), "slice_fwd": lambda: hp.slice_fwd(keys, conv),
             "slice_bwd": lambda: hp.slice_bwd(keys, conv, go), "splat_bwd": lambda: hp.splat_bwd(keys, feat, gz)}
    hp.fwd_bwd(keys, feat, conv, go, gz)
    torch.cuda.synchronize()
    cls_ms = 0.0
    for op, fn in calls.items():
        if fn is None:
            continue
        ts = []
        for _ in range(args.reps):
            bench.flush_l2(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = statistics.median(ts)
        cls_ms += ms
        nbytes = ab.get(op, 0) * args.batch * bench.H
        print("%s %-9s %.4f ms  %7.1f GB/s  %.3f of peak" % (name, op, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak))
    print("%s total %.4f ms  (%.3f of peak)" % (name, cls_ms, ab["total"] * args.batch * bench.H / cls_ms / 1e6 / peak))
    tot += cls_ms
    if args.check and hp.plan is not None:
        # the same ops without a plan (tile scatters) must agree: z / arg bit-exact, grad_grid to rel 1e-5
        z1, a1 = hp.z.clone(), hp.arg.clone()
        hp.slice_bwd(keys, conv, go)
        gg1 = hp.grad_grid.clone()
        plan, hp.plan = hp.plan, None
        hp.splat_fwd_only(keys, feat)
        hp.slice_bwd(keys, conv, go)
        torch.cuda.synchronize()
        hp.plan = plan
        ok_z = torch.equal(z1, hp.z)
        ok_a = torch.equal(a1, hp.arg)
        err = (gg1.float() - hp.grad_grid.float()).abs().max().item() / max(1e-30, hp.grad_grid.float().abs().max().item())
        print("%s check: z %s arg %s grad_grid max rel err %.2e" % (name, ok_z, ok_a, err))
print("class set total %.4f ms -> step %.3f ms" % (tot, tot * bench.REPEATS))


if args.pairs:
    # the 2-D and the 3-D head of a MultiHeadUnion are independent (layers/multihead_ct.py:191-194): run their passes on
    # two streams and compare with back-to-back execution on one
    gen = torch.Generator(device=dev).manual_seed(42)
    cls = {c[0]: c for c in bench.CLASSES}
    for a, b in (("a2d", "a3d"), ("b2d", "b3d"), ("c2d", "c3d")):
        items = []
        for name in (a, b):
            _, dim, W, F = cls[name]
            items.append((HotPath(W, bench.H, dim, args.batch, F, bench.N_PTS, dev, mode=args.mode),
                          bench.make_class_inputs(gen, dim, W, F, args.batch, dev)))
        streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]

        def seq():
            for hp, d in items:
                hp.fwd_bwd(*d)

        def par():
            main = torch.cuda.current_stream()
            for st, (hp, d) in zip(streams, items):
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    hp.fwd_bwd(*d)
            for st in streams:
                main.wait_stream(st)

        for fn, label in ((seq, "one stream"), (par, "two streams")):
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.reps):
                bench.flush_l2(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print("%s+%s %-11s %.4f ms" % (a, b, label, statistics.median(ts)))
