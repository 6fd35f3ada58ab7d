"""bench.py -- Splat+Slice fwd+bwd throughput of the B200 hot path (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode auto|atomic|deterministic]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload = "scanobjectnn_hotpath"): the Splat / Slice operators of ONE ScanObjectNN
classifier training step (BASELINE.json configs[1]: model_zoo/scanobject/classifier.py, B=32, N=2048, H=16):
24 MHCT blocks = 4 x the six shape classes {2D 128^2 F4, 3D 32^3 F4, 2D 64^2 F16, 3D 16^3 F16, 2D 16^2 F16,
3D 8^3 F32}; a "step" runs positions+Splat fwd, Slice fwd, Slice bwd, Splat bwd (incl. coordinate
gradients) for all 24.  The grid convolution between Splat and Slice is outside the metric.

  value  = Gpt-heads/s = sum over the 24 blocks of B*H*N / time, inputs resident in HBM (C-ABI calls only)
  e2e    = same metric through the public modules (DifferentiablePositions / Splat / Slice + autograd) with
           HOST pinned inputs copied H2D every step and the loss read back D2H
  roofline = the op with the largest share of the step, algorithmic bytes / CUDA-event time
  cpu_baseline = the reference's torch composition (oracle/ct_torch.py, kind "port") on the host cores
Multi-GPU: weak scaling, every rank runs its own batch, no data-path collective.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# (name, dim, W, F) of the six shape classes, classifier.py:46-63; H = 16, N = 2048, B = 32 per GPU
CLASSES = [("a2d", 2, 128, 4), ("a3d", 3, 32, 4), ("b2d", 2, 64, 16), ("b3d", 3, 16, 16), ("c2d", 2, 16, 16),
           ("c3d", 3, 8, 32)]
REPEATS = 4
H, N_PTS, B_PER_GPU = 16, 2048, 32
METRIC = "splat_slice_fwd_bwd_gpt_heads_per_s"
UNIT = "Gpt-heads/s"


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def surface_clouds(gen, B, N, device):
    """Synthetic ScanObjectNN-like clouds: points on a union of random planes / ellipsoid shells, centred and
    max-norm normalised like datasets/scanobjectnn.py:44-62."""
    import torch
    k = 4
    which = torch.randint(0, k, (B, N), generator=gen, device=device)
    u = torch.rand(B, N, 2, generator=gen, device=device) * 2 - 1
    basis = torch.randn(B, k, 3, 3, generator=gen, device=device)
    offs = torch.randn(B, k, 3, generator=gen, device=device) * 0.3
    b = torch.gather(basis, 1, which[:, :, None, None].expand(-1, -1, 3, 3))
    o = torch.gather(offs, 1, which[:, :, None].expand(-1, -1, 3))
    plane = u[..., 0:1] * b[:, :, 0] + u[..., 1:2] * b[:, :, 1] + o
    sph = torch.nn.functional.normalize(torch.randn(B, N, 3, generator=gen, device=device), dim=-1) * \
        (0.3 + 0.5 * torch.rand(B, 1, 1, generator=gen, device=device)) + o
    pts = torch.where((which % 2 == 0)[..., None], plane, sph)
    pts = pts - pts.mean(1, keepdim=True)
    pts = pts / pts.norm(dim=-1).max(dim=1)[0][:, None, None]
    return pts.permute(0, 2, 1).contiguous()          # [B, 3, N]


def make_class_inputs(gen, dim, W, F, B, device, grid_dtype=None):
    """keys = tanh(per-head rotated + shifted cloud) (what MultiHead feeds DifferentiablePositions,
    multihead_ct.py:93-99), features ~ N(0,1), plus the conv output and the two incoming gradients."""
    import torch
    grid_dtype = grid_dtype or torch.float32
    pcd = surface_clouds(gen, B, N_PTS, device)
    rot = torch.linalg.qr(torch.randn(H, 3, 3, generator=gen, device=device))[0]
    keys3 = torch.einsum("bcp,hcn->bhnp", pcd * 1.5, rot) + 0.05 * torch.randn(B, H, 3, N_PTS, generator=gen,
                                                                              device=device)
    keys = torch.tanh(keys3[:, :, :dim].reshape(B, H * dim, N_PTS)).contiguous()
    feat = torch.randn(B, H * F, N_PTS, generator=gen, device=device)
    grid_shape = (B, H * F) + (W,) * dim
    conv = torch.randn(grid_shape, generator=gen, device=device).to(grid_dtype)
    gz = torch.randn(grid_shape, generator=gen, device=device).to(grid_dtype)
    go = torch.randn(B, H * F, N_PTS, generator=gen, device=device)
    return keys, feat, conv, go, gz


def run_ours(args):
    import torch
    import torch.distributed as dist
    import cloud_transformers_b200 as ctb
    from cloud_transformers_b200 import _lib
    from cloud_transformers_b200.hotpath import HotPath, algorithmic_bytes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    ctb.config.mode = args.mode
    gen = torch.Generator(device=dev).manual_seed(42 + rank)
    B = B_PER_GPU
    gdt = torch.bfloat16 if args.grid_dtype == "bf16" else torch.float32
    eg = 2 if args.grid_dtype == "bf16" else 4
    data, paths = {}, {}
    for name, dim, W, F in CLASSES:
        data[name] = make_class_inputs(gen, dim, W, F, B, dev, gdt)
        paths[name] = HotPath(W, H, dim, B, F, N_PTS, dev, mode=args.mode, grid_dtype=gdt)
    order = [c for _ in range(REPEATS) for c in CLASSES]          # same class never back to back => L2 is cold
    step_bytes = sum(algorithmic_bytes(N_PTS, dim, F, W ** dim, e_grid=eg)["total"] * B * H for _, dim, W, F in order)
    pt_heads_step = len(order) * B * H * N_PTS

    def step():
        for name, dim, W, F in order:
            keys, feat, conv, go, gz = data[name]
            paths[name].fwd_bwd(keys, feat, conv, go, gz)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    barrier()
    ms_per_step = ms / args.steps
    value = world * pt_heads_step / (ms_per_step * 1e-3) / 1e9

    # ---- per-op CUDA-event timing for the roofline (same stream, same buffers, after the timed region)
    ops = []
    reps = max(3, min(args.steps, 10))
    for name, dim, W, F in CLASSES:
        keys, feat, conv, go, gz = data[name]
        hp = paths[name]
        ab = algorithmic_bytes(N_PTS, dim, F, W ** dim, e_grid=eg)
        calls = {"splat_fwd": lambda: hp.splat_fwd(keys, feat), "slice_fwd": lambda: hp.slice_fwd(keys, conv),
                 "slice_bwd": lambda: hp.slice_bwd(keys, conv, go), "splat_bwd": lambda: hp.splat_bwd(keys, feat, gz)}
        for op, fn in calls.items():
            ts = []
            for _ in range(reps):
                flush_l2(dev)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            t_ms = statistics.median(ts)
            nbytes = ab[op] * B * H
            ops.append({"op": op, "class": name, "ms": round(t_ms, 4), "gbs": round(nbytes / t_ms / 1e6, 1),
                        "bytes": nbytes})
    peak, peak_kind = peak_hbm()
    total_op_ms = sum(o["ms"] for o in ops)
    top = max(ops, key=lambda o: o["ms"])
    clocks = sampler.stop() if rank == 0 else None

    # ---- context: the same step with the grids stored as bf16 (fp32 arithmetic; tolerance rel 1e-2, not the headline)
    bf16_mode = None
    if args.grid_dtype == "f32" and rank == 0:
        try:
            g2 = torch.Generator(device=dev).manual_seed(43)
            d16, p16 = {}, {}
            for name, dim, W, F in CLASSES:
                d16[name] = make_class_inputs(g2, dim, W, F, B, dev, torch.bfloat16)
                p16[name] = HotPath(W, H, dim, B, F, N_PTS, dev, mode=args.mode, grid_dtype=torch.bfloat16)

            def step16():
                for name, dim, W, F in order:
                    p16[name].fwd_bwd(*d16[name])

            step16()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                step16()
            b.record()
            torch.cuda.synchronize()
            ms16 = a.elapsed_time(b) / 3
            bytes16 = sum(algorithmic_bytes(N_PTS, dim, F, W ** dim, e_grid=2)["total"] * B * H for _, dim, W, F in order)
            bf16_mode = {"value": round(pt_heads_step / (ms16 * 1e-3) / 1e9, 4), "unit": UNIT, "ms_per_step": round(ms16, 3),
                         "algorithmic_gbs": round(bytes16 / (ms16 * 1e-3) / 1e9, 1),
                         "what": "grids (z, convolved, grad_grid, grad_z) stored as bf16, fp32 arithmetic, this rank only"}
            del d16, p16
        except Exception as exc:  # noqa: BLE001
            bf16_mode = {"unavailable": repr(exc)[:200]}

    # ---- e2e through the public modules with host buffers -----------------------------------------
    e2e = run_e2e(args, dev, world, rank, data, order)
    ref_gpu = reference_composition_on_gpu(dev, data) if rank == 0 else None
    n_launches = args.steps * sum(paths[n].launches_per_pass() for n, _, _, _ in order)
    del data, paths
    torch.cuda.empty_cache()
    train = run_train(args, dev, world, rank) if args.train_steps > 0 else None

    if rank == 0:
        cpu = cpu_baseline(sample_batch=8, repeats=3)
        line = {
            "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.grid_dtype == "f32" else "f32 arithmetic, bf16 grid storage", "data": "synthetic",
            "config": {"workload": "scanobjectnn_hotpath", "blocks_per_step": len(order), "batch_per_gpu": B,
                       "heads": H, "points": N_PTS, "classes": [c[0] for c in CLASSES], "mode": args.mode,
                       "l2": "working set per step %.1f GB >> 126 MB L2; same class never back to back" %
                             (step_bytes / 1e9)},
            "algorithmic_gbs": round(step_bytes / (ms_per_step * 1e-3) / 1e9, 1),
            "frac_of_hbm_peak": round(step_bytes / (ms_per_step * 1e-3) / 1e9 / peak, 4),
            "roofline": {"bound": "hbm", "kernel": "%s/%s" % (top["op"], top["class"]), "achieved": top["gbs"],
                         "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": round(top["gbs"] / peak, 4),
                         "traffic": ncu_traffic(top["op"], top["class"]), "algorithmic_bytes": top["bytes"],
                         "share_of_step": round(top["ms"] / total_op_ms, 4)},
            "ops": ops,
            "e2e": e2e,
            "train": train,
            "gpu_launches": n_launches,
            "clocks": clocks,
            "cpu_baseline": cpu,
            "reference_composition_gpu": ref_gpu,
            "bf16_grid_mode": bf16_mode,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(op, cls):
    """dram__bytes_read.sum + dram__bytes_write.sum of the kernel(s) of one op, per launch, from the committed
    `ncu --set full` capture of tools/profile_ops.py (profiles/r01_ncu_full_summary.csv: 5 kernels per class in the order
    Splat fwd | Slice fwd | Slice bwd scatter, Slice bwd gather | Splat bwd; classes in CLASSES order)."""
    import csv
    path = os.path.join(ROOT, "profiles", "r01_ncu_full_summary.csv")
    try:
        rows = list(csv.reader(open(path)))
        hdr = rows[0]
        ir = [i for i, h in enumerate(hdr) if h.startswith("dram__bytes_read.sum")][0]
        iw = [i for i, h in enumerate(hdr) if h.startswith("dram__bytes_write.sum")][0]
        unit_r, unit_w = hdr[ir], hdr[iw]
        scale = lambda u: 1e9 if "Gbyte" in u else (1e6 if "Mbyte" in u else (1e3 if "Kbyte" in u else 1.0))
        ci = [c[0] for c in CLASSES].index(cls)
        pick = {"splat_fwd": [0], "slice_fwd": [1], "slice_bwd": [2, 3], "splat_bwd": [4]}[op]
        body = rows[1:]
        if len(body) < 5 * len(CLASSES):
            return None
        return int(sum(float(body[ci * 5 + k][ir]) * scale(unit_r) + float(body[ci * 5 + k][iw]) * scale(unit_w)
                       for k in pick))
    except Exception:
        return None


_FLUSH = {}


def flush_l2(dev):
    import torch
    if dev not in _FLUSH:
        _FLUSH[dev] = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    _FLUSH[dev].zero_()


def run_e2e(args, dev, world, rank, data, order):
    """Same step through DifferentiablePositions / Splat / Slice + autograd, inputs in pinned host memory."""
    import torch
    import torch.distributed as dist
    import cloud_transformers_b200 as ctb
    mods, host = {}, {}
    for name, dim, W, F in CLASSES:
        mods[name] = (ctb.DifferentiablePositions(tensor_size=W, heads=H, dim=dim).to(dev),
                      ctb.Splat(tensor_size=W, heads=H, dim=dim,
                                out_dtype=torch.bfloat16 if args.grid_dtype == "bf16" else None).to(dev),
                      ctb.Slice(tensor_size=W, heads=H, dim=dim).to(dev))
        keys, feat = data[name][0], data[name][1]
        host[name] = (keys.cpu().pin_memory(), feat.cpu().pin_memory())
    h2d = sum(host[n][0].numel() * 4 + host[n][1].numel() * 4 for n, _, _, _ in order)
    loss_host = torch.zeros(len(order), dtype=torch.float32).pin_memory()

    # Input pipeline: the next block's host->device copies run on a copy stream while the current block computes
    # (one block of prefetch).  Every copy still starts after the step's first timing event and is waited on by the
    # compute stream, so the timed region covers all of them.
    copy_stream = torch.cuda.Stream(device=dev)

    def upload(name):
        with torch.cuda.stream(copy_stream):
            k = host[name][0].to(dev, non_blocking=True)
            f = host[name][1].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return k, f, ev

    def step():
        main = torch.cuda.current_stream()
        copy_stream.wait_stream(main)
        nxt = upload(order[0][0])
        for i, (name, dim, W, F) in enumerate(order):
            dp, sp, sl = mods[name]
            k, f, ev = nxt
            if i + 1 < len(order):
                nxt = upload(order[i + 1][0])
            main.wait_event(ev)
            k.record_stream(main)
            f.record_stream(main)
            k.requires_grad_(True)
            f.requires_grad_(True)
            lc, idx = dp(k)
            z = sp(lc, idx, f)
            out = sl(lc, idx, z)
            loss = out.square().mean()
            loss.backward()
            loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.synchronize()
        return float(loss_host.sum())

    steps = max(2, min(args.steps, 5))
    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    val = world * len(order) * B_PER_GPU * H * N_PTS / (ms * 1e-3) / 1e9
    return {"value": round(val, 4), "unit": UNIT, "ms_per_step": round(ms, 3), "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": 4 * len(order), "api": "DifferentiablePositions/Splat/Slice modules + autograd",
            "input_pipeline": "pinned host buffers, one block of H2D prefetch on a copy stream"}


def reference_composition_on_gpu(dev, data):
    """Context only (not the reference arm): the reference's torch op composition (materialised pre_splat, scatter-max,
    expanded int64 index, gather; oracle/ct_torch.py, scatter_max via the first-winner shim) run on the SAME B200 on
    the same six class inputs at the same batch -- the bar a user of the reference sees on this GPU."""
    import torch
    from oracle import ct_torch as T
    try:
        def once():
            for name, dim, W, F in CLASSES:
                keys, feat = data[name][0], data[name][1]
                T.hot_path_fwd_bwd(keys, feat, W, H, dim)
        once()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            once()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        val = len(CLASSES) * B_PER_GPU * H * N_PTS / (ms * 1e-3) / 1e9
        return {"value": round(val, 4), "unit": UNIT, "ms_per_class_set": round(ms, 3),
                "what": "torch composition of the reference ops on this GPU, six classes once each at B=%d" % B_PER_GPU}
    except Exception as exc:  # e.g. out of memory on a smaller part
        return {"unavailable": repr(exc)[:200]}
    finally:
        torch.cuda.empty_cache()


def run_train(args, dev, world, rank):
    """Second half of the BASELINE metric: MHCT training samples/s.  A ScanObjectNN-classifier-shaped trunk built from
    the block mirrors (cloud_transformers_b200/mhct.py: 12 MultiHeadUnion + 2 MultiHeadPool, 24 Splat + 24 Slice + 2
    pool Splats per forward, like model_zoo/scanobject/classifier.py), Adam, cross-entropy on synthetic labels, clouds
    copied from pinned host memory every step; DDP + SyncBatchNorm over NCCL when world > 1 (what
    train_classification.py:107-109 does).  Convolutions / BatchNorm / Linear are stock PyTorch."""
    import torch
    import torch.distributed as dist
    from cloud_transformers_b200.mhct import ScanObjectTrunk
    try:
        torch.manual_seed(1234 + rank)
        torch.backends.cudnn.benchmark = True          # as train_classification.py:58
        B = args.train_batch
        model = ScanObjectTrunk().to(dev)
        if world > 1:
            model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
            # SyncBN keeps the running statistics identical on all ranks, so the per-step buffer broadcast of stock
            # DDP is redundant; gradients are reduced in place in the buckets
            model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], broadcast_buffers=False,
                                                              gradient_as_bucket_view=True)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        gen = torch.Generator(device=dev).manual_seed(7 + rank)
        clouds = [surface_clouds(gen, B, N_PTS, dev).cpu().pin_memory() for _ in range(4)]
        labels = [torch.randint(0, 15, (B,)).pin_memory() for _ in range(4)]
        loss_host = torch.zeros(1).pin_memory()

        def step(i):
            pcd = clouds[i % 4].to(dev, non_blocking=True)
            y = labels[i % 4].to(dev, non_blocking=True)
            logits, _ = model(pcd)
            loss = torch.nn.functional.cross_entropy(logits, y)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)

        for i in range(5 if world > 1 else 3):      # (DDP rebuilds its buckets after the first step)
            step(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.train_steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.train_steps
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        n_params = sum(p.numel() for p in model.parameters())
        return {"samples_per_s": round(world * B / (ms * 1e-3), 2), "ms_per_step": round(ms, 2), "batch_per_gpu": B,
                "points": N_PTS, "steps": args.train_steps, "params_m": round(n_params / 1e6, 2),
                "final_loss": round(float(loss_host[0]), 4),
                "model": "ScanObjectTrunk (classifier.py trunk: 12 MultiHeadUnion + 2 MultiHeadPool; towers -> avgpool)",
                "parallelism": "dp%d (DDP + SyncBN over NCCL)" % world if world > 1 else "single GPU"}
    except Exception as exc:
        return {"unavailable": repr(exc)[:300]}


def cpu_step(sample_batch, threads):
    """One bounded sample of the workload on the host: the six shape classes once each at batch
    `sample_batch` through the reference's torch composition (oracle/ct_torch.py)."""
    import torch
    from oracle import ct_torch as T
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    t0 = time.perf_counter()
    for name, dim, W, F in CLASSES:
        keys = torch.tanh(torch.randn(sample_batch, H * dim, N_PTS, generator=g))
        feat = torch.randn(sample_batch, H * F, N_PTS, generator=g)
        T.hot_path_fwd_bwd(keys, feat, W, H, dim)
    dt = time.perf_counter() - t0
    return len(CLASSES) * sample_batch * H * N_PTS / dt / 1e9, dt


def cpu_baseline(sample_batch=1, repeats=1):
    import torch
    cores = os.cpu_count() or 1
    cpu_step(sample_batch, cores)                     # warm
    vals = [cpu_step(sample_batch, cores) for _ in range(repeats)]
    best = max(v for v, _ in vals)
    return {"value": round(best, 8), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "six shape classes once each at B=%d (H=16, N=2048), torch CPU port of the reference "
                      "composition (oracle/ct_torch.py), %d threads, %.1f s per pass" % (sample_batch, cores,
                                                                                         vals[-1][1])}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (its torch composition; the
    reference is Python and cannot travel, so the port in oracle/ct_torch.py stands in), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_batch = 8
    for _ in range(args.warmup):
        cpu_step(sample_batch, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(sample_batch, cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = len(CLASSES) * sample_batch * H * N_PTS / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 8), "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "scanobjectnn_hotpath", "heads": H, "points": N_PTS,
                   "classes": [c[0] for c in CLASSES], "sample_batch": sample_batch},
        "cpu_baseline": {"value": round(val, 8), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "six shape classes once each at B=%d per step" % sample_batch},
        "e2e": {"value": round(val, 8), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--train-steps", type=int, default=10, help="steps of the MHCT training throughput add-on (0 = skip)")
    ap.add_argument("--train-batch", type=int, default=B_PER_GPU)
    ap.add_argument("--grid-dtype", default="f32", choices=["f32", "bf16"],
                    help="bf16: grids (z, convolved, grad_grid, grad_z) stored as bf16, arithmetic stays fp32")
    ap.add_argument("--mode", default=os.environ.get("CTB_MODE", "auto"), choices=["auto", "atomic", "tile", "deterministic"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
