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
  cpu_baseline = the reference's OWN layers/cloud_transform.py (staged verbatim in oracle/_ref by oracle/stage_ref.py,
           imported under the two dependency shims; kind "reference") on the host cores
  train  = MHCT training samples/s: the reference's own model_zoo/scanobject/classifier.py (24.02 M parameters)
           running on the B200 kernels through dropin/, loop as train_classification.py:181-273
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
# further BASELINE.json configs, timed as extra keys of the line: (name, dim, W, F, N, B, H)
OTHER_CONFIGS = [("s3dis_a3d", 3, 32, 4, 4096, 8, 16), ("s3dis_b3d", 3, 16, 16, 4096, 8, 16),
                 ("inpaint_dec_a2d", 2, 128, 4, 16384, 2, 16), ("inpaint_dec_a3d", 3, 32, 4, 16384, 2, 16),
                 ("sweep_n256k_2d256", 2, 256, 4, 262144, 4, 16), ("sweep_h64_3d64", 3, 64, 4, 4096, 8, 64)]


def workload_config(mode):
    """the `config` object of the line: identical for both arms (the reference arm times a bounded sample of it)"""
    step_bytes = 0
    from_bytes = lambda N, d, F, C: N * (24 * d + 20 * F) + 24 * F * C
    for _, dim, W, F in CLASSES:
        step_bytes += REPEATS * B_PER_GPU * H * from_bytes(N_PTS, dim, F, W ** dim)
    return {"workload": "scanobjectnn_hotpath", "blocks_per_step": REPEATS * len(CLASSES), "batch_per_gpu": B_PER_GPU,
            "heads": H, "points": N_PTS, "classes": [c[0] for c in CLASSES],
            "block_pairs": "the 2-D and the 3-D head of a MultiHeadUnion run concurrently (two streams), pairs in sequence",
            "l2": "working set per step %.1f GB >> 126 MB L2; same class never back to back" % (step_bytes / 1e9)}


def reference_tree_root():
    """where the unmodified reference files are: $CTB_REFERENCE_ROOT, the build container's /root/reference, or the
    verbatim copy staged by oracle/stage_ref.py (the only one that exists on the GPU box)"""
    for cand in (os.environ.get("CTB_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "layers", "cloud_transform.py")):
            return cand
    return None


def load_reference_model_through_dropin(rel_path, **params):
    """What a user of the reference does to switch: dropin/ ahead of the reference checkout on
    sys.path.  `layers` has no __init__.py (namespace package), so layers.cloud_transform resolves to
    dropin/layers/cloud_transform.py and everything else -- layers.multihead_ct*, unet2d.*, the model file, exec'd as
    in utils/train_util.py:23-27 -- is the reference's own, unmodified code."""
    root = reference_tree_root()
    if root is None:
        raise RuntimeError("no reference tree (run oracle/stage_ref.py in the build container)")
    for k in [k for k in sys.modules if k.split(".")[0] in ("layers", "unet2d", "utils", "model_zoo")]:
        del sys.modules[k]
    paths = [os.path.join(ROOT, "dropin"), root]
    try:
        import pytorch3d.transforms.so3  # noqa: F401
    except Exception:
        paths.insert(0, os.path.join(ROOT, "dropin", "_optional_shims"))
    for pth in reversed(paths):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    env = {}
    with open(os.path.join(root, rel_path)) as f:
        exec(compile(f.read(), os.path.join(root, rel_path), "exec"), env)
    import layers.cloud_transform as ct
    assert os.path.realpath(ct.__file__).startswith(os.path.realpath(os.path.join(ROOT, "dropin")))
    return env["Model"](**params), root


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
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")     # (CUDA-graph capture of NCCL collectives)
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

    # The 2-D and the 3-D head of a MultiHeadUnion block are independent (layers/multihead_ct.py:191-194 loops over
    # them; classifier.py:46-63 pairs 128^2 with 32^3, 64^2 with 16^3, 16^2 with 8^3), block k + 1 needs both heads of
    # block k: each pair runs on two streams and joins before the next pair, so one head's CTAs fill the SMs the other
    # head's last wave leaves idle.
    side = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]

    def step():
        main = torch.cuda.current_stream()
        for i in range(0, len(order), 2):
            for st, (name, dim, W, F) in zip(side, order[i:i + 2]):
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    paths[name].fwd_bwd(*data[name])
            for st in side:
                main.wait_stream(st)

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
    from cloud_transformers_b200.sharding import aggregate_throughput
    units, ms = aggregate_throughput(pt_heads_step, e0.elapsed_time(e1), dev)      # sum of units, max of time over ranks
    barrier()
    ms_per_step = ms / args.steps
    value = units / (ms_per_step * 1e-3) / 1e9

    # ---- per-op CUDA-event timing for the roofline (same stream, same buffers, after the timed region)
    ops = []
    reps = max(3, min(args.steps, 10))
    for name, dim, W, F in CLASSES:
        keys, feat, conv, go, gz = data[name]
        hp = paths[name]
        ab = algorithmic_bytes(N_PTS, dim, F, W ** dim, e_grid=eg)
        # plan = the once-per-block sort / entry lists where the class uses them (0 algorithmic bytes: pure overhead of
        # the path, counted in the step and in the class totals); every other row but slice_bwd is ONE kernel
        calls = {"plan": (lambda: hp.build_plan(keys)) if hp.plan is not None else None,
                 "splat_fwd": lambda: hp.splat_fwd_only(keys, feat), "slice_fwd": lambda: hp.slice_fwd(keys, conv),
                 "slice_bwd": lambda: hp.slice_bwd(keys, conv, go), "splat_bwd": lambda: hp.splat_bwd(keys, feat, gz)}
        for op, fn in calls.items():
            if fn is None:
                continue
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
            nbytes = ab.get(op, 0) * B * H
            ops.append({"op": op, "class": name, "ms": round(t_ms, 4), "gbs": round(nbytes / t_ms / 1e6, 1),
                        "bytes": nbytes})
    peak, peak_kind = peak_hbm()
    total_op_ms = sum(o["ms"] for o in ops)
    # the dominant single KERNEL: slice_bwd is two kernels (grad_grid scatter + key-gradient gather), plan is overhead
    top = max((o for o in ops if o["op"] in ("splat_fwd", "slice_fwd", "splat_bwd")), key=lambda o: o["ms"])
    clocks = sampler.stop() if rank == 0 else None

    # ---- context: the same step with the grids stored as bf16 (fp32 arithmetic; tolerance rel 1e-2, not the headline)
    bf16_mode = None
    if args.grid_dtype == "f32" and rank == 0 and "bf16" not in args.skip.split(","):
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
    skip = set(args.skip.split(","))
    e2e = run_e2e(args, dev, world, rank, data, order) if "e2e" not in skip else None
    ref_gpu = reference_composition_on_gpu(dev, data) if (rank == 0 and "refgpu" not in skip) else None
    det_mode = run_deterministic(args, dev, data, order) if (rank == 0 and args.mode == "auto" and "det" not in skip) else None
    n_launches = args.steps * sum(paths[n].launches_per_pass() for n, _, _, _ in order)
    del data, paths
    torch.cuda.empty_cache()
    others = run_other_configs(args, dev) if (rank == 0 and "others" not in skip) else None
    losses = run_losses(dev) if (rank == 0 and "losses" not in skip) else None
    completion = run_completion_train(dev) if (rank == 0 and "completion" not in skip and args.train_steps > 0) else None
    torch.cuda.empty_cache()
    s3dis = run_s3dis_train(dev) if (rank == 0 and "s3dis" not in skip and args.train_steps > 0) else None
    torch.cuda.empty_cache()
    barrier()
    train = train_eager = train_fused = train_nccl = None
    if args.train_steps > 0 and "train" not in skip:
        if "eager" not in skip:
            train_eager = run_train(args, dev, world, rank, graphed=False)
            torch.cuda.empty_cache()
            barrier()
        train = run_train(args, dev, world, rank, graphed=True)
        if "unavailable" in train and train_eager is not None:          # capture failed: the eager loop is the number
            train, train_eager = train_eager, train
        if world > 1 and "fusedbn" not in skip:
            torch.cuda.empty_cache()
            barrier()
            train_fused = run_train(args, dev, world, rank, graphed=True, fused_syncbn=True)
            if "unavailable" not in train_fused:       # the framework's own SyncBN is the headline training number,
                train, train_nccl = train_fused, train  # torch.nn.SyncBatchNorm over NCCL is reported beside it

    if rank == 0:
        cpu = cpu_baseline() if "cpu" not in skip else None
        line = {
            "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.grid_dtype == "f32" else "f32 arithmetic, bf16 grid storage", "data": "synthetic",
            "config": workload_config(args.mode),
            "mode": args.mode,
            "algorithmic_gbs": round(step_bytes / (ms_per_step * 1e-3) / 1e9, 1),
            "frac_of_hbm_peak": round(step_bytes / (ms_per_step * 1e-3) / 1e9 / peak, 4),
            "roofline": {"bound": "hbm", "op": "%s/%s" % (top["op"], top["class"]), "kernel": kernel_of(top["op"], top["class"]),
                         "achieved": top["gbs"], "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                         "frac": round(top["gbs"] / peak, 4), "traffic": ncu_traffic(top["op"], top["class"]),
                         "traffic_source": NCU_SUMMARY, "algorithmic_bytes": top["bytes"],
                         "share_of_step": round(top["ms"] / total_op_ms, 4)},
            "ops": ops,
            "e2e": e2e,
            "train": train,
            "train_eager": train_eager,
            "train_nccl_syncbn": train_nccl,
            "gpu_launches": n_launches,
            "clocks": clocks,
            "cpu_baseline": cpu,
            "reference_composition_gpu": ref_gpu,
            "bf16_grid_mode": bf16_mode,
            "deterministic_mode": det_mode,
            "other_configs": others,
            "completion_losses": losses,
            "train_completion": completion,
            "train_s3dis": s3dis,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


NCU_SUMMARY = "profiles/r02_ncu_full_summary.csv"


def _ncu_rows():
    """rows of the committed `ncu --set full` summary (tools/ncu_summary.py over tools/profile_ops.py: per class the
    kernels in launch order -- [plan] | Splat fwd | Slice fwd | Slice bwd scatter, Slice bwd gather | Splat bwd)."""
    import csv
    path = os.path.join(ROOT, NCU_SUMMARY)
    rows = list(csv.reader(open(path)))
    hdr, body = rows[0], rows[1:]
    per_class, i = {}, 0
    for name, _, _, _ in CLASSES:
        ks = {}
        if i < len(body) and body[i][0].replace("void ", "").startswith("plan_build"):
            ks["plan"] = body[i]
            i += 1
        for key in ("splat_fwd", "slice_fwd", "slice_bwd_scatter", "slice_bwd_gather", "splat_bwd"):
            ks[key] = body[i]
            i += 1
        per_class[name] = ks
    return hdr, per_class


def kernel_of(op, cls):
    """name of the dominant kernel of an op (slice_bwd is two kernels: the grad_grid scatter dominates)"""
    try:
        _, pc = _ncu_rows()
        key = {"slice_bwd": "slice_bwd_scatter"}.get(op, op)
        return pc[cls][key][0].replace("void ", "")
    except Exception:
        return None


def ncu_traffic(op, cls):
    """dram__bytes_read.sum + dram__bytes_write.sum of the kernel(s) of one op, per launch, from the committed ncu
    capture of the same library (profiles/r02_ncu_full_summary.csv); None if the file is missing or malformed."""
    try:
        hdr, pc = _ncu_rows()
        ir = [i for i, h in enumerate(hdr) if h.startswith("dram__bytes_read.sum")][0]
        iw = [i for i, h in enumerate(hdr) if h.startswith("dram__bytes_write.sum")][0]
        scale = lambda u: 1e9 if "Gbyte" in u else (1e6 if "Mbyte" in u else (1e3 if "Kbyte" in u else 1.0))
        keys = {"plan": ["plan"], "splat_fwd": ["splat_fwd"], "slice_fwd": ["slice_fwd"],
                "slice_bwd": ["slice_bwd_scatter", "slice_bwd_gather"], "splat_bwd": ["splat_bwd"]}[op]
        return int(sum(float(pc[cls][k][ir]) * scale(hdr[ir]) + float(pc[cls][k][iw]) * scale(hdr[iw])
                       for k in keys if k in pc[cls]))
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
    import cloud_transformers_b200 as ctb
    from cloud_transformers_b200.sharding import aggregate_throughput
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

    # Input pipeline: the host->device copies of the next DEPTH blocks are in flight on two copy streams (keys on one,
    # features on the other) while the current block computes.  Every copy still starts after the step's first timing
    # event and is waited on by the compute stream, so the timed region covers all of them.
    DEPTH = 3
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]

    def upload(name):
        out = []
        for st, src in zip(streams, host[name]):
            with torch.cuda.stream(st):
                t = src.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(st)
            out.append((t, ev))
        return out

    def step():
        main = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(main)
        queue = [upload(order[i][0]) for i in range(min(DEPTH, len(order)))]
        for i, (name, dim, W, F) in enumerate(order):
            dp, sp, sl = mods[name]
            (k, evk), (f, evf) = queue.pop(0)
            if i + DEPTH < len(order):
                queue.append(upload(order[i + DEPTH][0]))
            main.wait_event(evk)
            main.wait_event(evf)
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
    _barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    units, ms = aggregate_throughput(len(order) * B_PER_GPU * H * N_PTS, e0.elapsed_time(e1) / steps, dev)
    return {"value": round(units / (ms * 1e-3) / 1e9, 4), "unit": UNIT, "ms_per_step": round(ms, 3),
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * len(order),
            "h2d_gbs_per_gpu": round(h2d / (ms * 1e-3) / 1e9, 1),
            "api": "DifferentiablePositions/Splat/Slice modules + autograd",
            "input_pipeline": "pinned host buffers, %d blocks of H2D prefetch on two copy streams; the step is bound by "
                              "the host link when h2d_gbs_per_gpu is near the PCIe rate of the box" % DEPTH}


def _barrier(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run_losses(dev):
    """Row N3: the completion losses at the shapes of train_inpainter.py:187-192 / :267-269 (EMD eps 0.005 x 50 iterations
    in training, 0.004 x 3000 in validation; clouds of 2048 and, from the inpainting decoder, 16384 points) and the
    Chamfer distance, through the public modules, ms per call on this rank (the reference's own kernels on the same
    GPU: profiles/r02_losses_vs_reference_kernels.txt)."""
    import torch
    from cloud_transformers_b200.chamfer import ChamferFunction
    from cloud_transformers_b200.emd import emdModule
    out = {}
    gen = torch.Generator(device=dev).manual_seed(11)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return round(a.elapsed_time(b) / reps, 4)

    try:
        for B, n, eps, iters, reps in [(32, 2048, 0.005, 50, 5), (32, 2048, 0.004, 3000, 2), (2, 16384, 0.005, 50, 3)]:
            x = torch.rand(B, n, 3, device=dev, generator=gen)
            y = (x[:, torch.randperm(n, device=dev, generator=gen)] + 0.02 * torch.randn(B, n, 3, device=dev, generator=gen)).clamp(0, 1).contiguous()
            out["emd_%dx%d_eps%g_%dit_ms" % (B, n, eps, iters)] = timed(lambda: emdModule()(x, y, eps, iters), reps)
        for B, n, m in [(32, 2048, 2048), (8, 8192, 8192)]:
            x, y = torch.rand(B, n, 3, device=dev, generator=gen), torch.rand(B, m, 3, device=dev, generator=gen)
            out["chamfer_%dx%dx%d_ms" % (B, n, m)] = timed(lambda: ChamferFunction.apply(x, y), 5)
    except Exception as exc:  # noqa: BLE001
        out["unavailable"] = repr(exc)[:200]
    return out


def run_completion_train(dev, steps=10):
    """BASELINE config 4 (inpainting): the reference's own model_zoo/completion/inpainter.py (53.4 M parameters, AdaIN
    blocks on 16384 decoder points) and the step of train_inpainter.py:176-196 -- partial_postproces on the host, EMD
    (0.005, 50) + Chamfer, Adam 1e-4 -- through dropin/ (Splat / Slice, emd_linear.emd_module and
    chamfer_extension.dist_chamfer are this library's), batch 2 as configs/inpainting.yaml, this rank only."""
    import torch
    try:
        torch.manual_seed(42)
        generator, root = load_reference_model_through_dropin("model_zoo/completion/inpainter.py")
        import chamfer_extension.dist_chamfer as dist_chamfer
        import emd_linear.emd_module as emd
        from utils.pcd_utils import partial_postproces
        n_params = sum(p.numel() for p in generator.parameters())
        generator = generator.to(dev).train()
        optimizer = torch.optim.Adam(generator.parameters(), lr=1e-4, betas=(0.9, 0.999), weight_decay=0.0)
        EMD = emd.emdModule()
        B, n_in, n_gt = 2, 2048, 16384
        g = torch.Generator().manual_seed(3)
        u = torch.randn(B, n_gt, 3, generator=g)
        gt = (0.5 * u / u.norm(dim=-1, keepdim=True) * torch.tensor([1.0, 0.6, 0.4])).pin_memory()
        partial = gt[:, :n_in].clone()
        partial[:, 1500:] = 0.0
        last = [None]

        def step():
            pcd_gt = 2 * gt.permute(0, 2, 1)[:, :, None].to(dev, non_blocking=True)
            pcd_part_enc, pcd_part_noise = partial_postproces(2 * partial, pcd_gt.shape[-1])
            pcd_part_enc = pcd_part_enc.permute(0, 2, 1)[:, :, None].to(dev)
            pcd_part_noise = pcd_part_noise.permute(0, 2, 1).to(dev)
            reconstruction, _ = generator(pcd_part_noise, pcd_part_enc)
            dist, _ = EMD(reconstruction[:, :, 0].permute(0, 2, 1), pcd_gt[:, :, 0].permute(0, 2, 1), 0.005, 50)
            loss_emd = torch.sqrt(dist).mean(1).mean()
            loss_chamfer = dist_chamfer.loss_chamfer(reconstruction, pcd_gt)
            loss = loss_emd + 0.0 * loss_chamfer                  # chamfer_weight 0.0, configs/inpainting.yaml:24
            loss.backward()
            optimizer.step()
            optimizer.zero_grad()
            last[0] = loss.item()                                 # the script logs it every step (:200-207)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"samples_per_s": round(B / (ms * 1e-3), 2), "ms_per_step": round(ms, 2), "batch": B, "input_points": n_in,
                "decoder_points": n_gt, "params_m": round(n_params / 1e6, 2), "steps": steps, "final_loss": round(last[0], 4),
                "model": "model_zoo/completion/inpainter.py (reference file, unmodified) through dropin/",
                "loop": "train_inpainter.py:176-196: host partial_postproces, EMD (0.005, 50 iterations) + Chamfer, Adam"}
    except Exception as exc:  # noqa: BLE001
        import traceback
        return {"unavailable": repr(exc)[:300], "trace": traceback.format_exc()[-600:]}


def run_s3dis_train(dev, steps=20):
    """BASELINE config 3 (S3DIS 1x1): the reference's own model_zoo/s3dis/segmenter.py (9.22 M parameters) and the step
    of train_segmentation.py:177-200 -- clouds of 4096 points with xyz + rgb from pinned host memory, per-point cross
    entropy over 13 classes, Adam 1e-3, the loss and the arg-max predictions read back every step -- through dropin/,
    batch 8 as configs/s3dis.yaml, this rank only."""
    import torch
    try:
        torch.manual_seed(42)
        generator, root = load_reference_model_through_dropin("model_zoo/s3dis/segmenter.py")
        n_params = sum(p.numel() for p in generator.parameters())
        generator = generator.to(dev).train()
        optimizer = torch.optim.Adam(generator.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=0.0)
        ce = torch.nn.CrossEntropyLoss()
        B, N = 8, 4096
        g = torch.Generator().manual_seed(4)
        clouds = [torch.cat([torch.rand(B, N, 2, generator=g), 3 * torch.rand(B, N, 1, generator=g),
                             torch.rand(B, N, 3, generator=g)], dim=2).pin_memory() for _ in range(4)]
        labels = [torch.randint(0, 13, (B, N), generator=g).pin_memory() for _ in range(4)]
        last = [None]

        def step(i):
            pcd = clouds[i % 4].to(dev, non_blocking=True).permute(0, 2, 1)[:, :, None]
            lab = labels[i % 4].to(dev, non_blocking=True)
            pred, _ = generator(pcd)
            loss = ce(pred[:, :, 0], lab)
            loss.backward()
            optimizer.step()
            optimizer.zero_grad()
            last[0] = loss.item()
            pred[:, :, 0].detach().cpu().numpy().argmax(1)          # confusion-matrix update of the script (:198-200)

        for i in range(3):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"samples_per_s": round(B / (ms * 1e-3), 2), "ms_per_step": round(ms, 2), "batch": B, "points": N,
                "params_m": round(n_params / 1e6, 2), "steps": steps, "final_loss": round(last[0], 4),
                "model": "model_zoo/s3dis/segmenter.py (reference file, unmodified) through dropin/",
                "loop": "train_segmentation.py:177-200: per-point cross entropy, Adam, loss and predictions read back every step"}
    except Exception as exc:  # noqa: BLE001
        import traceback
        return {"unavailable": repr(exc)[:300], "trace": traceback.format_exc()[-600:]}


def run_other_configs(args, dev):
    """The other BASELINE.json configs (S3DIS block, completion decoder, two sweep corners) and the deterministic
    mode of the main workload: ms per fwd+bwd pass and Gpt-heads/s, C-ABI calls, inputs resident, this rank only."""
    import torch
    from cloud_transformers_b200.hotpath import HotPath, algorithmic_bytes
    peak, _ = peak_hbm()
    out = {}
    gen = torch.Generator(device=dev).manual_seed(5)
    for name, dim, W, F, N, B, heads in OTHER_CONFIGS:
        try:
            keys = torch.tanh(torch.randn(B, heads * dim, N, generator=gen, device=dev))
            feat = torch.randn(B, heads * F, N, generator=gen, device=dev)
            grid = (B, heads * F) + (W,) * dim
            conv = torch.randn(grid, generator=gen, device=dev)
            gz = torch.randn(grid, generator=gen, device=dev)
            go = torch.randn(B, heads * F, N, generator=gen, device=dev)
            hp = HotPath(W, heads, dim, B, F, N, dev, mode=args.mode)
            for _ in range(2):
                hp.fwd_bwd(keys, feat, conv, go, gz)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                flush_l2(dev)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                hp.fwd_bwd(keys, feat, conv, go, gz)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = statistics.median(ts)
            nbytes = algorithmic_bytes(N, dim, F, W ** dim)["total"] * B * heads
            modes = {0: "atomic", 1: "deterministic", 2: "tile"}
            out[name] = {"shape": {"dim": dim, "W": W, "F": F, "N": N, "B": B, "H": heads}, "ms": round(ms, 4),
                         "gpt_heads_per_s": round(B * heads * N / ms / 1e6, 4),
                         "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peak, 4),
                         "kernels": [modes[m] for m in hp.modes]}
            del hp, keys, feat, conv, gz, go
        except Exception as exc:  # noqa: BLE001
            out[name] = {"unavailable": repr(exc)[:160]}
        torch.cuda.empty_cache()
    return out


def run_deterministic(args, dev, data, order):
    """The main workload with mode = deterministic (plan-based scatters with a fixed summation order)."""
    import torch
    from cloud_transformers_b200.hotpath import HotPath
    try:
        paths = {name: HotPath(W, H, dim, B_PER_GPU, F, N_PTS, dev, mode="deterministic") for name, dim, W, F in CLASSES}

        def step():
            for name, dim, W, F in order:
                paths[name].fwd_bwd(*data[name])
        step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        return {"value": round(len(order) * B_PER_GPU * H * N_PTS / (ms * 1e-3) / 1e9, 4), "unit": UNIT,
                "ms_per_step": round(ms, 3), "what": "same step, mode = deterministic, this rank only"}
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": repr(exc)[:200]}


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


def run_train(args, dev, world, rank, graphed, fused_syncbn=False):
    """Second half of the BASELINE metric: MHCT training samples/s on the reference's OWN ScanObjectNN classifier
    (model_zoo/scanobject/classifier.py, 24.02 M parameters) running on the B200 kernels through dropin/.  The loop is
    train_classification.py:181-273 on synthetic clouds (configs/scanobjectnn.yaml: Adam 1e-3, StepLR, seg_weight 0.5):
    DDP(SyncBatchNorm.convert_sync_batchnorm(model)) as :107-109, loss = 0.5 CE + 0.5 BCE, optimizer step, then the
    script's per-step bookkeeping -- utils/train_util_distributed.py reduce_loss_dict (:12-34) and the pickled
    all_gather of predictions (:37-77), the reference's own unmodified functions -- and the .item() reads of the losses
    and of the 26 x 3 lattice statistics rank 0 logs (:249-260).  Clouds come from pinned host memory every step.
    graphed=True: forward + backward + optimizer step (SyncBN / DDP collectives included) replay as ONE CUDA graph
    (cloud_transformers_b200/graphed.py); the bookkeeping stays eager, after the replay.
    fused_syncbn=True: cloud_transformers_b200.syncbn.convert_sync_batchnorm instead of torch's -- same layers, same
    state_dict, the per-layer statistics exchange done by the normalisation kernels themselves over NVLink peer memory
    (csrc/ctb_syncbn.cuh) instead of one NCCL all_gather / all_reduce per layer and direction."""
    import torch
    import torch.distributed as dist
    from cloud_transformers_b200.graphed import GraphedTrainStep
    from cloud_transformers_b200.sharding import aggregate_throughput
    from cloud_transformers_b200 import syncbn as ctb_syncbn
    try:
        to_sync = ctb_syncbn.convert_sync_batchnorm if fused_syncbn else torch.nn.SyncBatchNorm.convert_sync_batchnorm
        torch.manual_seed(42)                          # train_classification.py:96
        torch.backends.cudnn.benchmark = True          # :58
        B = args.train_batch
        model, root = load_reference_model_through_dropin("model_zoo/scanobject/classifier.py")
        import utils.train_util_distributed as tud     # the reference's own helpers
        n_params = sum(p.numel() for p in model.parameters())
        model = model.to(dev)
        side = torch.cuda.Stream(device=dev)
        if world > 1:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):              # (the graph recipe wants DDP built on the warm-up stream)
                model = torch.nn.parallel.DistributedDataParallel(to_sync(model), device_ids=[dev.index],
                                                                  output_device=dev.index)
            torch.cuda.current_stream().wait_stream(side)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=0, capturable=graphed)
        sched = torch.optim.lr_scheduler.StepLR(opt, gamma=0.7, step_size=25000)
        ce, bce = torch.nn.CrossEntropyLoss(), torch.nn.BCEWithLogitsLoss()
        gen = torch.Generator(device=dev).manual_seed(7 + rank)
        clouds = [surface_clouds(gen, B, N_PTS, dev).permute(0, 2, 1).contiguous().cpu().pin_memory() for _ in range(4)]
        labels = [torch.randint(0, 15, (B,)).pin_memory() for _ in range(4)]
        masks = [(torch.rand(B, N_PTS) < 0.7).float().pin_memory() for _ in range(4)]
        sink = []
        model.train()

        def loss_fn(outs, mask, y):
            class_pred, mask_pred, lattices = outs
            seg_loss = bce(mask_pred[:, 0, 0], mask)
            cls_loss = ce(class_pred, y)
            return 0.5 * cls_loss + 0.5 * seg_loss, (cls_loss, seg_loss)

        def batch(i):
            pcd = clouds[i % 4].to(dev, non_blocking=True).permute(0, 2, 1)[:, :, None]     # [B,3,1,N], :195
            return pcd, masks[i % 4].to(dev, non_blocking=True), labels[i % 4].to(dev, non_blocking=True)

        runner = None
        if graphed:
            pcd, mask, y = batch(0)
            runner = GraphedTrainStep(model, opt, loss_fn, (pcd.contiguous(), mask, y))

        def step(i):
            pcd, mask, y = batch(i)
            if runner is not None:
                (class_pred, mask_pred, lattices), loss, (cls_loss, seg_loss) = runner(pcd, mask, y)
            else:
                outs = model(pcd)
                class_pred, mask_pred, lattices = outs
                loss, (cls_loss, seg_loss) = loss_fn(outs, mask, y)
                loss.backward()
                opt.step()
                opt.zero_grad()
            if world > 1:                                                                     # gather_results, :157-174
                loss_all = tud.reduce_loss_dict({"loss": loss.detach(), "loss_cls": cls_loss.detach(),
                                                 "loss_seg": seg_loss.detach()})
            else:
                loss_all = {"loss": loss, "loss_cls": cls_loss, "loss_seg": seg_loss}
            with torch.no_grad():
                pred_np = class_pred.detach().cpu().numpy().argmax(1)
                mpred_np = (torch.sigmoid(mask_pred[:, 0, 0]) > 0.5).detach().cpu().numpy()
                mask_np, y_np = mask.detach().cpu().numpy(), y.detach().cpu().numpy()
            gathered = tud.all_gather((pred_np, mpred_np, y_np, mask_np)) if world > 1 else [(pred_np, mpred_np, y_np, mask_np)]
            if rank == 0:                                                                     # tensorboard scalars, :249-260
                sink.append([loss_all[k].item() for k in loss_all] +
                            [float(v[0]) + v[1].item() + v[2].item() for v in lattices] + [len(gathered)])
            sched.step()

        for i in range(6 if world > 1 else 4):      # (DDP rebuilds its buckets after the first step; cudnn.benchmark)
            step(i)
        _barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.train_steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        units, ms = aggregate_throughput(B, e0.elapsed_time(e1) / args.train_steps, dev)
        return {"samples_per_s": round(units / (ms * 1e-3), 2), "ms_per_step": round(ms, 2), "batch_per_gpu": B,
                "points": N_PTS, "steps": args.train_steps, "params_m": round(n_params / 1e6, 2),
                "final_loss": round(sink[-1][0], 4) if sink else None,
                "model": "model_zoo/scanobject/classifier.py (reference file, unmodified) through dropin/",
                "loop": "train_classification.py:181-273: DDP + SyncBN (:107-109), Adam + StepLR, reduce_loss_dict + pickled "
                        "all_gather + .item() logging every step",
                "step_execution": "one CUDA graph per step (forward + backward + optimizer + SyncBN / DDP collectives), "
                                  "bookkeeping eager" if graphed else "eager, as the script",
                "syncbn": None if world == 1 else ("cloud_transformers_b200.syncbn: statistics exchanged by the "
                                                   "normalisation kernels over NVLink peer memory" if fused_syncbn
                                                   else "torch.nn.SyncBatchNorm (NCCL collective per layer and direction)"),
                "parallelism": "dp%d (DDP over NCCL + SyncBN)" % world if world > 1 else "single GPU"}
    except Exception as exc:
        import traceback
        return {"unavailable": repr(exc)[:300], "trace": traceback.format_exc()[-900:]}


# ---- the reference's own CPU implementation of the path (cpu_baseline leg and --impl reference) ------------------
_REF_MODS = {}


def _reference_modules():
    """layers/cloud_transform.py of the reference, unmodified, imported from the staged tree under the two dependency
    shims (torch_scatter.scatter_max, pytorch3d so3) of oracle/reference_loader.py."""
    if not _REF_MODS:
        from oracle import reference_loader as RL
        if RL.available():
            ct, _, _ = RL.load_reference_layers()
            _REF_MODS["ct"] = ct
            _REF_MODS["kind"] = "reference"
        else:
            from oracle import ct_torch as T
            _REF_MODS["port"] = T
            _REF_MODS["kind"] = "port"
    return _REF_MODS


def cpu_step(sample_batch, threads, repeats_per_class=1):
    """One bounded sample of the workload on the host: the six shape classes at batch `sample_batch` through the
    reference's own DifferentiablePositions / Splat / Slice modules, forward and backward."""
    import torch
    mods = _reference_modules()
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    t0 = time.perf_counter()
    for name, dim, W, F in CLASSES:
        keys = torch.tanh(torch.randn(sample_batch, H * dim, N_PTS, generator=g))
        feat = torch.randn(sample_batch, H * F, N_PTS, generator=g)
        for _ in range(repeats_per_class):
            if "ct" in mods:
                ct = mods["ct"]
                k, f = keys.clone().requires_grad_(True), feat.clone().requires_grad_(True)
                lc, idx = ct.DifferentiablePositions(tensor_size=W, heads=H, dim=dim)(k)
                z = ct.Splat(tensor_size=W, heads=H, dim=dim)(lc, idx, f)
                out = ct.Slice(tensor_size=W, heads=H, dim=dim)(lc, idx, z)
                out.square().mean().backward()
            else:
                mods["port"].hot_path_fwd_bwd(keys, feat, W, H, dim)
    dt = time.perf_counter() - t0
    return len(CLASSES) * repeats_per_class * sample_batch * H * N_PTS / dt / 1e9, dt


CPU_SAMPLE_BATCH = 8      # clouds per class in one CPU sample (the config's batch is 32; the metric is per point-head)


def cpu_baseline(sample_batch=CPU_SAMPLE_BATCH, repeats=2):
    cores = os.cpu_count() or 1
    cpu_step(sample_batch, cores)                     # warm
    vals = [cpu_step(sample_batch, cores) for _ in range(repeats)]
    best = max(v for v, _ in vals)
    kind = _reference_modules()["kind"]
    return {"value": round(best, 8), "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "the six shape classes once each at B=%d of the config's 32 (H=16, N=2048), %s, "
                      "%d threads, %.1f s per pass" % (sample_batch,
                                                      "the reference's own layers/cloud_transform.py (oracle/_ref)"
                                                      if kind == "reference" else "torch port oracle/ct_torch.py",
                                                      cores, vals[-1][1])}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (its unmodified layers/cloud_transform.py
    from the staged tree), all host threads, on a bounded sample of the SAME workload config: per step the six shape
    classes once each at B=8 of the 32 clouds (the metric is per point-head, so the sample size cancels)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        cpu_step(CPU_SAMPLE_BATCH, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(CPU_SAMPLE_BATCH, cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = len(CLASSES) * CPU_SAMPLE_BATCH * H * N_PTS / dt / 1e9
    kind = _reference_modules()["kind"]
    line = {
        "impl": "reference", "metric": METRIC, "value": round(val, 8), "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.mode),
        "cpu_baseline": {"value": round(val, 8), "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "per step: the six shape classes once each at B=%d of the config's 32 clouds, the "
                                   "reference's own DifferentiablePositions / Splat / Slice forward + backward (scatter_max "
                                   "through the shim of oracle/reference_loader.py)" % CPU_SAMPLE_BATCH},
        "e2e": {"value": round(val, 8), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--train-steps", type=int, default=100, help="steps of the MHCT training throughput add-on (0 = skip)")
    ap.add_argument("--train-batch", type=int, default=B_PER_GPU)
    ap.add_argument("--grid-dtype", default="f32", choices=["f32", "bf16"],
                    help="bf16: grids (z, convolved, grad_grid, grad_z) stored as bf16, arithmetic stays fp32")
    ap.add_argument("--mode", default=os.environ.get("CTB_MODE", "auto"), choices=["auto", "atomic", "tile", "deterministic"])
    ap.add_argument("--skip", default="", help="developer switch: comma list of add-on sections to skip "
                                               "(e2e, refgpu, det, others, losses, completion, s3dis, train, eager, fusedbn, cpu, bf16)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
