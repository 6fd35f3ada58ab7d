"""Per-phase cycle shares of tile_scatter_kernel (developer tool).  Needs a library built with -DCTB_PHASE_TIMERS:
    nvcc <flags of __graft_entry__.NVCC_FLAGS> -DCTB_PHASE_TIMERS csrc/ctb200.cu -o cloud_transformers_b200/libctb200.so
usage: python tools/phase_timers.py [classes]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cloud_transformers_b200 import _lib  # noqa: E402
from cloud_transformers_b200.hotpath import HotPath  # noqa: E402

lib = _lib.load()
fn = lib.ctb_debug_phase
fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
names = ["init", "compact", "prepass", "pt loads", "max pass", "arg pass", "to fence", "tma store", "pre-fence", "fence",
         "E:staging", "E:tile init", "E:wait for slowest window", "E:fold", "E:write-out", "E:windows (thread 0)"]
classes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["a2d", "a3d"]
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(42)
for name, dim, W, F in bench.CLASSES:
    if name not in classes:
        continue
    keys, feat, conv, go, gz = bench.make_class_inputs(gen, dim, W, F, bench.B_PER_GPU, dev)
    hp = HotPath(W, bench.H, dim, bench.B_PER_GPU, F, bench.N_PTS, dev, mode="auto")
    hp.fwd_bwd(keys, feat, conv, go, gz)
    buf = (ctypes.c_ulonglong * 16)()
    for op in ("splat_fwd", "slice_bwd"):
        fn(buf, 1)
        if op == "splat_fwd":
            hp.splat_fwd_only(keys, feat)
        else:
            hp.slice_bwd(keys, conv, go)
        fn(buf, 0)
        tot = float(sum(buf[:16])) or 1.0
        print(name, op, " ".join("%s=%.0f%%" % (n, 100.0 * buf[i] / tot) for i, n in enumerate(names) if buf[i]),
              "cycles/CTA-sum", int(tot))
