"""Time the four hot-path ops on arbitrary shapes in several modes (CUDA events, inputs resident).
usage: python tools/time_shapes.py  [dim,W,F,N,B ...]   (default: the S3DIS and completion shapes of SURVEY 8(d) C3 / C4)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloud_transformers_b200.hotpath import HotPath, algorithmic_bytes  # noqa: E402

H = 16
shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [
    (3, 32, 4, 4096, 8), (2, 128, 4, 4096, 8), (3, 16, 16, 4096, 8), (3, 8, 32, 4096, 8),
    (2, 128, 4, 16384, 2), (3, 32, 4, 16384, 2), (2, 64, 16, 16384, 2), (3, 16, 16, 16384, 2)]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for dim, W, F, N, B in shapes:
    keys = torch.tanh(torch.randn(B, H * dim, N, generator=g, device=dev))
    feat = torch.randn(B, H * F, N, generator=g, device=dev)
    grid = (B, H * F) + (W,) * dim
    conv = torch.randn(grid, generator=g, device=dev)
    gz = torch.randn(grid, generator=g, device=dev)
    go = torch.randn(B, H * F, N, generator=g, device=dev)
    line = []
    for mode in ("auto", "atomic"):
        try:
            hp = HotPath(W, H, dim, B, F, N, dev, mode=mode)
            for _ in range(3):
                hp.fwd_bwd(keys, feat, conv, go, gz)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                hp.fwd_bwd(keys, feat, conv, go, gz)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            by = algorithmic_bytes(N, dim, F, W ** dim)["total"] * B * H
            line.append("%s %.3f ms %.2f Gpt-heads/s %.0f GB/s" % (mode, ms, B * H * N / ms / 1e6, by / ms / 1e6))
        except Exception as exc:  # noqa: BLE001
            line.append("%s failed: %s" % (mode, str(exc)[:80]))
    print("d=%d W=%d F=%d N=%d B=%d | %s" % (dim, W, F, N, B, " | ".join(line)))
