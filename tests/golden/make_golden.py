"""Mint the golden fixtures for the Splat / Slice path by running the UNMODIFIED reference
(/root/reference/layers/cloud_transform.py + layers/utils.py, imported under the two dependency shims of
oracle/reference_loader.py) on seeded synthetic inputs.  Build-container only; the resulting .npz files
are committed and are what the CPU and GPU test suites compare against.

    python tests/golden/make_golden.py

The reference ships no tests or vectors of its own, so these are the pins.
Inputs are continuous random values (no exact ties / zeros among the features) except for the explicit
adversarial key sweep, because tie behaviour of scatter_max is implementation-defined.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_loader as RL  # noqa: E402

CASES = [
    # name, dim, tensor_size, heads, F, N, B, pad
    ("c2d_w8", 2, 8, 2, 3, 40, 2, False),
    ("c2d_w16_pad", 2, 16, 3, 4, 64, 2, True),
    ("c3d_w4", 3, 4, 2, 2, 33, 2, False),
    ("c3d_w8_pad", 3, 8, 2, 4, 48, 1, True),
    ("c2d_rect", 2, (6, 10), 2, 2, 37, 1, False),
    ("c3d_rect", 3, (4, 6, 5), 2, 3, 29, 2, True),
]


def adversarial_keys():
    vals = [0.0, -0.0, 1.0, -1.0, 2.0, -2.0, 1e-8, -1e-8, 0.5, -0.5]
    one = np.float32(1.0)
    for k in range(1, 4):
        vals += [float(np.nextafter(one, np.float32(0), dtype=np.float32) if k == 1 else 1 - k * 2.0 ** -24)]
        vals += [-(1 - k * 2.0 ** -24), 1 - k * 2.0 ** -23, -(1 - k * 2.0 ** -23)]
    for W in (8, 16, 32, 64, 128, 256):
        for j in range(0, W, max(1, W // 8)):
            b = 2.0 * j / (W - 1) - 1.0          # exact cell boundary in key space
            b32 = np.float32(b)
            vals += [float(b32), float(np.nextafter(b32, np.float32(2))), float(np.nextafter(b32, np.float32(-2)))]
    return np.asarray(vals, dtype=np.float32)


def main():
    ct, ut, mh = RL.load_reference_layers()
    torch.manual_seed(1234)
    out = {}
    for name, dim, W, H, F, N, B, use_pad in CASES:
        keys = torch.tanh(torch.randn(B, H * dim, N) * 1.3)
        feat = torch.randn(B, H * F, N)
        pad = (torch.rand(B, N) > 0.25).float() if use_pad else None
        dp = ct.DifferentiablePositions(tensor_size=W, heads=H, dim=dim)
        sp = ct.Splat(tensor_size=W, heads=H, dim=dim)
        sl = ct.Slice(tensor_size=W, heads=H, dim=dim)
        k = keys.clone().requires_grad_(True)
        f = feat.clone().requires_grad_(True)
        lc, idx = dp(k)
        z = sp(lc, idx, f, pad)
        conv = torch.randn_like(z).requires_grad_(True)
        o = sl(lc, idx, conv, pad)
        go = torch.randn_like(o)
        gz = torch.randn_like(z)
        (o * go).sum().backward(retain_graph=True)
        gk_slice = k.grad.clone()
        k.grad = None
        (z * gz).sum().backward()
        d = dict(keys=keys, feat=feat, lc=lc.detach(), idx=idx.to(torch.int32), z=z.detach(), conv=conv.detach(),
                 out=o.detach(), go=go, gz=gz, gconv=conv.grad, gk_slice=gk_slice, gk_splat=k.grad, gfeat=f.grad)
        if pad is not None:
            d["pad"] = pad
        for kk, v in d.items():
            out["%s/%s" % (name, kk)] = v.numpy()
    # adversarial index sweep: 1 head, the same key on every axis
    ak = adversarial_keys()
    out["adv/keys"] = ak
    for dim in (2, 3):
        for W in ((8, 16, 32, 64, 128, 256) if dim == 2 else (8, 16, 32, 64)):
            keys = torch.from_numpy(np.tile(ak[None, None, :], (1, dim, 1)).copy())
            dp = ct.DifferentiablePositions(tensor_size=W, heads=1, dim=dim)
            lc, idx = dp(keys)
            out["adv/idx_d%d_w%d" % (dim, W)] = idx.to(torch.int32).numpy()
            out["adv/lc_d%d_w%d" % (dim, W)] = lc.numpy()
    path = os.path.join(HERE, "splat_slice_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
