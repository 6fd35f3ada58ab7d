"""Allocation-free runner of the fused hot path over the C ABI.

`HotPath` owns the output / workspace buffers of one shape class and issues, per call,
    plan build (deterministic mode) + Splat forward      ctb_plan_build, ctb_splat_fwd_keys
    Slice forward                                         ctb_slice_fwd_keys
    Slice backward (grad_grid, grad_keys)                 ctb_slice_bwd_keys
    Splat backward (grad_features, grad_keys)             ctb_splat_bwd_keys
on the current stream with no host synchronisation and no allocation -- the four passes an MHCT block
(layers/multihead_ct.py:99-107) drives through Splat / Slice in training.  Used by bench.py for the
device-resident throughput figure and usable for CUDA-graph capture.
"""
import ctypes

import torch

from . import _lib
from .functional import Geometry, _call, _ptr, _stream


class HotPath:
    def __init__(self, tensor_size, heads, dim, B, F, N, device, mode="auto", reduce="max", grid_dtype=torch.float32):
        sizes = [tensor_size] * dim if isinstance(tensor_size, int) else list(tensor_size)
        self.geom = Geometry(sizes, heads, dim)
        self.B, self.F, self.N = B, F, N
        self.device = torch.device(device)
        self.reduce = _lib.REDUCE_MAX if reduce == "max" else _lib.REDUCE_SUM
        self.grid_dtype = grid_dtype
        if grid_dtype == torch.bfloat16 and mode not in ("auto", "tile"):
            raise ValueError("bf16 grid storage needs the tile kernels")
        self.shape = self.geom.shape(B, F, N, _lib.DTYPE_BF16 if grid_dtype == torch.bfloat16 else _lib.DTYPE_F32)
        lib = _lib.load()
        want = {"auto": _lib.MODE_TILE, "tile": _lib.MODE_TILE, "deterministic": _lib.MODE_DETERMINISTIC,
                "atomic": _lib.MODE_ATOMIC}[mode]
        sup = [lib.ctb_mode_supported(ctypes.byref(self.shape), op, self.reduce, want) for op in range(4)]
        if mode in ("tile", "deterministic") and not all(sup):
            raise _lib.CtbError("ctb_mode_supported", _lib.CTB_ERR_UNSUPPORTED, "shape not covered")
        self.modes = [want if s else _lib.MODE_ATOMIC for s in sup]
        H, C = heads, self.geom.C
        f32 = dict(dtype=torch.float32, device=self.device)
        if grid_dtype == torch.bfloat16 and not all(sup):
            raise _lib.CtbError("ctb_mode_supported", _lib.CTB_ERR_UNSUPPORTED, "bf16 grids need a tile-supported shape")
        self.z = torch.empty((B, H * F) + self.geom.sizes, dtype=grid_dtype, device=self.device)
        self.arg = torch.empty((B, H * F, C), dtype=torch.int32, device=self.device)
        self.out = torch.empty((B, H * F, N), **f32)
        self.grad_grid = torch.empty((B, H * F) + self.geom.sizes, dtype=grid_dtype, device=self.device)
        self.grad_keys_slice = torch.empty((B, H * dim, N), **f32)
        self.grad_keys_splat = torch.empty((B, H * dim, N), **f32)
        self.grad_feat = torch.empty((B, H * F, N), **f32)
        self.plan = None
        if any(m != _lib.MODE_ATOMIC and lib.ctb_plan_used(ctypes.byref(self.shape), m)
               for m in (self.modes[_lib.OP_SPLAT_FWD], self.modes[_lib.OP_SLICE_BWD])):
            nbytes = lib.ctb_plan_bytes(ctypes.byref(self.shape))
            self.plan = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._sh = ctypes.byref(self.shape)
        # Where only the backward reads the plan (16^3 x F16: forward max = tile scatter, grad_grid sum = plan-based), it
        # is built on a side stream under the two forward passes instead of in front of them.
        self.plan_in_fwd = self.plan is not None and bool(lib.ctb_op_uses_plan(
            self._sh, _lib.OP_SPLAT_FWD, self.reduce, self.modes[_lib.OP_SPLAT_FWD]))
        self.slice_fwd_uses_plan = self.plan is not None and bool(lib.ctb_op_uses_plan(
            self._sh, _lib.OP_SLICE_FWD, self.reduce, self.modes[_lib.OP_SLICE_FWD]))
        self._side = None

    # number of kernel launches (ours) per full fwd+bwd pass, for bench.py's gpu_launches claim
    def launches_per_pass(self):
        m_sf, m_sb = self.modes[_lib.OP_SPLAT_FWD], self.modes[_lib.OP_SLICE_BWD]
        n = 1 if self.plan is not None else 0                             # plan build
        n += 1 if m_sf != _lib.MODE_ATOMIC else (2 if self.reduce == 0 else 1)   # splat fwd
        n += 1                                                            # slice fwd
        n += 1 if m_sb == _lib.MODE_ATOMIC else 2                         # slice bwd: scatter + gather
        n += 1                                                            # splat bwd
        return n

    def _call(self, name, *args):
        # kernels, stream and pointers must all belong to self.device, whatever the caller's current device is
        with torch.cuda.device(self.device):
            _call(name, *args)

    def build_plan(self, keys):
        self._call("ctb_plan_build", _ptr(keys), _ptr(self.plan), ctypes.c_size_t(self.plan.numel()), self._sh,
              _stream(keys))

    def splat_fwd(self, keys, feat, pad=None):
        if self.plan is not None:
            self.build_plan(keys)
        return self.splat_fwd_only(keys, feat, pad)

    def splat_fwd_only(self, keys, feat, pad=None):
        self._call("ctb_splat_fwd_keys", _ptr(keys), _ptr(feat), _ptr(pad), _ptr(self.z), _ptr(self.arg), self._sh,
              self.reduce, self.modes[_lib.OP_SPLAT_FWD], _ptr(self.plan), _stream(keys))
        return self.z

    def slice_fwd(self, keys, grid, pad=None):
        self._call("ctb_slice_fwd_keys", _ptr(keys), _ptr(grid), _ptr(pad), _ptr(self.out), self._sh,
              self.modes[_lib.OP_SLICE_FWD], _ptr(self.plan) if self.slice_fwd_uses_plan else None, _stream(keys))
        return self.out

    def slice_bwd(self, keys, grid, grad_out, pad=None):
        self._call("ctb_slice_bwd_keys", _ptr(keys), _ptr(grid), _ptr(pad), _ptr(grad_out), _ptr(self.grad_grid),
              _ptr(self.grad_keys_slice), self._sh, self.modes[_lib.OP_SLICE_BWD], _ptr(self.plan), _stream(keys))
        return self.grad_grid, self.grad_keys_slice

    def splat_bwd(self, keys, feat, grad_z, pad=None):
        self._call("ctb_splat_bwd_keys", _ptr(keys), _ptr(feat), _ptr(pad), _ptr(grad_z), _ptr(self.arg),
              _ptr(self.grad_feat), _ptr(self.grad_keys_splat), self._sh, self.reduce,
              self.modes[_lib.OP_SPLAT_BWD], _stream(keys))
        return self.grad_feat, self.grad_keys_splat

    def fwd_bwd(self, keys, feat, conv, grad_out, grad_z, pad=None):
        """One pass of the hot path; `conv` stands for the convolved grid (the conv itself is outside the
        metric), `grad_z` for the gradient the conv backward hands to Splat."""
        if self.plan is not None and not self.plan_in_fwd:
            cur = torch.cuda.current_stream(self.device)
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self.build_plan(keys)
            self.splat_fwd_only(keys, feat, pad)
            if self.slice_fwd_uses_plan:
                cur.wait_stream(self._side)
            self.slice_fwd(keys, conv, pad)
            if not self.slice_fwd_uses_plan:
                cur.wait_stream(self._side)
        else:
            self.splat_fwd(keys, feat, pad)
            self.slice_fwd(keys, conv, pad)
        self.slice_bwd(keys, conv, grad_out, pad)
        self.splat_bwd(keys, feat, grad_z, pad)


def algorithmic_bytes(N, dim, F, C, e=4, e_grid=None):
    """minimal HBM bytes per (batch, head) unit, per pass and in total (reduce = max).
    e = bytes per feature element, e_grid = bytes per grid element (bf16 grid storage mode: 2)."""
    d = dim
    eg = e if e_grid is None else e_grid
    per = {
        "splat_fwd": N * (4 * d + e * F) + eg * F * C,
        "slice_fwd": N * (4 * d + e * F) + eg * F * C,
        "slice_bwd": N * (8 * d + e * F) + 2 * eg * F * C,
        "splat_bwd": N * (8 * d + 2 * e * F) + 2 * eg * F * C,
    }
    per["total"] = sum(per.values())
    assert per["total"] == N * (24 * d + 5 * e * F) + 6 * eg * F * C
    return per
