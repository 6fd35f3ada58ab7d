"""torch.autograd.Functions over the C ABI (include/ctb200.h).  PyTorch is plumbing here: it owns the
device memory and the stream; all arithmetic happens in libctb200.so.

Two families:
  * reference-API ops (`positions`, `splat`, `slice_`) take / return the reference's tensors
    (local_coordinate, flattened_index), layers/cloud_transform.py:72-121, :131-180, :190-227;
  * fused ops (`fused_splat`, `fused_slice`) take the keys through a `PositionsHandle`, never read
    lc / idx from memory and return grad_keys directly (SURVEY.md rows A1-A7 in four kernels).
"""
import ctypes
import os

import torch
from torch.autograd.function import once_differentiable

from . import _lib

_MODES = {"auto": None, "atomic": _lib.MODE_ATOMIC, "tile": _lib.MODE_TILE, "deterministic": _lib.MODE_DETERMINISTIC}


class _Config:
    """Process-wide switches.  mode: 'auto' (shared-memory tile kernels when the shape is supported, else
    the point-stationary L2-atomic kernels), 'atomic', 'tile' or 'deterministic' (error if unsupported).
    fused: let Splat / Slice bypass lc / idx when they were produced by our DifferentiablePositions.
    use_plan: see below."""

    def __init__(self):
        self.mode = os.environ.get("CTB_MODE", "auto")
        self.fused = os.environ.get("CTB_FUSED", "1") != "0"
        # tile mode: build the per-unit plan (sorted entry lists) where the plan-based scatters are faster; off = always
        # the shared-memory tile scatters (fixed-point sums, invariant under permutations of the points)
        self.use_plan = os.environ.get("CTB_USE_PLAN", "1") != "0"


config = _Config()


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cloud_transformers_b200 runs on CUDA tensors only (sm_100a); there is no CPU "
                               "fallback. Got a tensor on %s." % t.device)


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class Geometry:
    """tensor_size / heads / dim of a DifferentiableGridModule (cloud_transform.py:29-59)."""

    __slots__ = ("sizes", "heads", "dim", "S", "C")

    def __init__(self, sizes, heads, dim):
        self.sizes = tuple(int(s) for s in sizes)
        self.heads = int(heads)
        self.dim = int(dim)
        assert self.dim in (2, 3) and len(self.sizes) == self.dim
        self.S = 1 << self.dim
        c = 1
        for s in self.sizes:
            c *= s
        self.C = c

    def shape(self, B, F, N, grid_dtype=0):
        return _lib.make_shape(B, self.heads, F, N, self.dim, self.sizes, grid_dtype)


def _call(name, *args):
    lib = _lib.load()
    _lib.check(name, getattr(lib, name)(*args))


# ---------------------------------------------------------------------------------------------------
# reference-API ops
class _PositionsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, keys, geom):
        _require_cuda(keys)
        k = _f32c(keys)
        B, HD, N = k.shape
        assert HD == geom.heads * geom.dim  # cloud_transform.py:84
        lc = torch.empty((B, geom.heads, geom.S, N), dtype=torch.float32, device=k.device)
        idx = torch.empty((B, geom.heads, geom.S, N), dtype=torch.int64, device=k.device)
        with torch.cuda.device(k.device):
            _call("ctb_positions_fwd", _ptr(k), _ptr(lc), _ptr(idx), ctypes.byref(geom.shape(B, 1, N)), _stream(k))
        ctx.save_for_backward(k)
        ctx.geom = geom
        ctx.mark_non_differentiable(idx)
        return lc, idx

    @staticmethod
    @once_differentiable
    def backward(ctx, g_lc, _g_idx):
        (k,) = ctx.saved_tensors
        geom = ctx.geom
        B, _, N = k.shape
        g = _f32c(g_lc)
        gk = torch.empty_like(k)
        with torch.cuda.device(k.device):
            _call("ctb_positions_bwd", _ptr(k), _ptr(g), _ptr(gk), ctypes.byref(geom.shape(B, 1, N)), _stream(k))
        return gk, None


class _SplatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lc, idx, features, pad, geom, reduce):
        _require_cuda(lc, idx, features, pad)
        lc_c, f_c, p_c = _f32c(lc), _f32c(features), _f32c(pad)
        idx_c = idx.contiguous()
        assert idx_c.dtype == torch.int64
        B, N = f_c.size(0), f_c.size(-1)
        F = f_c.size(1) // geom.heads
        z = torch.empty((B, geom.heads * F) + geom.sizes, dtype=torch.float32, device=f_c.device)
        arg = torch.empty((B, geom.heads * F, geom.C), dtype=torch.int32, device=f_c.device) \
            if reduce == _lib.REDUCE_MAX else None
        with torch.cuda.device(f_c.device):
            _call("ctb_splat_fwd", _ptr(lc_c), _ptr(idx_c), _ptr(f_c), _ptr(p_c), _ptr(z), _ptr(arg),
                  ctypes.byref(geom.shape(B, F, N)), reduce, _stream(f_c))
        ctx.save_for_backward(lc_c, idx_c, f_c, p_c, arg)
        ctx.geom, ctx.reduce = geom, reduce
        return z

    @staticmethod
    @once_differentiable
    def backward(ctx, gz):
        lc_c, idx_c, f_c, p_c, arg = ctx.saved_tensors
        geom = ctx.geom
        B, N = f_c.size(0), f_c.size(-1)
        F = f_c.size(1) // geom.heads
        gz_c = _f32c(gz)
        gf = torch.empty_like(f_c)
        glc = torch.empty_like(lc_c)
        with torch.cuda.device(f_c.device):
            _call("ctb_splat_bwd", _ptr(lc_c), _ptr(idx_c), _ptr(f_c), _ptr(p_c), _ptr(gz_c), _ptr(arg), _ptr(gf),
                  _ptr(glc), ctypes.byref(geom.shape(B, F, N)), ctx.reduce, _stream(f_c))
        return glc, None, gf, None, None, None


class _SliceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lc, idx, grid, pad, geom):
        _require_cuda(lc, idx, grid, pad)
        lc_c, g_c, p_c = _f32c(lc), _f32c(grid), _f32c(pad)
        idx_c = idx.contiguous()
        B, N = lc_c.size(0), lc_c.size(-1)
        F = g_c.size(1) // geom.heads
        out = torch.empty((B, geom.heads * F, N), dtype=torch.float32, device=g_c.device)
        with torch.cuda.device(g_c.device):
            _call("ctb_slice_fwd", _ptr(lc_c), _ptr(idx_c), _ptr(g_c), _ptr(p_c), _ptr(out),
                  ctypes.byref(geom.shape(B, F, N)), _stream(g_c))
        ctx.save_for_backward(lc_c, idx_c, g_c, p_c)
        ctx.geom = geom
        ctx.grid_dtype = grid.dtype
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, go):
        lc_c, idx_c, g_c, p_c = ctx.saved_tensors
        geom = ctx.geom
        B, N = lc_c.size(0), lc_c.size(-1)
        F = g_c.size(1) // geom.heads
        go_c = _f32c(go)
        gg = torch.empty_like(g_c)
        glc = torch.empty_like(lc_c)
        with torch.cuda.device(g_c.device):
            _call("ctb_slice_bwd", _ptr(lc_c), _ptr(idx_c), _ptr(g_c), _ptr(p_c), _ptr(go_c), _ptr(gg), _ptr(glc),
                  ctypes.byref(geom.shape(B, F, N)), _stream(g_c))
        return glc, None, gg.to(ctx.grid_dtype), None, None


def positions(keys, geom):
    return _PositionsFn.apply(keys, geom)


def splat(lc, idx, features, pad, geom, reduce=_lib.REDUCE_MAX):
    return _SplatFn.apply(lc, idx, features, pad, geom, reduce)


def slice_(lc, idx, grid, pad, geom):
    return _SliceFn.apply(lc, idx, grid, pad, geom)


# ---------------------------------------------------------------------------------------------------
# fused ops
class PositionsHandle:
    """What DifferentiablePositions knows about one call: the keys tensor (autograd-connected), the
    geometry, and lazily the cell-sorted plan shared by the Splat forward and the Slice backward."""

    def __init__(self, keys, geom):
        self.keys = keys
        self.geom = geom
        self._keys_c = None
        self._plan = None
        self._plan_event = None

    def keys_c(self):
        if self._keys_c is None:
            self._keys_c = _f32c(self.keys.detach())
        return self._keys_c

    def mode_for(self, op, F, reduce=_lib.REDUCE_MAX):
        want = config.mode
        if want not in _MODES:
            raise ValueError("CTB mode must be one of %s" % sorted(_MODES))
        if want == "atomic":
            return _lib.MODE_ATOMIC
        k = self.keys
        mode = _lib.MODE_DETERMINISTIC if want == "deterministic" else _lib.MODE_TILE
        ok = _lib.load().ctb_mode_supported(ctypes.byref(self.geom.shape(k.size(0), F, k.size(-1))), op, reduce, mode)
        if ok:
            return mode
        if want != "auto":
            raise _lib.CtbError("ctb_mode_supported", _lib.CTB_ERR_UNSUPPORTED,
                                "shape not covered by the %s kernels" % want)
        return _lib.MODE_ATOMIC

    def plan_for(self, mode, F):
        """The plan if the kernels of `mode` use one on this shape (built once, shared by Splat fwd and Slice bwd)."""
        if mode == _lib.MODE_ATOMIC or (mode == _lib.MODE_TILE and not config.use_plan):
            return None
        k = self.keys
        sh = self.geom.shape(k.size(0), F, k.size(-1))
        if not _lib.load().ctb_plan_used(ctypes.byref(sh), mode):
            return None
        return self.plan()

    def _build_plan(self):
        k = self.keys_c()
        B, _, N = k.shape
        sh = self.geom.shape(B, 1, N)
        lib = _lib.load()
        nbytes = lib.ctb_plan_bytes(ctypes.byref(sh))
        if nbytes == 0:
            raise _lib.CtbError("ctb_plan_bytes", _lib.CTB_ERR_UNSUPPORTED, "shape cannot be planned")
        plan = torch.empty(nbytes, dtype=torch.uint8, device=k.device)
        with torch.cuda.device(k.device):
            _call("ctb_plan_build", _ptr(k), _ptr(plan), ctypes.c_size_t(nbytes), ctypes.byref(sh), _stream(k))
        return plan

    def plan(self):
        if self._plan is None:
            self._plan = self._build_plan()
        elif self._plan_event is not None:
            # built ahead on the side stream (prefetch_plan): order this stream behind it, once
            cur = torch.cuda.current_stream(self._plan.device)
            cur.wait_event(self._plan_event)
            self._plan.record_stream(cur)
            self._plan_event = None
        return self._plan

    def prefetch_plan(self, mode, F):
        """Splat forward calls this when the plan is only read later (the grad_grid sum of Slice backward on grids
        whose forward max keeps the tile scatter): build it on a side stream under the forward passes instead of on
        the backward's critical path.  Skipped under CUDA-graph capture (a fork that a forward-only capture never joins)."""
        if self._plan is not None or not config.use_plan or not torch.is_grad_enabled():
            return
        k = self.keys
        if torch.cuda.is_current_stream_capturing():
            return
        sh = self.geom.shape(k.size(0), F, k.size(-1))
        lib = _lib.load()
        if not lib.ctb_plan_used(ctypes.byref(sh), mode) or lib.ctb_op_uses_plan(ctypes.byref(sh), _lib.OP_SPLAT_FWD, 0, mode):
            return
        kc = self.keys_c()
        cur = torch.cuda.current_stream(kc.device)
        side = _side_stream(kc.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self._plan = self._build_plan()
            self._plan_event = torch.cuda.Event()
            self._plan_event.record(side)
        kc.record_stream(side)


_SIDE = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


def _grid_dtype_code(dtype, mode):
    """bf16 grids are handled natively by the TILE / DETERMINISTIC kernels; everything else goes through fp32."""
    return _lib.DTYPE_BF16 if (dtype == torch.bfloat16 and mode != _lib.MODE_ATOMIC) else _lib.DTYPE_F32


def _grid_c(t, code):
    return t.contiguous() if code == _lib.DTYPE_BF16 else _f32c(t)


class _FusedSplatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, keys, features, pad, handle, reduce, out_dtype=None):
        _require_cuda(keys, features, pad)
        geom = handle.geom
        k, f_c, p_c = handle.keys_c(), _f32c(features), _f32c(pad)
        B, N = f_c.size(0), f_c.size(-1)
        F = f_c.size(1) // geom.heads
        mode = handle.mode_for(_lib.OP_SPLAT_FWD, F, reduce)
        uses_plan = mode != _lib.MODE_ATOMIC and _lib.load().ctb_op_uses_plan(
            ctypes.byref(geom.shape(B, F, N)), _lib.OP_SPLAT_FWD, reduce, mode)
        plan = handle.plan_for(mode, F) if uses_plan else None
        gd = _grid_dtype_code(out_dtype, mode)
        ctx.gd = gd
        z = torch.empty((B, geom.heads * F) + geom.sizes,
                        dtype=torch.bfloat16 if gd == _lib.DTYPE_BF16 else torch.float32, device=f_c.device)
        arg = torch.empty((B, geom.heads * F, geom.C), dtype=torch.int32, device=f_c.device) \
            if reduce == _lib.REDUCE_MAX else None
        with torch.cuda.device(f_c.device):
            _call("ctb_splat_fwd_keys", _ptr(k), _ptr(f_c), _ptr(p_c), _ptr(z), _ptr(arg),
                  ctypes.byref(geom.shape(B, F, N, gd)), reduce, mode, _ptr(plan), _stream(f_c))
        ctx.save_for_backward(k, f_c, p_c, arg)
        ctx.handle, ctx.reduce = handle, reduce
        if plan is None and mode == _lib.MODE_TILE and (features.requires_grad or keys.requires_grad):
            handle.prefetch_plan(mode, F)
        if out_dtype is not None and z.dtype != out_dtype:
            z = z.to(out_dtype)
        return z

    @staticmethod
    @once_differentiable
    def backward(ctx, gz):
        k, f_c, p_c, arg = ctx.saved_tensors
        handle = ctx.handle
        geom = handle.geom
        B, N = f_c.size(0), f_c.size(-1)
        F = f_c.size(1) // geom.heads
        mode = handle.mode_for(_lib.OP_SPLAT_BWD, F, ctx.reduce)
        gd = _grid_dtype_code(gz.dtype, mode)
        gz_c = _grid_c(gz, gd)
        gf = torch.empty_like(f_c)
        gk = torch.empty_like(k)
        with torch.cuda.device(f_c.device):
            _call("ctb_splat_bwd_keys", _ptr(k), _ptr(f_c), _ptr(p_c), _ptr(gz_c), _ptr(arg), _ptr(gf), _ptr(gk),
                  ctypes.byref(geom.shape(B, F, N, gd)), ctx.reduce, mode, _stream(f_c))
        return gk, gf, None, None, None, None


class _FusedSliceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, keys, grid, pad, handle):
        _require_cuda(keys, grid, pad)
        geom = handle.geom
        k, p_c = handle.keys_c(), _f32c(pad)
        B, N = k.size(0), k.size(-1)
        F = grid.size(1) // geom.heads
        mode = handle.mode_for(_lib.OP_SLICE_FWD, F)
        gd = _grid_dtype_code(grid.dtype, mode)
        g_c = _grid_c(grid, gd)
        out = torch.empty((B, geom.heads * F, N), dtype=torch.float32, device=g_c.device)
        sh = geom.shape(B, F, N, gd)
        # the sorted-order gather reads the plan where the block has one anyway: already built (Splat forward on the
        # coarse grids), being built on the side stream, or due in this op's backward
        plan = None
        if (mode != _lib.MODE_ATOMIC and config.use_plan
                and _lib.load().ctb_op_uses_plan(ctypes.byref(sh), _lib.OP_SLICE_FWD, 0, mode)
                and (handle._plan is not None
                     or (torch.is_grad_enabled() and (grid.requires_grad or keys.requires_grad)))):
            plan = handle.plan()
        with torch.cuda.device(g_c.device):
            _call("ctb_slice_fwd_keys", _ptr(k), _ptr(g_c), _ptr(p_c), _ptr(out), ctypes.byref(sh), mode, _ptr(plan),
                  _stream(g_c))
        ctx.save_for_backward(k, g_c, p_c)
        ctx.handle = handle
        ctx.grid_dtype = grid.dtype
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, go):
        k, g_c, p_c = ctx.saved_tensors
        handle = ctx.handle
        geom = handle.geom
        B, N = k.size(0), k.size(-1)
        F = g_c.size(1) // geom.heads
        mode = handle.mode_for(_lib.OP_SLICE_BWD, F)
        plan = handle.plan_for(mode, F)
        gd = _grid_dtype_code(g_c.dtype, mode)
        if gd == _lib.DTYPE_F32 and g_c.dtype != torch.float32:
            g_c = g_c.float()
        go_c = _f32c(go)
        gg = torch.empty_like(g_c)
        gk = torch.empty_like(k)
        with torch.cuda.device(g_c.device):
            _call("ctb_slice_bwd_keys", _ptr(k), _ptr(g_c), _ptr(p_c), _ptr(go_c), _ptr(gg), _ptr(gk),
                  ctypes.byref(geom.shape(B, F, N, gd)), mode, _ptr(plan), _stream(g_c))
        return gk, gg.to(ctx.grid_dtype), None, None


def fused_splat(handle, features, pad=None, reduce=_lib.REDUCE_MAX, out_dtype=None):
    """out_dtype=torch.bfloat16 selects the bf16 grid storage mode (fp32 arithmetic, bf16 z)."""
    return _FusedSplatFn.apply(handle.keys, features, pad, handle, reduce, out_dtype)


def fused_slice(handle, grid, pad=None):
    return _FusedSliceFn.apply(handle.keys, grid, pad, handle)


def count_occupied(z):
    """(|z| > 1e-9).sum() as a device tensor (multihead_ct.py:104-105), one pass, no host sync."""
    _require_cuda(z)
    z_c = _f32c(z)
    cnt = torch.zeros(1, dtype=torch.int64, device=z_c.device)
    with torch.cuda.device(z_c.device):
        _call("ctb_count_occupied", _ptr(z_c), ctypes.c_uint64(z_c.numel()), _ptr(cnt), _stream(z_c))
    return cnt[0]


# ---------------------------------------------------------------------------------------------------
# A8: fused per-head projection + tanh (VolTransformer / PlaneTransformer + torch.tanh of the MHCT blocks)
class _ProjectFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pcd, keys_res, res_scale, shift, rot, scales, heads, dim, key_stats):
        _require_cuda(pcd, keys_res, shift, rot, scales)
        pcd_c, res_c = _f32c(pcd), _f32c(keys_res)
        shift_c, rot_c, scales_c = _f32c(shift), _f32c(rot), _f32c(scales)
        B, _, N = pcd_c.shape
        rs = 1.0 if res_scale is None else float(res_scale)
        keys = torch.empty((B, heads * dim, N), dtype=torch.float32, device=pcd_c.device)
        sh = _lib.make_shape(B, heads, 1, N, dim, (2,) * dim)
        with torch.cuda.device(pcd_c.device):
            _call("ctb_project_fwd_stats", _ptr(pcd_c), _ptr(res_c), ctypes.c_float(rs), _ptr(shift_c), _ptr(rot_c),
                  _ptr(scales_c), _ptr(keys), _ptr(key_stats), ctypes.byref(sh), _stream(pcd_c))
        ctx.save_for_backward(pcd_c, res_c, shift_c, rot_c, scales_c, keys)
        ctx.rs, ctx.heads, ctx.dim = rs, heads, dim
        ctx.res_scale_is_tensor = isinstance(res_scale, torch.Tensor)
        return keys

    @staticmethod
    @once_differentiable
    def backward(ctx, g_keys):
        pcd_c, res_c, shift_c, rot_c, scales_c, keys = ctx.saved_tensors
        B, _, N = pcd_c.shape
        heads, dim = ctx.heads, ctx.dim
        g = _f32c(g_keys)
        g_pcd = torch.empty_like(pcd_c)
        g_res = torch.empty_like(res_c) if res_c is not None else None
        acc = torch.zeros((heads, 16), dtype=torch.float32, device=pcd_c.device)
        sh = _lib.make_shape(B, heads, 1, N, dim, (2,) * dim)
        ws_bytes = _lib.load().ctb_project_bwd_workspace_bytes(ctypes.byref(sh))
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=pcd_c.device)
        with torch.cuda.device(pcd_c.device):
            _call("ctb_project_bwd", _ptr(pcd_c), _ptr(res_c), ctypes.c_float(ctx.rs), _ptr(shift_c), _ptr(rot_c),
                  _ptr(scales_c), _ptr(keys), _ptr(g), _ptr(g_pcd), _ptr(g_res), _ptr(acc), _ptr(ws),
                  ctypes.c_size_t(ws_bytes), ctypes.byref(sh), _stream(pcd_c))
        g_shift = acc[:, 0:3].contiguous()
        g_rot = acc[:, 3:12].reshape(heads, 3, 3)
        g_scales = acc[:, 12:12 + dim].contiguous() if scales_c is not None else None
        g_rs = acc[:, 15].sum() if ctx.res_scale_is_tensor else None
        return g_pcd, g_res, g_rs, g_shift, g_rot, g_scales, None, None, None


def project_keys(orig_pcd, keys_res, shift, rot, scales=None, res_scale=None, *, heads, dim, key_stats=None):
    """tanh(((orig_pcd + res_scale * keys_res + shift) . rot)[:dim] * scales) -> keys [B, heads*dim, N].

    orig_pcd [B,3,N]; keys_res [B, heads*3, N] or [B,heads,3,N] or None; shift [heads,3]; rot [heads,3,3];
    scales [heads,dim] or None; res_scale None / python float / 0-dim tensor (MultiHeadAdaIn's `scale`).
    key_stats: optional zeroed float64 [2] device tensor that receives sum / sum of squares of the PRE-tanh keys."""
    return _ProjectFn.apply(orig_pcd, keys_res, res_scale, shift, rot, scales, heads, dim, key_stats)


def key_mean_var(key_stats, count):
    """mean and (unbiased, like torch.var) variance of the pre-tanh keys from the sums ctb_project_fwd_stats filled."""
    mean = key_stats[0] / count
    var = (key_stats[1] - key_stats[0] * mean) / max(count - 1, 1)
    return mean.float(), var.float()
