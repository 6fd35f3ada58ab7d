"""ctypes binding of libctb200.so (C ABI in include/ctb200.h).

The product path has NO CPU or PyTorch fallback: if the CUDA library is missing or a call fails, this
module raises.  Build the library with `python __graft_entry__.py build` (or `build()`), which runs
nvcc for sm_100a and leaves `cloud_transformers_b200/libctb200.so` in-tree.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libctb200.so")

CTB_OK = 0
CTB_ERR_INVALID_ARGUMENT = -1
CTB_ERR_UNSUPPORTED = -2
CTB_ERR_CUDA = -3
CTB_ERR_WORKSPACE = -4

REDUCE_MAX, REDUCE_SUM = 0, 1
MODE_ATOMIC, MODE_DETERMINISTIC, MODE_TILE = 0, 1, 2
OP_SPLAT_FWD, OP_SPLAT_BWD, OP_SLICE_FWD, OP_SLICE_BWD = 0, 1, 2, 3
DTYPE_F32, DTYPE_BF16 = 0, 1


class CtbShape(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int32), ("H", ctypes.c_int32), ("F", ctypes.c_int32), ("N", ctypes.c_int32),
                ("dim", ctypes.c_int32), ("size", ctypes.c_int32 * 3), ("grid_dtype", ctypes.c_int32)]


class CtbBnExchange(ctypes.Structure):
    _fields_ = [("peer_data", ctypes.c_void_p), ("peer_flag", ctypes.c_void_p), ("epoch", ctypes.c_void_p),
                ("done", ctypes.c_void_p), ("scratch", ctypes.c_void_p), ("rank", ctypes.c_int32),
                ("world", ctypes.c_int32)]


class CtbError(RuntimeError):
    def __init__(self, fn, status, detail=""):
        self.status = status
        super().__init__("libctb200: %s failed with status %d (%s)%s" % (fn, status, detail, ""))


_P = ctypes.c_void_p
_SH = ctypes.POINTER(CtbShape)
_I = ctypes.c_int

# name -> (restype, argtypes); every symbol include/ctb200.h declares
SIGNATURES = {
    "ctb_version": (_I, []),
    "ctb_strerror": (ctypes.c_char_p, [_I]),
    "ctb_last_cuda_error": (_I, []),
    "ctb_positions_fwd": (_I, [_P, _P, _P, _SH, _P]),
    "ctb_positions_bwd": (_I, [_P, _P, _P, _SH, _P]),
    "ctb_splat_fwd": (_I, [_P, _P, _P, _P, _P, _P, _SH, _I, _P]),
    "ctb_splat_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _SH, _I, _P]),
    "ctb_slice_fwd": (_I, [_P, _P, _P, _P, _P, _SH, _P]),
    "ctb_slice_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _SH, _P]),
    "ctb_mode_supported": (_I, [_SH, _I, _I, _I]),
    "ctb_plan_used": (_I, [_SH, _I]),
    "ctb_op_uses_plan": (_I, [_SH, _I, _I, _I]),
    "ctb_plan_bytes": (ctypes.c_size_t, [_SH]),
    "ctb_plan_build": (_I, [_P, _P, ctypes.c_size_t, _SH, _P]),
    "ctb_splat_fwd_keys": (_I, [_P, _P, _P, _P, _P, _SH, _I, _I, _P, _P]),
    "ctb_splat_bwd_keys": (_I, [_P, _P, _P, _P, _P, _P, _P, _SH, _I, _I, _P]),
    "ctb_slice_fwd_keys": (_I, [_P, _P, _P, _P, _SH, _I, _P, _P]),
    "ctb_slice_bwd_keys": (_I, [_P, _P, _P, _P, _P, _P, _SH, _I, _P, _P]),
    "ctb_project_fwd": (_I, [_P, _P, ctypes.c_float, _P, _P, _P, _P, _SH, _P]),
    "ctb_project_fwd_stats": (_I, [_P, _P, ctypes.c_float, _P, _P, _P, _P, _P, _SH, _P]),
    "ctb_project_bwd": (_I, [_P, _P, ctypes.c_float, _P, _P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_size_t, _SH, _P]),
    "ctb_project_bwd_workspace_bytes": (ctypes.c_size_t, [_SH]),
    "ctb_chamfer_workspace_bytes": (ctypes.c_size_t, [_I, _I, _I]),
    "ctb_chamfer_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, ctypes.c_size_t, _I, _I, _I, _P]),
    "ctb_chamfer_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "ctb_emd_max_points": (_I, []),
    "ctb_emd_workspace_bytes": (ctypes.c_size_t, [_I, _I]),
    "ctb_emd_fwd": (_I, [_P, _P, _P, _P, _P, ctypes.c_size_t, _I, _I, ctypes.c_float, _I, _P]),
    "ctb_emd_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _P]),
    "ctb_syncbn_scratch_bytes": (ctypes.c_uint64, [_I]),
    "ctb_syncbn_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, ctypes.POINTER(CtbBnExchange), _I, _I, _I, ctypes.c_float,
                            ctypes.c_float, _P]),
    "ctb_syncbn_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, ctypes.POINTER(CtbBnExchange), _I, _I, _I, _P]),
    "ctb_count_occupied": (_I, [_P, ctypes.c_uint64, _P, _P]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load libctb200.so (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                "cloud_transformers_b200: %s not found. Build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a). There is no CPU / PyTorch fallback for Splat / Slice." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def make_shape(B, H, F, N, dim, size, grid_dtype=0):
    s = CtbShape()
    s.B, s.H, s.F, s.N, s.dim = int(B), int(H), int(F), int(N), int(dim)
    s.grid_dtype = int(grid_dtype)
    for a in range(3):
        s.size[a] = int(size[a]) if a < len(size) else 1
    return s


def check(fn_name, status):
    if status != CTB_OK:
        lib = load()
        detail = lib.ctb_strerror(status).decode()
        if status == CTB_ERR_CUDA:
            detail += ", cudaError=%d" % lib.ctb_last_cuda_error()
        raise CtbError(fn_name, status, detail)
