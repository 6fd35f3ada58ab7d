// ctb200.cu -- C ABI of libctb200.so (declared in include/ctb200.h).  sm_100a only.
#include <cuda_runtime.h>
#include <initializer_list>
#include <stdint.h>
#include "../../include/ctb200.h"
#include "ctb_positions.cuh"
#include "ctb_generic.cuh"
#include "ctb_plan.cuh"
#include "ctb_tile.cuh"
#include "ctb_tile_cl.cuh"
#include "ctb_binned.cuh"
#include "ctb_sgather.cuh"
#include "ctb_project.cuh"
#include "ctb_chamfer.cuh"
#include "ctb_emd.cuh"
#include "ctb_syncbn.cuh"

namespace {

thread_local int g_last_cuda_error = 0;

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = (int)e;
  return CTB_ERR_CUDA;
}

inline int cuda_status(cudaError_t e) {
  if (e == cudaSuccess) return CTB_OK;
  if (e == cudaErrorNotSupported) return CTB_ERR_UNSUPPORTED;
  return cuda_fail(e);
}

#define CTB_CUDA(expr)                                \
  do {                                                \
    cudaError_t _e = (expr);                          \
    if (_e != cudaSuccess) return cuda_fail(_e);      \
  } while (0)

#define CTB_LAUNCH_CHECK()                            \
  do {                                                \
    cudaError_t _e = cudaGetLastError();              \
    if (_e != cudaSuccess) return cuda_fail(_e);      \
  } while (0)

int check_shape(const ctb_shape* s, bool need_f) {
  if (!s) return CTB_ERR_INVALID_ARGUMENT;
  if (s->B <= 0 || s->H <= 0 || s->N <= 0) return CTB_ERR_INVALID_ARGUMENT;
  if (need_f && s->F <= 0) return CTB_ERR_INVALID_ARGUMENT;
  if (s->dim != 2 && s->dim != 3) return CTB_ERR_INVALID_ARGUMENT;
  if (s->grid_dtype != CTB_DTYPE_F32 && s->grid_dtype != CTB_DTYPE_BF16) return CTB_ERR_INVALID_ARGUMENT;
  long long C = 1;
  for (int a = 0; a < s->dim; ++a) {
    if (s->size[a] < 2) return CTB_ERR_INVALID_ARGUMENT;
    C *= s->size[a];
    if (C >= (1ll << 30)) return CTB_ERR_UNSUPPORTED;
  }
  const long long S = 1ll << s->dim;
  if (S * (long long)s->N >= (1ll << 31) - 1) return CTB_ERR_UNSUPPORTED;  // arg is int32
  const long long units = (long long)s->B * s->H;
  const long long chunks = (s->N + ctb::kGenericBlock - 1) / ctb::kGenericBlock;
  if (units * chunks >= (1ll << 31) - 1) return CTB_ERR_UNSUPPORTED;
  return CTB_OK;
}

inline long long cells(const ctb_shape* s) {
  long long C = 1;
  for (int a = 0; a < s->dim; ++a) C *= s->size[a];
  return C;
}

struct Launch {
  int chunks;
  unsigned blocks;
  cudaStream_t stream;
};

inline Launch make_launch(const ctb_shape* s, void* stream) {
  Launch l;
  l.chunks = (s->N + ctb::kGenericBlock - 1) / ctb::kGenericBlock;
  l.blocks = (unsigned)((long long)s->B * s->H * l.chunks);
  l.stream = (cudaStream_t)stream;
  return l;
}

// ---- typed implementations ----------------------------------------------------------------------
template <int D>
int positions_fwd_impl(const float* keys, float* lc, int64_t* idx, const ctb_shape* s, void* stream) {
  const ctb::Grid<D> g = ctb::make_grid<D>(s->size);
  const Launch l = make_launch(s, stream);
  ctb::positions_fwd_kernel<D><<<l.blocks, ctb::kGenericBlock, 0, l.stream>>>(keys, lc, (long long*)idx, g,
                                                                              s->N, l.chunks);
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

template <int D>
int positions_bwd_impl(const float* keys, const float* grad_lc, float* grad_keys, const ctb_shape* s,
                       void* stream) {
  const ctb::Grid<D> g = ctb::make_grid<D>(s->size);
  const Launch l = make_launch(s, stream);
  ctb::positions_bwd_kernel<D><<<l.blocks, ctb::kGenericBlock, 0, l.stream>>>(keys, grad_lc, grad_keys, g,
                                                                              s->N, l.chunks);
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

template <int D, bool KEYS>
int splat_fwd_atomic_impl(ctb::PointSource src, const float* feat, const float* pad, float* z, int32_t* arg,
                          const ctb_shape* s, int reduce, void* stream) {
  const ctb::Grid<D> g = ctb::make_grid<D>(s->size);
  const Launch l = make_launch(s, stream);
  const size_t n_grid = (size_t)s->B * s->H * s->F * g.C;
  CTB_CUDA(cudaMemsetAsync(z, 0, n_grid * sizeof(float), l.stream));
  if (reduce == CTB_REDUCE_SUM) {
    ctb::splat_fwd_atomic_kernel<D, KEYS, 0, true><<<l.blocks, ctb::kGenericBlock, 0, l.stream>>>(
        src, feat, pad, z, nullptr, g, s->H, s->F, s->N, l.chunks);
    CTB_LAUNCH_CHECK();
    return CTB_OK;
  }
  CTB_CUDA(cudaMemsetAsync(arg, 0xFF, n_grid * sizeof(int32_t), l.stream));
  ctb::splat_fwd_atomic_kernel<D, KEYS, 0, false><<<l.blocks, ctb::kGenericBlock, 0, l.stream>>>(
      src, feat, pad, z, arg, g, s->H, s->F, s->N, l.chunks);
  CTB_LAUNCH_CHECK();
  ctb::splat_fwd_atomic_kernel<D, KEYS, 1, false><<<l.blocks, ctb::kGenericBlock, 0, l.stream>>>(
      src, feat, pad, z, arg, g, s->H, s->F, s->N, l.chunks);
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

template <int D, bool KEYS>
int splat_bwd_impl(ctb::PointSource src, const float* feat, const float* pad, const float* grad_z,
                   const int32_t* arg, float* grad_feat, float* grad_pos, const ctb_shape* s, int reduce,
                   void* stream) {
  const ctb::Grid<D> g = ctb::make_grid<D>(s->size);
  const Launch l = make_launch(s, stream);
  if (reduce == CTB_REDUCE_SUM)
    ctb::splat_bwd_kernel<D, KEYS, true><<<l.blocks, ctb::kGenericBlock, 0, l.stream>>>(
        src, feat, pad, grad_z, arg, grad_feat, grad_pos, g, s->H, s->F, s->N, l.chunks);
  else
    ctb::splat_bwd_kernel<D, KEYS, false><<<l.blocks, ctb::kGenericBlock, 0, l.stream>>>(
        src, feat, pad, grad_z, arg, grad_feat, grad_pos, g, s->H, s->F, s->N, l.chunks);
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

template <int D, bool KEYS>
int slice_fwd_impl(ctb::PointSource src, const float* grid, const float* pad, float* out, const ctb_shape* s,
                   void* stream) {
  const ctb::Grid<D> g = ctb::make_grid<D>(s->size);
  const Launch l = make_launch(s, stream);
  ctb::slice_fwd_kernel<D, KEYS><<<l.blocks, ctb::kGenericBlock, 0, l.stream>>>(src, grid, pad, out, g, s->H,
                                                                                s->F, s->N, l.chunks);
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

template <int D, bool KEYS>
int slice_bwd_atomic_impl(ctb::PointSource src, const float* grid, const float* pad, const float* grad_out,
                          float* grad_grid, float* grad_pos, const ctb_shape* s, void* stream) {
  const ctb::Grid<D> g = ctb::make_grid<D>(s->size);
  const Launch l = make_launch(s, stream);
  CTB_CUDA(cudaMemsetAsync(grad_grid, 0, (size_t)s->B * s->H * s->F * g.C * sizeof(float), l.stream));
  ctb::slice_bwd_kernel<D, KEYS, true><<<l.blocks, ctb::kGenericBlock, 0, l.stream>>>(
      src, grid, pad, grad_out, grad_grid, grad_pos, g, s->H, s->F, s->N, l.chunks);
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

// TILE-mode scatter: channel-lane kernel for coarse dense grids, point-lane kernel otherwise
template <int D, typename GT>
cudaError_t tile_scatter_any(const float* keys, const float* feat, const float* pad, GT* z, int* arg,
                             const ctb_shape* s, bool sum, cudaStream_t stream) {
  static const bool no_cl = getenv("CTB_NO_CHANNEL_LANE") != nullptr || getenv("CTB_QUAD_SUM") != nullptr;
  cudaError_t e = cudaSuccess;
  // measured (profiles/r01_*): lanes = channels wins for the contended float-sum atomics (c3d 0.52 -> 0.32 ms) but
  // not for max, whose second pass is broadcast-friendly shared loads
  static const bool cl_max = getenv("CTB_CHANNEL_LANE_MAX") != nullptr;
  if (!no_cl && (sum || cl_max) && ctb::cl_scatter_try<D, GT>(keys, feat, pad, z, arg, s, sum, stream, &e)) return e;
  return ctb::tile_scatter<D, GT>(keys, feat, pad, z, arg, s, sum, stream);
}

// dimension x grid-dtype dispatch for the tile kernels
typedef __nv_bfloat16 bf16_t;
inline cudaError_t scatter_dispatch(const float* keys, const float* feat, const float* pad, void* z, int* arg,
                                    const ctb_shape* s, bool sum, cudaStream_t st) {
  const bool bf = s->grid_dtype == CTB_DTYPE_BF16;
  if (s->dim == 2)
    return bf ? tile_scatter_any<2, bf16_t>(keys, feat, pad, (bf16_t*)z, arg, s, sum, st)
              : tile_scatter_any<2, float>(keys, feat, pad, (float*)z, arg, s, sum, st);
  return bf ? tile_scatter_any<3, bf16_t>(keys, feat, pad, (bf16_t*)z, arg, s, sum, st)
            : tile_scatter_any<3, float>(keys, feat, pad, (float*)z, arg, s, sum, st);
}

template <int MODE>
cudaError_t gather_dispatch(const float* keys, const void* t1, const int* t2, const float* in, const float* pad,
                            float* out, float* grad_keys, const ctb_shape* s, cudaStream_t st, bool overlap_prev = false) {
  const bool bf = s->grid_dtype == CTB_DTYPE_BF16;
  if (s->dim == 2)
    return bf ? ctb::tile_gather<2, MODE, bf16_t>(keys, (const bf16_t*)t1, t2, in, pad, out, grad_keys, s, st, overlap_prev)
              : ctb::tile_gather<2, MODE, float>(keys, (const float*)t1, t2, in, pad, out, grad_keys, s, st, overlap_prev);
  return bf ? ctb::tile_gather<3, MODE, bf16_t>(keys, (const bf16_t*)t1, t2, in, pad, out, grad_keys, s, st, overlap_prev)
            : ctb::tile_gather<3, MODE, float>(keys, (const float*)t1, t2, in, pad, out, grad_keys, s, st, overlap_prev);
}

// binned (output-stationary) scatter over the plan's entry lists: reduce x grid-dtype dispatch
inline cudaError_t binned_dispatch(const float* feat, const float* pad, const void* plan, void* z, int* arg,
                                   const ctb_shape* s, bool sum, cudaStream_t st) {
  const bool bf = s->grid_dtype == CTB_DTYPE_BF16;
  if (sum)
    return bf ? ctb::bin_scatter<true, bf16_t>(feat, pad, plan, (bf16_t*)z, arg, s, st)
              : ctb::bin_scatter<true, float>(feat, pad, plan, (float*)z, arg, s, st);
  return bf ? ctb::bin_scatter<false, bf16_t>(feat, pad, plan, (bf16_t*)z, arg, s, st)
            : ctb::bin_scatter<false, float>(feat, pad, plan, (float*)z, arg, s, st);
}

inline bool binned_ok(const ctb_shape* s) {
  ctb::BinConfig bc;
  return ctb::bin_config(s, false, &bc);      // the max variant needs the larger footprint
}
// TILE mode takes the binned scatters only where they win (measured, profiles/r02_*): the sums on every dense grid
// (at least one point per four cells), the max on coarse grids (<= 1024 cells: the whole unit is one tile and every
// window gets a long run of entries).  DETERMINISTIC always takes them.
inline bool binned_in_tile(const ctb_shape* s, bool sum) {
  static const bool off = getenv("CTB_NO_BINNED") != nullptr;
  static const bool all = getenv("CTB_BINNED_ALL") != nullptr;
  if (off || !binned_ok(s)) return false;
  if (all) return true;
  return ctb::plan_dense(s) && (sum || ctb::shape_cells(s) <= 1024);
}

#define CTB_DISPATCH_DIM(shape, call2, call3) ((shape)->dim == 2 ? (call2) : (call3))

}  // namespace

// ---- exported entries -----------------------------------------------------------------------------
extern "C" {

int ctb_version(void) { return CTB_VERSION; }

#ifdef CTB_PHASE_TIMERS
// developer tool (not in the header): copy out / reset the phase cycle counters of ctb_tile.cuh
int ctb_debug_phase(unsigned long long* host16, int reset) {
  cudaDeviceSynchronize();
  if (host16 && cudaMemcpyFromSymbol(host16, ctb::g_phase, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[16] = {0};
    if (cudaMemcpyToSymbol(ctb::g_phase, z, sizeof(z)) != cudaSuccess) return -1;
  }
  return 0;
}
#endif

const char* ctb_strerror(int status) {
  switch (status) {
    case CTB_OK: return "ok";
    case CTB_ERR_INVALID_ARGUMENT: return "invalid argument";
    case CTB_ERR_UNSUPPORTED: return "unsupported shape";
    case CTB_ERR_CUDA: return "CUDA runtime error (see ctb_last_cuda_error)";
    case CTB_ERR_WORKSPACE: return "plan / workspace missing or too small";
    default: return "unknown status";
  }
}

int ctb_last_cuda_error(void) { return g_last_cuda_error; }

int ctb_positions_fwd(const float* keys, float* lc, int64_t* idx, const ctb_shape* shape, void* stream) {
  int st = check_shape(shape, false);
  if (st) return st;
  if (!keys || !lc || !idx) return CTB_ERR_INVALID_ARGUMENT;
  return CTB_DISPATCH_DIM(shape, positions_fwd_impl<2>(keys, lc, idx, shape, stream),
                          positions_fwd_impl<3>(keys, lc, idx, shape, stream));
}

int ctb_positions_bwd(const float* keys, const float* grad_lc, float* grad_keys, const ctb_shape* shape,
                      void* stream) {
  int st = check_shape(shape, false);
  if (st) return st;
  if (!keys || !grad_lc || !grad_keys) return CTB_ERR_INVALID_ARGUMENT;
  return CTB_DISPATCH_DIM(shape, positions_bwd_impl<2>(keys, grad_lc, grad_keys, shape, stream),
                          positions_bwd_impl<3>(keys, grad_lc, grad_keys, shape, stream));
}

int ctb_splat_fwd(const float* lc, const int64_t* idx, const float* features, const float* pad, float* z,
                  int32_t* arg, const ctb_shape* shape, int reduce, void* stream) {
  int st = check_shape(shape, true);
  if (st) return st;
  if (!lc || !idx || !features || !z) return CTB_ERR_INVALID_ARGUMENT;
  if (reduce != CTB_REDUCE_MAX && reduce != CTB_REDUCE_SUM) return CTB_ERR_INVALID_ARGUMENT;
  if (reduce == CTB_REDUCE_MAX && !arg) return CTB_ERR_INVALID_ARGUMENT;
  ctb::PointSource src{nullptr, lc, idx};
  return CTB_DISPATCH_DIM(shape,
                          (splat_fwd_atomic_impl<2, false>(src, features, pad, z, arg, shape, reduce, stream)),
                          (splat_fwd_atomic_impl<3, false>(src, features, pad, z, arg, shape, reduce, stream)));
}

int ctb_splat_bwd(const float* lc, const int64_t* idx, const float* features, const float* pad,
                  const float* grad_z, const int32_t* arg, float* grad_features, float* grad_lc,
                  const ctb_shape* shape, int reduce, void* stream) {
  int st = check_shape(shape, true);
  if (st) return st;
  if (!lc || !idx || !features || !grad_z || !grad_features || !grad_lc) return CTB_ERR_INVALID_ARGUMENT;
  if (reduce != CTB_REDUCE_MAX && reduce != CTB_REDUCE_SUM) return CTB_ERR_INVALID_ARGUMENT;
  if (reduce == CTB_REDUCE_MAX && !arg) return CTB_ERR_INVALID_ARGUMENT;
  ctb::PointSource src{nullptr, lc, idx};
  return CTB_DISPATCH_DIM(
      shape,
      (splat_bwd_impl<2, false>(src, features, pad, grad_z, arg, grad_features, grad_lc, shape, reduce, stream)),
      (splat_bwd_impl<3, false>(src, features, pad, grad_z, arg, grad_features, grad_lc, shape, reduce, stream)));
}

int ctb_slice_fwd(const float* lc, const int64_t* idx, const float* grid, const float* pad, float* out,
                  const ctb_shape* shape, void* stream) {
  int st = check_shape(shape, true);
  if (st) return st;
  if (!lc || !idx || !grid || !out) return CTB_ERR_INVALID_ARGUMENT;
  ctb::PointSource src{nullptr, lc, idx};
  return CTB_DISPATCH_DIM(shape, (slice_fwd_impl<2, false>(src, grid, pad, out, shape, stream)),
                          (slice_fwd_impl<3, false>(src, grid, pad, out, shape, stream)));
}

int ctb_slice_bwd(const float* lc, const int64_t* idx, const float* grid, const float* pad,
                  const float* grad_out, float* grad_grid, float* grad_lc, const ctb_shape* shape,
                  void* stream) {
  int st = check_shape(shape, true);
  if (st) return st;
  if (!lc || !idx || !grid || !grad_out || !grad_grid || !grad_lc) return CTB_ERR_INVALID_ARGUMENT;
  ctb::PointSource src{nullptr, lc, idx};
  return CTB_DISPATCH_DIM(
      shape, (slice_bwd_atomic_impl<2, false>(src, grid, pad, grad_out, grad_grid, grad_lc, shape, stream)),
      (slice_bwd_atomic_impl<3, false>(src, grid, pad, grad_out, grad_grid, grad_lc, shape, stream)));
}

int ctb_mode_supported(const ctb_shape* shape, int op, int reduce, int mode) {
  if (check_shape(shape, true)) return 0;
  if (mode == CTB_MODE_ATOMIC) return 1;
  if (mode != CTB_MODE_TILE && mode != CTB_MODE_DETERMINISTIC) return 0;
  const bool sum = reduce == CTB_REDUCE_SUM;
  const bool det = mode == CTB_MODE_DETERMINISTIC;
  ctb::TileConfig tc;
  ctb::TileConfig gc;
  // DETERMINISTIC: the two scatters are the binned kernels (fixed summation order); TILE uses them too when the
  // caller hands in a plan, and the shared-memory tile scatters otherwise
  switch (op) {
    case CTB_OP_SPLAT_FWD:
      return det ? binned_ok(shape) : (binned_in_tile(shape, sum) || ctb::tile_scatter_config(shape, sum, !sum, &tc));
    case CTB_OP_SPLAT_BWD:
      return sum ? (ctb::gather_config(shape, ctb::GATHER_SLICE_FWD, &gc) &&
                    ctb::gather_config(shape, ctb::GATHER_SLICE_BWD_KEYS, &gc))
                 : ctb::gather_config(shape, ctb::GATHER_SPLAT_BWD, &gc);
    case CTB_OP_SLICE_FWD: return ctb::gather_config(shape, ctb::GATHER_SLICE_FWD, &gc);
    case CTB_OP_SLICE_BWD:
      return (det ? binned_ok(shape) : (binned_in_tile(shape, true) || ctb::tile_scatter_config(shape, true, false, &tc))) &&
             ctb::gather_config(shape, ctb::GATHER_SLICE_BWD_KEYS, &gc);
    default: return 0;
  }
}

int ctb_plan_used(const ctb_shape* shape, int mode) {
  if (check_shape(shape, true)) return 0;
  if (mode != CTB_MODE_TILE && mode != CTB_MODE_DETERMINISTIC) return 0;
  return (mode == CTB_MODE_DETERMINISTIC ? binned_ok(shape) : binned_in_tile(shape, true)) ? 1 : 0;
}

// Slice forward in sorted point order (ctb_sgather.cuh) where the shape's plan exists anyway
inline bool sorted_slice_ok(const ctb_shape* s, int mode) {
  static const bool off = getenv("CTB_NO_SORTED_SLICE") != nullptr;
  ctb::SortGatherConfig sc;
  return !off && (mode == CTB_MODE_TILE || mode == CTB_MODE_DETERMINISTIC) && ctb_plan_used(s, mode) &&
         ctb::sgather_config(s, &sc);
}

int ctb_op_uses_plan(const ctb_shape* shape, int op, int reduce, int mode) {
  if (check_shape(shape, true)) return 0;
  if (op == CTB_OP_SLICE_FWD) return sorted_slice_ok(shape, mode) ? 1 : 0;
  if (op != CTB_OP_SPLAT_FWD && op != CTB_OP_SLICE_BWD) return 0;          // the backward gathers never read it
  const bool sum = op == CTB_OP_SLICE_BWD || reduce == CTB_REDUCE_SUM;
  if (mode == CTB_MODE_DETERMINISTIC) return binned_ok(shape) ? 1 : 0;
  if (mode == CTB_MODE_TILE) return binned_in_tile(shape, sum) ? 1 : 0;
  return 0;
}

size_t ctb_plan_bytes(const ctb_shape* shape) {
  if (check_shape(shape, false)) return 0;
  if (!ctb::plan_supported(shape)) return 0;
  return ctb::plan_layout(shape).bytes;
}

int ctb_plan_build(const float* keys, void* plan, size_t plan_bytes, const ctb_shape* shape, void* stream) {
  int st = check_shape(shape, false);
  if (st) return st;
  if (!keys || !plan) return CTB_ERR_INVALID_ARGUMENT;
  if (!ctb::plan_supported(shape)) return CTB_ERR_UNSUPPORTED;
  if (plan_bytes < ctb::plan_layout(shape).bytes) return CTB_ERR_WORKSPACE;
  return cuda_status(shape->dim == 2 ? ctb::plan_build<2>(keys, plan, shape, (cudaStream_t)stream)
                                     : ctb::plan_build<3>(keys, plan, shape, (cudaStream_t)stream));
}

int ctb_splat_fwd_keys(const float* keys, const float* features, const float* pad, void* z_any, int32_t* arg,
                       const ctb_shape* shape, int reduce, int mode, const void* plan, void* stream) {
  float* z = (float*)z_any;
  int st = check_shape(shape, true);
  if (st) return st;
  if (!keys || !features || !z) return CTB_ERR_INVALID_ARGUMENT;
  if (reduce != CTB_REDUCE_MAX && reduce != CTB_REDUCE_SUM) return CTB_ERR_INVALID_ARGUMENT;
  if (reduce == CTB_REDUCE_MAX && !arg) return CTB_ERR_INVALID_ARGUMENT;
  if (shape->grid_dtype != CTB_DTYPE_F32 && mode == CTB_MODE_ATOMIC) return CTB_ERR_UNSUPPORTED;
  if (mode == CTB_MODE_DETERMINISTIC || (mode == CTB_MODE_TILE && plan && binned_in_tile(shape, reduce == CTB_REDUCE_SUM))) {
    if (!plan) return CTB_ERR_WORKSPACE;
    if (!binned_ok(shape)) return CTB_ERR_UNSUPPORTED;
    const bool sum = reduce == CTB_REDUCE_SUM;
    return cuda_status(binned_dispatch(features, pad, plan, z_any, sum ? nullptr : arg, shape, sum,
                                       (cudaStream_t)stream));
  }
  if (mode == CTB_MODE_TILE) {
    const bool sum = reduce == CTB_REDUCE_SUM;
    return cuda_status(scatter_dispatch(keys, features, pad, z_any, sum ? nullptr : arg, shape, sum, (cudaStream_t)stream));
  }
  if (mode != CTB_MODE_ATOMIC) return CTB_ERR_INVALID_ARGUMENT;
  ctb::PointSource src{keys, nullptr, nullptr};
  return CTB_DISPATCH_DIM(shape,
                          (splat_fwd_atomic_impl<2, true>(src, features, pad, z, arg, shape, reduce, stream)),
                          (splat_fwd_atomic_impl<3, true>(src, features, pad, z, arg, shape, reduce, stream)));
}

int ctb_splat_bwd_keys(const float* keys, const float* features, const float* pad, const void* grad_z_any,
                       const int32_t* arg, float* grad_features, float* grad_keys, const ctb_shape* shape,
                       int reduce, int mode, void* stream) {
  const float* grad_z = (const float*)grad_z_any;
  int st = check_shape(shape, true);
  if (st) return st;
  if (!keys || !features || !grad_z || !grad_features || !grad_keys) return CTB_ERR_INVALID_ARGUMENT;
  if (reduce != CTB_REDUCE_MAX && reduce != CTB_REDUCE_SUM) return CTB_ERR_INVALID_ARGUMENT;
  if (reduce == CTB_REDUCE_MAX && !arg) return CTB_ERR_INVALID_ARGUMENT;
  if (mode == CTB_MODE_DETERMINISTIC || mode == CTB_MODE_TILE) {
    if (reduce == CTB_REDUCE_SUM) {
      // every entry contributes: grad_features is the Slice forward of grad_z, and the weight gradients are those of
      // a Slice backward whose "grid" is grad_z and whose upstream gradient is the features -- both gathers exist
      st = cuda_status(gather_dispatch<ctb::GATHER_SLICE_FWD>(keys, grad_z_any, nullptr, nullptr, pad, grad_features,
                                                              nullptr, shape, (cudaStream_t)stream));
      if (st) return st;
      return cuda_status(gather_dispatch<ctb::GATHER_SLICE_BWD_KEYS>(keys, grad_z_any, nullptr, features, pad, nullptr,
                                                                      grad_keys, shape, (cudaStream_t)stream));
    }
    return cuda_status(gather_dispatch<ctb::GATHER_SPLAT_BWD>(keys, grad_z_any, arg, features, pad, grad_features,
                                                               grad_keys, shape, (cudaStream_t)stream));
  }
  if (shape->grid_dtype != CTB_DTYPE_F32) return CTB_ERR_UNSUPPORTED;
  if (mode != CTB_MODE_ATOMIC) return CTB_ERR_INVALID_ARGUMENT;
  ctb::PointSource src{keys, nullptr, nullptr};
  return CTB_DISPATCH_DIM(
      shape,
      (splat_bwd_impl<2, true>(src, features, pad, grad_z, arg, grad_features, grad_keys, shape, reduce, stream)),
      (splat_bwd_impl<3, true>(src, features, pad, grad_z, arg, grad_features, grad_keys, shape, reduce, stream)));
}

int ctb_slice_fwd_keys(const float* keys, const void* grid_any, const float* pad, float* out,
                       const ctb_shape* shape, int mode, const void* plan, void* stream) {
  const float* grid = (const float*)grid_any;
  int st = check_shape(shape, true);
  if (st) return st;
  if (!keys || !grid || !out) return CTB_ERR_INVALID_ARGUMENT;
  if (plan && sorted_slice_ok(shape, mode) && (((uintptr_t)grid | (uintptr_t)out) & 15u) == 0)
    return cuda_status(shape->dim == 2 ? ctb::sorted_slice_fwd<2>(keys, grid, pad, out, plan, shape, (cudaStream_t)stream)
                                       : ctb::sorted_slice_fwd<3>(keys, grid, pad, out, plan, shape, (cudaStream_t)stream));
  if (mode == CTB_MODE_DETERMINISTIC || mode == CTB_MODE_TILE)
    return cuda_status(gather_dispatch<ctb::GATHER_SLICE_FWD>(keys, grid_any, nullptr, nullptr, pad, out, nullptr, shape,
                                                               (cudaStream_t)stream));
  if (shape->grid_dtype != CTB_DTYPE_F32) return CTB_ERR_UNSUPPORTED;
  if (mode != CTB_MODE_ATOMIC) return CTB_ERR_INVALID_ARGUMENT;
  ctb::PointSource src{keys, nullptr, nullptr};
  return CTB_DISPATCH_DIM(shape, (slice_fwd_impl<2, true>(src, grid, pad, out, shape, stream)),
                          (slice_fwd_impl<3, true>(src, grid, pad, out, shape, stream)));
}

int ctb_slice_bwd_keys(const float* keys, const void* grid_any, const float* pad, const float* grad_out,
                       void* grad_grid_any, float* grad_keys, const ctb_shape* shape, int mode, const void* plan,
                       void* stream) {
  const float* grid = (const float*)grid_any;
  float* grad_grid = (float*)grad_grid_any;
  int st = check_shape(shape, true);
  if (st) return st;
  if (!keys || !grid || !grad_out || !grad_grid || !grad_keys) return CTB_ERR_INVALID_ARGUMENT;
  if (mode == CTB_MODE_DETERMINISTIC || mode == CTB_MODE_TILE) {
    if (!ctb_mode_supported(shape, CTB_OP_SLICE_BWD, CTB_REDUCE_SUM, mode)) return CTB_ERR_UNSUPPORTED;
    // grad_grid: scatter-add of grad_out * pad (the Splat-sum kernel of the mode) ...
    if (mode == CTB_MODE_DETERMINISTIC || (plan && binned_in_tile(shape, true))) {
      if (!plan) return CTB_ERR_WORKSPACE;
      st = cuda_status(binned_dispatch(grad_out, pad, plan, grad_grid_any, nullptr, shape, true, (cudaStream_t)stream));
    } else {
      st = cuda_status(scatter_dispatch(keys, grad_out, pad, grad_grid_any, nullptr, shape, true, (cudaStream_t)stream));
    }
    if (st) return st;
    // ... and grad_keys: tile gather against the convolved grid.  It reads keys / grid / grad_out and writes
    // grad_keys only -- nothing the scatter above touches -- so in TILE mode it may overlap the scatter's last wave.
    return cuda_status(gather_dispatch<ctb::GATHER_SLICE_BWD_KEYS>(keys, grid_any, nullptr, grad_out, pad, nullptr,
                                                                    grad_keys, shape, (cudaStream_t)stream,
                                                                    mode == CTB_MODE_TILE));
  }
  if (shape->grid_dtype != CTB_DTYPE_F32) return CTB_ERR_UNSUPPORTED;
  if (mode != CTB_MODE_ATOMIC) return CTB_ERR_INVALID_ARGUMENT;
  ctb::PointSource src{keys, nullptr, nullptr};
  return CTB_DISPATCH_DIM(
      shape, (slice_bwd_atomic_impl<2, true>(src, grid, pad, grad_out, grad_grid, grad_keys, shape, stream)),
      (slice_bwd_atomic_impl<3, true>(src, grid, pad, grad_out, grad_grid, grad_keys, shape, stream)));
}

int ctb_project_fwd_stats(const float* pcd, const float* keys_res, float res_scale, const float* shift, const float* rot,
                          const float* scales, float* keys, double* key_stats, const ctb_shape* shape, void* stream) {
  int st = check_shape(shape, false);
  if (st) return st;
  if (!pcd || !shift || !rot || !keys) return CTB_ERR_INVALID_ARGUMENT;
  const int chunks = (shape->N + ctb::kProjBlock - 1) / ctb::kProjBlock;
  const unsigned blocks = (unsigned)((long long)shape->B * shape->H * chunks);
  if (shape->dim == 2)
    ctb::project_fwd_kernel<2><<<blocks, ctb::kProjBlock, 0, (cudaStream_t)stream>>>(
        pcd, keys_res, res_scale, shift, rot, scales, keys, key_stats, shape->H, shape->N, chunks);
  else
    ctb::project_fwd_kernel<3><<<blocks, ctb::kProjBlock, 0, (cudaStream_t)stream>>>(
        pcd, keys_res, res_scale, shift, rot, scales, keys, key_stats, shape->H, shape->N, chunks);
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

int ctb_project_fwd(const float* pcd, const float* keys_res, float res_scale, const float* shift, const float* rot,
                    const float* scales, float* keys, const ctb_shape* shape, void* stream) {
  return ctb_project_fwd_stats(pcd, keys_res, res_scale, shift, rot, scales, keys, nullptr, shape, stream);
}

int ctb_project_bwd(const float* pcd, const float* keys_res, float res_scale, const float* shift, const float* rot,
                    const float* scales, const float* keys, const float* grad_keys, float* grad_pcd,
                    float* grad_keys_res, float* param_acc, float* workspace, size_t workspace_bytes,
                    const ctb_shape* shape, void* stream) {
  int st = check_shape(shape, false);
  if (st) return st;
  if (!pcd || !shift || !rot || !keys || !grad_keys || !grad_pcd || !param_acc) return CTB_ERR_INVALID_ARGUMENT;
  if (keys_res && !grad_keys_res) return CTB_ERR_INVALID_ARGUMENT;
  if (!workspace || workspace_bytes < ctb_project_bwd_workspace_bytes(shape)) return CTB_ERR_WORKSPACE;
  const int chunks = (shape->N + 31) / 32;
  const unsigned blocks = (unsigned)((long long)shape->B * chunks);
  const dim3 block(32, shape->H < 32 ? shape->H : 32);
  if (shape->dim == 2)
    ctb::project_bwd_kernel<2><<<blocks, block, 0, (cudaStream_t)stream>>>(
        pcd, keys_res, res_scale, shift, rot, scales, keys, grad_keys, grad_pcd, grad_keys_res, workspace, shape->H,
        shape->N, chunks);
  else
    ctb::project_bwd_kernel<3><<<blocks, block, 0, (cudaStream_t)stream>>>(
        pcd, keys_res, res_scale, shift, rot, scales, keys, grad_keys, grad_pcd, grad_keys_res, workspace, shape->H,
        shape->N, chunks);
  CTB_LAUNCH_CHECK();
  const int cells = shape->H * ctb::kProjAcc;
  ctb::project_param_reduce_kernel<<<cells, 128, 0, (cudaStream_t)stream>>>(workspace, param_acc, (int)blocks, shape->H);
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

size_t ctb_project_bwd_workspace_bytes(const ctb_shape* shape) {
  if (check_shape(shape, false)) return 0;
  const size_t chunks = (size_t)(shape->N + 31) / 32;
  return (size_t)shape->B * chunks * shape->H * ctb::kProjAcc * sizeof(float);
}

size_t ctb_chamfer_workspace_bytes(int B, int n, int m) {
  if (B <= 0 || n <= 0 || m <= 0) return 0;
  return ((size_t)B * n + (size_t)B * m) * sizeof(unsigned long long);
}

int ctb_chamfer_fwd(const float* xyz1, const float* xyz2, float* dist1, float* dist2, int32_t* idx1, int32_t* idx2,
                    void* workspace, size_t workspace_bytes, int B, int n, int m, void* stream) {
  if (B <= 0 || n <= 0 || m <= 0) return CTB_ERR_INVALID_ARGUMENT;
  if (!xyz1 || !xyz2 || !dist1 || !dist2 || !idx1 || !idx2) return CTB_ERR_INVALID_ARGUMENT;
  if (!workspace || workspace_bytes < ctb_chamfer_workspace_bytes(B, n, m)) return CTB_ERR_WORKSPACE;
  if (B > 65535) return CTB_ERR_UNSUPPORTED;
  return cuda_status(ctb::chamfer_forward(xyz1, xyz2, dist1, dist2, idx1, idx2, (unsigned long long*)workspace, B, n, m,
                                          (cudaStream_t)stream));
}

int ctb_chamfer_bwd(const float* xyz1, const float* xyz2, const float* grad_dist1, const float* grad_dist2,
                    const int32_t* idx1, const int32_t* idx2, float* grad_xyz1, float* grad_xyz2, int B, int n, int m,
                    void* stream) {
  if (B <= 0 || n <= 0 || m <= 0) return CTB_ERR_INVALID_ARGUMENT;
  if (!xyz1 || !xyz2 || !grad_dist1 || !grad_dist2 || !idx1 || !idx2 || !grad_xyz1 || !grad_xyz2)
    return CTB_ERR_INVALID_ARGUMENT;
  if (B > 65535) return CTB_ERR_UNSUPPORTED;
  return cuda_status(ctb::chamfer_backward(xyz1, xyz2, grad_dist1, grad_dist2, idx1, idx2, grad_xyz1, grad_xyz2, B, n, m,
                                           (cudaStream_t)stream));
}

int ctb_emd_max_points(void) { return ctb::kEmdMaxPoints; }

size_t ctb_emd_workspace_bytes(int B, int n) {
  return (B > 0 && n > 0 && n <= ctb::kEmdMaxPoints) ? ctb::emd_workspace_bytes(B, n) : 0;
}

int ctb_emd_fwd(const float* xyz1, const float* xyz2, float* dist, int32_t* assignment, void* workspace,
                size_t workspace_bytes, int B, int n, float eps, int iters, void* stream) {
  if (!xyz1 || !xyz2 || !dist || !assignment || B <= 0 || n <= 0 || iters <= 0) return CTB_ERR_INVALID_ARGUMENT;
  if (n > ctb::kEmdMaxPoints) return CTB_ERR_UNSUPPORTED;
  const size_t need = ctb::emd_workspace_bytes(B, n);
  if (need && (!workspace || workspace_bytes < need)) return CTB_ERR_WORKSPACE;
  return cuda_status(ctb::emd_forward(xyz1, xyz2, dist, assignment, workspace, B, n, eps, iters, (cudaStream_t)stream));
}

int ctb_emd_bwd(const float* xyz1, const float* xyz2, const float* grad_dist, const int32_t* assignment, float* grad_xyz1,
                int B, int n, void* stream) {
  if (!xyz1 || !xyz2 || !grad_dist || !assignment || !grad_xyz1 || B <= 0 || n <= 0) return CTB_ERR_INVALID_ARGUMENT;
  return cuda_status(ctb::emd_backward(xyz1, xyz2, grad_dist, assignment, grad_xyz1, B, n, (cudaStream_t)stream));
}

static int bn_exchange_of(const ctb_bn_exchange* ex, int C, ctb::BnExchange* out) {
  if (!ex || !ex->peer_data || !ex->peer_flag || !ex->epoch || !ex->done || !ex->scratch) return CTB_ERR_INVALID_ARGUMENT;
  if (ex->world < 1 || ex->world > 32 || ex->rank < 0 || ex->rank >= ex->world) return CTB_ERR_INVALID_ARGUMENT;
  out->peer_data = (float* const*)ex->peer_data;
  out->peer_flag = (unsigned* const*)ex->peer_flag;
  out->epoch = ex->epoch;
  out->done = ex->done;
  out->scratch = (float2*)ex->scratch;
  out->rank = ex->rank;
  out->world = ex->world;
  out->C = C;
  return CTB_OK;
}

static bool bn_vec4(int L, std::initializer_list<const void*> ptrs) {
  if (L % 4) return false;
  for (const void* p : ptrs)
    if ((uintptr_t)p & 15u) return false;
  return true;
}

uint64_t ctb_syncbn_scratch_bytes(int C) {
  return C > 0 ? (uint64_t)C * (ctb::kBnMaxChunks * sizeof(float2) + sizeof(unsigned)) : 0;
}

int ctb_syncbn_fwd(const float* x, const float* weight, const float* bias, float* y, float* save_mean, float* save_invstd,
                   float* running_mean, float* running_var, const ctb_bn_exchange* exchange, int B, int C, int L, float eps,
                   float momentum, void* stream) {
  if (!x || !y || !save_mean || !save_invstd || B <= 0 || C <= 0 || L <= 0) return CTB_ERR_INVALID_ARGUMENT;
  ctb::BnExchange ex;
  int st = bn_exchange_of(exchange, C, &ex);
  if (st) return st;
  const long long per = (long long)B * L;
  const dim3 gs(C, ctb::bn_chunks(C, per, ctb::kBnMaxChunks, 8192)), ga(C, ctb::bn_chunks(C, per, 65535, 4096));
  cudaStream_t sm = (cudaStream_t)stream;
  if (bn_vec4(L, {x, y})) {
    ctb::syncbn_fwd_stats_kernel<4><<<gs, ctb::kBnThreads, 0, sm>>>(x, ex, B, L);
    ctb::syncbn_fwd_apply_kernel<4><<<ga, ctb::kBnThreads, 0, sm>>>(x, weight, bias, y, save_mean, save_invstd, running_mean,
                                                                    running_var, ex, B, L, eps, momentum);
  } else {
    ctb::syncbn_fwd_stats_kernel<1><<<gs, ctb::kBnThreads, 0, sm>>>(x, ex, B, L);
    ctb::syncbn_fwd_apply_kernel<1><<<ga, ctb::kBnThreads, 0, sm>>>(x, weight, bias, y, save_mean, save_invstd, running_mean,
                                                                    running_var, ex, B, L, eps, momentum);
  }
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

int ctb_syncbn_bwd(const float* x, const float* grad_y, const float* weight, const float* save_mean, const float* save_invstd,
                   float* grad_x, float* grad_weight, float* grad_bias, const ctb_bn_exchange* exchange, int B, int C, int L,
                   void* stream) {
  if (!x || !grad_y || !save_mean || !save_invstd || !grad_x || B <= 0 || C <= 0 || L <= 0) return CTB_ERR_INVALID_ARGUMENT;
  ctb::BnExchange ex;
  int st = bn_exchange_of(exchange, C, &ex);
  if (st) return st;
  const long long per = (long long)B * L;
  const dim3 gs(C, ctb::bn_chunks(C, per, ctb::kBnMaxChunks, 8192)), ga(C, ctb::bn_chunks(C, per, 65535, 4096));
  cudaStream_t sm = (cudaStream_t)stream;
  if (bn_vec4(L, {x, grad_y, grad_x})) {
    ctb::syncbn_bwd_stats_kernel<4><<<gs, ctb::kBnThreads, 0, sm>>>(x, grad_y, save_mean, save_invstd, grad_weight, grad_bias,
                                                                    ex, B, L);
    ctb::syncbn_bwd_apply_kernel<4><<<ga, ctb::kBnThreads, 0, sm>>>(x, grad_y, weight, save_mean, save_invstd, grad_x, ex, B, L);
  } else {
    ctb::syncbn_bwd_stats_kernel<1><<<gs, ctb::kBnThreads, 0, sm>>>(x, grad_y, save_mean, save_invstd, grad_weight, grad_bias,
                                                                    ex, B, L);
    ctb::syncbn_bwd_apply_kernel<1><<<ga, ctb::kBnThreads, 0, sm>>>(x, grad_y, weight, save_mean, save_invstd, grad_x, ex, B, L);
  }
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

int ctb_count_occupied(const float* z, uint64_t n_elements, unsigned long long* count, void* stream) {
  if (!z || !count) return CTB_ERR_INVALID_ARGUMENT;
  if (n_elements == 0) return CTB_OK;
  unsigned blocks = (unsigned)((n_elements + 255) / 256);
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  ctb::count_occupied_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(z, n_elements, count);
  CTB_LAUNCH_CHECK();
  return CTB_OK;
}

}  // extern "C"
