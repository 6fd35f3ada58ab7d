// ctb_generic.cuh -- point-stationary kernels (one thread per point, loop over channels).
//
// These are the shape-agnostic kernels: any grid size, any N, any F.  The scatters use L2 atomics
// (red.global.max.s32 on the int view of positive floats / red.global.add.f32); the gathers read the
// NCHW grid directly.  They implement CTB_MODE_ATOMIC and serve the reference-API entries where the
// caller hands in lc / idx tensors (policy KEYS = false) as well as the fused entries (KEYS = true).
//
// Reference rows: A1/A7 positions, A2+A3 Splat fwd, A4 Slice fwd, A5 Slice bwd,
// A6 Splat bwd  --  layers/cloud_transform.py:72-121, :131-180, :190-227.
#pragma once
#include <cuda_runtime.h>
#include "ctb_positions.cuh"

namespace ctb {

constexpr int kGenericBlock = 256;

// where a kernel gets the per-point corner weights / cell indices from
struct PointSource {
  const float* keys;     // KEYS:  [B, H*D, N]
  const float* lc;       // !KEYS: [B, H, S, N]
  const int64_t* idx;    // !KEYS: [B, H, S, N]
};

template <int D, bool KEYS>
struct Corners {
  float w[1 << D];
  int cell[1 << D];   // -1 = out of range (skipped; only possible for caller-provided idx)
  Pos<D> pos;         // valid only when KEYS
};

template <int D, bool KEYS>
__device__ __forceinline__ void load_corners(const PointSource& src, const Grid<D>& g, int unit, int n, int N,
                                             Corners<D, KEYS>& c) {
  constexpr int S = 1 << D;
  if constexpr (KEYS) {
    c.pos = point_pos<D>(src.keys + (size_t)unit * D * N, n, N, g);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      c.w[s] = corner_weight<D>(c.pos, s);
      c.cell[s] = c.pos.base + corner_offset<D>(g, s);
    }
  } else {
    const size_t o = (size_t)unit * S * N + n;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      c.w[s] = __ldg(src.lc + o + (size_t)s * N);
      const long long i = __ldg((const long long*)src.idx + o + (size_t)s * N);
      c.cell[s] = (i >= 0 && i < (long long)g.C) ? (int)i : -1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// A1: positions forward (materialises lc / idx for the reference API)
template <int D>
__global__ void __launch_bounds__(kGenericBlock)
positions_fwd_kernel(const float* __restrict__ keys, float* __restrict__ lc, long long* __restrict__ idx,
                     Grid<D> g, int N, int chunks) {
  constexpr int S = 1 << D;
  const int unit = blockIdx.x / chunks;
  const int n = (blockIdx.x % chunks) * kGenericBlock + threadIdx.x;
  if (n >= N) return;
  const Pos<D> p = point_pos<D>(keys + (size_t)unit * D * N, n, N, g);
  const size_t o = (size_t)unit * S * N + n;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    lc[o + (size_t)s * N] = corner_weight<D>(p, s);
    idx[o + (size_t)s * N] = (long long)(p.base + corner_offset<D>(g, s));
  }
}

// A7: positions backward
template <int D>
__global__ void __launch_bounds__(kGenericBlock)
positions_bwd_kernel(const float* __restrict__ keys, const float* __restrict__ grad_lc,
                     float* __restrict__ grad_keys, Grid<D> g, int N, int chunks) {
  constexpr int S = 1 << D;
  const int unit = blockIdx.x / chunks;
  const int n = (blockIdx.x % chunks) * kGenericBlock + threadIdx.x;
  if (n >= N) return;
  const Pos<D> p = point_pos<D>(keys + (size_t)unit * D * N, n, N, g);
  float gw[S], gk[D];
#pragma unroll
  for (int s = 0; s < S; ++s) gw[s] = __ldg(grad_lc + ((size_t)unit * S + s) * N + n);
  weight_grad_to_key_grad<D>(p, gw, gk);
#pragma unroll
  for (int a = 0; a < D; ++a) grad_keys[((size_t)unit * D + a) * N + n] = gk[a];
}

// ---------------------------------------------------------------------------------------------
// A2+A3: Splat forward, atomic.  PHASE 0: max (or sum) into z.  PHASE 1 (max only): arg = min e among
// the entries equal to the cell's final value -- exactly torch-scatter's CPU rule "first strictly
// greater in ascending e".
template <int D, bool KEYS, int PHASE, bool SUM>
__global__ void __launch_bounds__(kGenericBlock)
splat_fwd_atomic_kernel(PointSource src, const float* __restrict__ feat, const float* __restrict__ pad,
                        float* __restrict__ z, int* __restrict__ arg, Grid<D> g, int H, int F, int N,
                        int chunks) {
  constexpr int S = 1 << D;
  const int unit = blockIdx.x / chunks;
  const int n = (blockIdx.x % chunks) * kGenericBlock + threadIdx.x;
  if (n >= N) return;
  Corners<D, KEYS> c;
  load_corners<D, KEYS>(src, g, unit, n, N, c);
  const float pd = pad ? __ldg(pad + (size_t)(unit / H) * N + n) : 1.0f;
  const float* fu = feat + (size_t)unit * F * N + n;
  float* zu = z + (size_t)unit * F * g.C;
  int* au = arg ? arg + (size_t)unit * F * g.C : nullptr;
  for (int f = 0; f < F; ++f) {
    float ft = __ldg(fu + (size_t)f * N);
    if (pad) ft = CTB_FMUL(ft, pd);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if (c.cell[s] < 0) continue;
      const float v = CTB_FMUL(ft, c.w[s]);
      const size_t o = (size_t)f * g.C + c.cell[s];
      if constexpr (SUM) {
        if (v != 0.0f) atomicAdd(zu + o, v);
      } else if constexpr (PHASE == 0) {
        if (v > 0.0f) atomicMax((int*)(zu + o), __float_as_int(v));
      } else {
        if (v > 0.0f && v == zu[o]) atomicMin((unsigned*)(au + o), (unsigned)(s * N + n));
      }
    }
  }
}

// A4: Slice forward (gather).  Sequential float32 accumulation over corners (first product, then FMAs).
template <int D, bool KEYS>
__global__ void __launch_bounds__(kGenericBlock)
slice_fwd_kernel(PointSource src, const float* __restrict__ grid, const float* __restrict__ pad,
                 float* __restrict__ out, Grid<D> g, int H, int F, int N, int chunks) {
  constexpr int S = 1 << D;
  const int unit = blockIdx.x / chunks;
  const int n = (blockIdx.x % chunks) * kGenericBlock + threadIdx.x;
  if (n >= N) return;
  Corners<D, KEYS> c;
  load_corners<D, KEYS>(src, g, unit, n, N, c);
  const float pd = pad ? __ldg(pad + (size_t)(unit / H) * N + n) : 1.0f;
  const float* gu = grid + (size_t)unit * F * g.C;
  float* ou = out + (size_t)unit * F * N + n;
  for (int f = 0; f < F; ++f) {
    float acc = 0.0f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float gv = c.cell[s] >= 0 ? __ldg(gu + (size_t)f * g.C + c.cell[s]) : 0.0f;
      acc = (s == 0) ? CTB_FMUL(gv, c.w[s]) : fmaf(gv, c.w[s], acc);
    }
    if (pad) acc = CTB_FMUL(acc, pd);
    ou[(size_t)f * N] = acc;
  }
}

// A5: Slice backward.  grad_grid (pre-zeroed) += lc * grad_out * pad (atomic scatter-add);
// grad_lc[s] = sum_f grid[cell_s] * grad_out * pad, written as grad_lc (!KEYS) or folded to grad_keys.
// SCATTER = false skips the scatter-add (used when a deterministic kernel produces grad_grid).
template <int D, bool KEYS, bool SCATTER>
__global__ void __launch_bounds__(kGenericBlock)
slice_bwd_kernel(PointSource src, const float* __restrict__ grid, const float* __restrict__ pad,
                 const float* __restrict__ grad_out, float* __restrict__ grad_grid,
                 float* __restrict__ grad_pos, Grid<D> g, int H, int F, int N, int chunks) {
  constexpr int S = 1 << D;
  const int unit = blockIdx.x / chunks;
  const int n = (blockIdx.x % chunks) * kGenericBlock + threadIdx.x;
  if (n >= N) return;
  Corners<D, KEYS> c;
  load_corners<D, KEYS>(src, g, unit, n, N, c);
  const float pd = pad ? __ldg(pad + (size_t)(unit / H) * N + n) : 1.0f;
  const float* gu = grid + (size_t)unit * F * g.C;
  float* ggu = grad_grid + (size_t)unit * F * g.C;
  const float* gou = grad_out + (size_t)unit * F * N + n;
  float gw[S];
#pragma unroll
  for (int s = 0; s < S; ++s) gw[s] = 0.0f;
  for (int f = 0; f < F; ++f) {
    float go = __ldg(gou + (size_t)f * N);
    if (pad) go *= pd;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if (c.cell[s] < 0) continue;
      const size_t o = (size_t)f * g.C + c.cell[s];
      gw[s] = fmaf(__ldg(gu + o), go, gw[s]);
      if constexpr (SCATTER) {
        const float t = c.w[s] * go;
        if (t != 0.0f) atomicAdd(ggu + o, t);
      }
    }
  }
  if constexpr (KEYS) {
    float gk[D];
    weight_grad_to_key_grad<D>(c.pos, gw, gk);
#pragma unroll
    for (int a = 0; a < D; ++a) grad_pos[((size_t)unit * D + a) * N + n] = gk[a];
  } else {
#pragma unroll
    for (int s = 0; s < S; ++s) grad_pos[((size_t)unit * S + s) * N + n] = gw[s];
  }
}

// A6: Splat backward.  MAX: the gradient of a cell goes to its single arg winner; SUM: to every entry.
template <int D, bool KEYS, bool SUM>
__global__ void __launch_bounds__(kGenericBlock)
splat_bwd_kernel(PointSource src, const float* __restrict__ feat, const float* __restrict__ pad,
                 const float* __restrict__ grad_z, const int* __restrict__ arg,
                 float* __restrict__ grad_feat, float* __restrict__ grad_pos, Grid<D> g, int H, int F,
                 int N, int chunks) {
  constexpr int S = 1 << D;
  const int unit = blockIdx.x / chunks;
  const int n = (blockIdx.x % chunks) * kGenericBlock + threadIdx.x;
  if (n >= N) return;
  Corners<D, KEYS> c;
  load_corners<D, KEYS>(src, g, unit, n, N, c);
  const float pd = pad ? __ldg(pad + (size_t)(unit / H) * N + n) : 1.0f;
  const float* fu = feat + (size_t)unit * F * N + n;
  const float* gzu = grad_z + (size_t)unit * F * g.C;
  const int* au = SUM ? nullptr : arg + (size_t)unit * F * g.C;
  float* gfu = grad_feat + (size_t)unit * F * N + n;
  float gw[S];
#pragma unroll
  for (int s = 0; s < S; ++s) gw[s] = 0.0f;
  for (int f = 0; f < F; ++f) {
    float ft = __ldg(fu + (size_t)f * N);
    if (pad) ft *= pd;
    float gf = 0.0f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if (c.cell[s] < 0) continue;
      const size_t o = (size_t)f * g.C + c.cell[s];
      bool win = true;
      if constexpr (!SUM) win = (__ldg(au + o) == s * N + n);
      if (win) {
        const float gz = __ldg(gzu + o);
        gf = fmaf(gz, c.w[s], gf);
        gw[s] = fmaf(gz, ft, gw[s]);
      }
    }
    if (pad) gf *= pd;
    gfu[(size_t)f * N] = gf;
  }
  if constexpr (KEYS) {
    float gk[D];
    weight_grad_to_key_grad<D>(c.pos, gw, gk);
#pragma unroll
    for (int a = 0; a < D; ++a) grad_pos[((size_t)unit * D + a) * N + n] = gk[a];
  } else {
#pragma unroll
    for (int s = 0; s < S; ++s) grad_pos[((size_t)unit * S + s) * N + n] = gw[s];
  }
}

// A9: occupancy count (multihead_ct.py:104-105)
__global__ void __launch_bounds__(256)
count_occupied_kernel(const float* __restrict__ z, unsigned long long n, unsigned long long* count) {
  unsigned long long local = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x)
    local += fabsf(__ldg(z + i)) > 1e-9f ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

}  // namespace ctb
