// ctb_tile.cuh -- CTB_MODE_TILE kernels: the grid lives in CTA-owned shared-memory tiles.
//
// Measured on B200 (tools/microbench.cu, profiles/r01_microbench.txt): ATOMS.MAX.S32 on random shared
// addresses runs at ~11 lane-ops/clk/SM (= plain random LDS), shared float add is an ATOMS.CAST.SPIN loop
// at ~3, L2 atomics (REDG) reach ~0.67 and scattered 4-byte LDG ~1.0 lane-ops/clk/SM.  So both the
// scatters and the gathers resolve their 2^d random corner accesses per point ON CHIP and touch HBM only
// with coalesced 16-byte tile moves:
//
//   work item = (unit (b,h), channel group, slab of R grid rows along axis 0)
//   1. (slabs > 1) compact the indices of the points whose base row falls into the slab into shared
//      memory -- one cheap axis-0 test per point -- so the main loops run with full warps;
//   2. scatter: accumulate into the tile with native shared atomics, store the slab once
//      gather : load the slab (+1 halo row) once, read corners from the tile;
//   every cell of z / arg / grad_grid is written exactly once, no zero-fill pass, no L2 atomics.
//
// tile_scatter_kernel  A2+A3 Splat forward (reduce = max: pass 1 atomicMax on the int view of the positive
//                      products -- the reference's zero floor makes non-positive products irrelevant --,
//                      pass 2 atomicMin(e) among the entries equal to the maximum == torch-scatter's "first
//                      strictly greater in ascending e" rule; order independent => reproducible without a
//                      sort) and the grad_grid half of A5 Slice backward (reduce = sum, shared float atomics).
// tile_gather_kernel   A4 Slice forward, the grad_keys half of A5, A6 Splat backward (+A7 folded in).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ctb200.h"
#include "ctb_positions.cuh"

namespace ctb {

constexpr int kTileThreads = 512;
constexpr int kTileSmemTwoCtas = 110 * 1024;
constexpr int kTileSmemMax = 220 * 1024;
constexpr int kTileMaxPoints = 65535;  // compacted point lists are uint16

struct TileConfig {
  int FG;      // channels per tile
  int R;       // base rows (axis 0) per slab
  int slabs;
  int PPT;     // gather: points per thread (compile-time bound)
  size_t smem;
};

// Pick (FG, R): prefer ALL channels of the unit in one tile (positions are then computed once per point and
// pass), in as few balanced slabs as fit next to a second resident CTA; fall back to channel groups when
// even two rows of all channels do not fit.  halo = 1 for gathers (the +1 corner row), 0 for scatters.
inline bool tile_config(const ctb_shape* s, size_t per_cell, int halo, TileConfig* out) {
  const int stride0 = s->dim == 2 ? s->size[1] : s->size[1] * s->size[2];
  const int rows = s->size[0] - halo;          // base rows (gather) / destination rows (scatter)
  if (s->N > kTileMaxPoints) return false;
  const size_t list_bytes = (((size_t)s->N + 7) & ~(size_t)7) * 2 + 16;  // uint16 list + counter
  for (int pass = 0; pass < 2; ++pass) {
    const size_t budget = (pass == 0 ? kTileSmemTwoCtas : kTileSmemMax) - list_bytes;
    for (int FG = s->F; FG >= 1; FG = (FG + 1) / 2) {
      const int groups = (s->F + FG - 1) / FG;
      const int fg = (s->F + groups - 1) / groups;      // balanced group size
      const size_t row_bytes = (size_t)stride0 * fg * per_cell;
      int R = (int)(budget / row_bytes) - halo;
      const int min_rows = rows < 3 ? rows : 3;         // thinner slabs => too much halo / boundary rework
      if (R >= min_rows) {
        if (R > rows) R = rows;
        const int slabs = (rows + R - 1) / R;
        R = (rows + slabs - 1) / slabs;
        out->FG = fg;
        out->R = R;
        out->slabs = (rows + R - 1) / R;
        out->smem = (size_t)(R + halo) * row_bytes + list_bytes;
        return true;
      }
      if (FG == 1) break;
    }
  }
  return false;
}

// Shared-memory list of the points of this slab.  Returns the count; sel[i] is the point index.
// first_row / last_row: a point is selected if any of its two corner rows c0, c0+1 (scatter) or its base
// row c0 (gather) lies in [x0, x1).
template <int D, bool BOTH_ROWS>
__device__ __forceinline__ int compact_slab_points(const float* __restrict__ ku, int N, const Grid<D>& g, int x0, int x1,
                                                   unsigned short* sel, int* counter) {
  if (threadIdx.x == 0) *counter = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int n0 = 0; n0 < N; n0 += kTileThreads) {
    const int n = n0 + threadIdx.x;
    bool take = false;
    if (n < N) {
      bool in_rng;
      float up0, dn0;
      int c0;
      axis_pos<D>(__ldg(ku + n), g.scale[0], g.W[0], up0, dn0, c0, in_rng);
      take = (c0 >= x0 && c0 < x1) || (BOTH_ROWS && (c0 + 1 >= x0 && c0 + 1 < x1));
    }
    const unsigned m = __ballot_sync(0xffffffffu, take);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (take) sel[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)n;
  }
  __syncthreads();
  return *counter;
}

// ------------------------------------------------------------------------------------------------------
template <int D, bool SUM, bool VEC4>
__global__ void __launch_bounds__(kTileThreads)
tile_scatter_kernel(const float* __restrict__ keys, const float* __restrict__ feat, const float* __restrict__ pad,
                    float* __restrict__ z, int* __restrict__ arg, Grid<D> g, int H, int F, int N, int FG, int R,
                    int slabs, int groups) {
  constexpr int S = 1 << D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int stride0 = g.stride[0];
  const int W0 = g.W[0];
  const int tile_cells = R * stride0;
  const bool want_arg = !SUM && arg != nullptr;
  float* tval = (float*)smem_raw;                                   // [fg][tile_cells]
  int* targ = (int*)(tval + (size_t)FG * tile_cells);               // [fg][tile_cells]  (max with arg only)
  unsigned short* sel = (unsigned short*)(targ + (want_arg ? (size_t)FG * tile_cells : 0));
  int* counter = (int*)(sel + ((N + 7) & ~7));

  int item = blockIdx.x;
  const int slab = item % slabs;
  item /= slabs;
  const int f0 = (item % groups) * FG;
  const int unit = item / groups;
  const int fg = min(FG, F - f0);
  const int x0 = slab * R, x1 = min(x0 + R, W0);
  const int cell0 = x0 * stride0;
  const int ncell = (x1 - x0) * stride0;

  if constexpr (VEC4) {
    float4* t4 = reinterpret_cast<float4*>(tval);
    int4* a4 = reinterpret_cast<int4*>(targ);
    for (int i = threadIdx.x; i < (fg * tile_cells) >> 2; i += kTileThreads) {
      t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (want_arg) a4[i] = make_int4(0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF);
    }
  } else {
    for (int i = threadIdx.x; i < fg * tile_cells; i += kTileThreads) {
      tval[i] = 0.0f;
      if (want_arg) targ[i] = 0x7FFFFFFF;
    }
  }
  const float* ku = keys + (size_t)unit * D * N;
  const float* pu = pad ? pad + (size_t)(unit / H) * N : nullptr;
  const float* fu = feat + ((size_t)unit * F + f0) * N;
  int cnt = N;
  if (slabs > 1) cnt = compact_slab_points<D, true>(ku, N, g, x0, x1, sel, counter);
  else __syncthreads();

#pragma unroll 1
  for (int pass = 0; pass < (want_arg ? 2 : 1); ++pass) {
    for (int i = threadIdx.x; i < cnt; i += kTileThreads) {
      const int n = slabs > 1 ? (int)sel[i] : i;
      const Pos<D> p = point_pos<D>(ku, n, N, g);
      const bool in0 = (p.c0 >= x0) && (p.c0 < x1);
      const bool in1 = (p.c0 + 1 >= x0) && (p.c0 + 1 < x1);
      float w[S];
      int lc[S];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        w[s] = corner_weight<D>(p, s);
        lc[s] = ((s & 1) ? in1 : in0) ? p.base + corner_offset<D>(g, s) - cell0 : -1;
      }
      const float pd = pu ? __ldg(pu + n) : 1.0f;
      for (int f = 0; f < fg; ++f) {
        float ft = __ldg(fu + (size_t)f * N + n);
        if (pu) ft = CTB_FMUL(ft, pd);
        float* tf = tval + (size_t)f * tile_cells;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          if (lc[s] < 0) continue;
          const float v = CTB_FMUL(ft, w[s]);
          if constexpr (SUM) {
            atomicAdd(tf + lc[s], v);
          } else {
            if (v > 0.0f) {
              if (pass == 0) {
                atomicMax((int*)tf + lc[s], __float_as_int(v));
              } else if (__float_as_int(v) == ((const int*)tf)[lc[s]]) {
                atomicMin(targ + (size_t)f * tile_cells + lc[s], s * N + n);
              }
            }
          }
        }
      }
    }
    __syncthreads();
  }

  // one coalesced store of the slab: z (and arg, with the "no winner" marker mapped to -1)
  if constexpr (VEC4) {
    const int n4 = ncell >> 2;
    for (int i = threadIdx.x; i < fg * n4; i += kTileThreads) {
      const int f = i / n4, r = i - f * n4;
      const size_t go = ((size_t)unit * F + f0 + f) * g.C + cell0;
      __stcs(reinterpret_cast<float4*>(z + go) + r, reinterpret_cast<const float4*>(tval + (size_t)f * tile_cells)[r]);
      if (want_arg) {
        int4 a = reinterpret_cast<const int4*>(targ + (size_t)f * tile_cells)[r];
        a.x = a.x == 0x7FFFFFFF ? -1 : a.x;
        a.y = a.y == 0x7FFFFFFF ? -1 : a.y;
        a.z = a.z == 0x7FFFFFFF ? -1 : a.z;
        a.w = a.w == 0x7FFFFFFF ? -1 : a.w;
        __stcs(reinterpret_cast<int4*>(arg + go) + r, a);
      }
    }
  } else {
    for (int i = threadIdx.x; i < fg * ncell; i += kTileThreads) {
      const int f = i / ncell, r = i - f * ncell;
      const size_t go = ((size_t)unit * F + f0 + f) * g.C + cell0 + r;
      z[go] = tval[(size_t)f * tile_cells + r];
      if (want_arg) {
        const int a = targ[(size_t)f * tile_cells + r];
        arg[go] = a == 0x7FFFFFFF ? -1 : a;
      }
    }
  }
}

inline bool tile_scatter_config(const ctb_shape* s, bool sum, bool want_arg, TileConfig* out) {
  return tile_config(s, (sum || !want_arg) ? 4 : 8, 0, out);
}

template <int D>
cudaError_t tile_scatter(const float* keys, const float* feat, const float* pad, float* z, int* arg,
                         const ctb_shape* s, bool sum, cudaStream_t stream) {
  TileConfig c;
  if (!tile_scatter_config(s, sum, arg != nullptr, &c)) return cudaErrorNotSupported;
  const Grid<D> g = make_grid<D>(s->size);
  const int groups = (s->F + c.FG - 1) / c.FG;
  const long long blocks = (long long)s->B * s->H * groups * c.slabs;
  if (blocks >= (1ll << 31)) return cudaErrorNotSupported;
  const bool vec4 = (g.C % 4 == 0) && (g.stride[0] % 4 == 0) && ((reinterpret_cast<uintptr_t>(z) & 15) == 0) &&
                    (arg == nullptr || (reinterpret_cast<uintptr_t>(arg) & 15) == 0);
#define CTB_TS(SUMV, VECV)                                                                                          \
  do {                                                                                                              \
    cudaError_t e = cudaFuncSetAttribute(tile_scatter_kernel<D, SUMV, VECV>,                                        \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);                 \
    if (e != cudaSuccess) return e;                                                                                 \
    tile_scatter_kernel<D, SUMV, VECV><<<(unsigned)blocks, kTileThreads, c.smem, stream>>>(                         \
        keys, feat, pad, z, arg, g, s->H, s->F, s->N, c.FG, c.R, c.slabs, groups);                                  \
    return cudaGetLastError();                                                                                      \
  } while (0)
  if (sum) {
    if (vec4) CTB_TS(true, true); else CTB_TS(true, false);
  } else {
    if (vec4) CTB_TS(false, true); else CTB_TS(false, false);
  }
#undef CTB_TS
}

// ------------------------------------------------------------------------------------------------------
enum GatherMode { GATHER_SLICE_FWD = 0, GATHER_SLICE_BWD_KEYS = 1, GATHER_SPLAT_BWD = 2 };

// One CTA = (unit, slab).  Loops over channel groups; every point is resolved in the single slab that holds
// its base row, so grad_keys needs no cross-CTA reduction.
template <int D, int MODE, int PPT, bool VEC4>
__global__ void __launch_bounds__(kTileThreads, 2)
tile_gather_kernel(const float* __restrict__ keys, const float* __restrict__ t1, const int* __restrict__ t2,
                   const float* __restrict__ in, const float* __restrict__ pad, float* __restrict__ out,
                   float* __restrict__ grad_keys, Grid<D> g, int H, int F, int N, int FG, int R, int slabs) {
  constexpr int S = 1 << D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int stride0 = g.stride[0];
  const int W0 = g.W[0];
  const int tile_cells = (R + 1) * stride0;
  float* s1 = (float*)smem_raw;                                     // [FG][tile_cells]
  int* s2 = (int*)(s1 + (size_t)FG * tile_cells);                   // [FG][tile_cells]  (SPLAT_BWD: arg)
  unsigned short* sel = (unsigned short*)(s2 + (MODE == GATHER_SPLAT_BWD ? (size_t)FG * tile_cells : 0));
  int* counter = (int*)(sel + ((N + 7) & ~7));

  const int slab = blockIdx.x % slabs;
  const int unit = blockIdx.x / slabs;
  const int x0 = slab * R;
  const int x1 = min(x0 + R, W0 - 1);           // base rows [x0, x1)
  const int xe = min(x1 + 1, W0);               // tile rows [x0, xe)
  const int cell0 = x0 * stride0;
  const int ncell = (xe - x0) * stride0;
  const float* ku = keys + (size_t)unit * D * N;
  const float* pu = pad ? pad + (size_t)(unit / H) * N : nullptr;

  int cnt = N;
  if (slabs > 1) cnt = compact_slab_points<D, false>(ku, N, g, x0, x1, sel, counter);

  float gk[PPT][D];
#pragma unroll
  for (int k = 0; k < PPT; ++k)
#pragma unroll
    for (int a = 0; a < D; ++a) gk[k][a] = 0.0f;

  for (int f0 = 0; f0 < F; f0 += FG) {
    const int fg = min(FG, F - f0);
    if (f0 > 0) __syncthreads();                 // previous group's readers are done with the tile
    if constexpr (VEC4) {
      const int n4 = ncell >> 2;
      for (int i = threadIdx.x; i < fg * n4; i += kTileThreads) {
        const int f = i / n4, r = i - f * n4;
        const size_t go = ((size_t)unit * F + f0 + f) * g.C + cell0;
        reinterpret_cast<float4*>(s1 + (size_t)f * tile_cells)[r] = __ldcs(reinterpret_cast<const float4*>(t1 + go) + r);
        if constexpr (MODE == GATHER_SPLAT_BWD)
          reinterpret_cast<int4*>(s2 + (size_t)f * tile_cells)[r] = __ldcs(reinterpret_cast<const int4*>(t2 + go) + r);
      }
    } else {
      for (int i = threadIdx.x; i < fg * ncell; i += kTileThreads) {
        const int f = i / ncell, r = i - f * ncell;
        const size_t go = ((size_t)unit * F + f0 + f) * g.C + cell0 + r;
        s1[(size_t)f * tile_cells + r] = __ldg(t1 + go);
        if constexpr (MODE == GATHER_SPLAT_BWD) s2[(size_t)f * tile_cells + r] = __ldg(t2 + go);
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const int i = threadIdx.x + k * kTileThreads;
      if (i >= cnt) break;
      const int n = slabs > 1 ? (int)sel[i] : i;
      const Pos<D> p = point_pos<D>(ku, n, N, g);
      const float pd = pu ? __ldg(pu + n) : 1.0f;
      float w[S];
      int lc[S];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        w[s] = corner_weight<D>(p, s);
        lc[s] = p.base + corner_offset<D>(g, s) - cell0;
      }
      float gw[S];
#pragma unroll
      for (int s = 0; s < S; ++s) gw[s] = 0.0f;
      for (int f = 0; f < fg; ++f) {
        const float* tf = s1 + (size_t)f * tile_cells;
        const size_t po = ((size_t)unit * F + f0 + f) * N + n;
        if constexpr (MODE == GATHER_SLICE_FWD) {
          float acc = 0.0f;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const float t = CTB_FMUL(tf[lc[s]], w[s]);
            acc = (s == 0) ? t : CTB_FADD(acc, t);
          }
          if (pu) acc = CTB_FMUL(acc, pd);
          out[po] = acc;
        } else if constexpr (MODE == GATHER_SLICE_BWD_KEYS) {
          float go = __ldg(in + po);
          if (pu) go *= pd;
#pragma unroll
          for (int s = 0; s < S; ++s) gw[s] = fmaf(tf[lc[s]], go, gw[s]);
        } else {
          const int* af = s2 + (size_t)f * tile_cells;
          float ft = __ldg(in + po);
          if (pu) ft *= pd;
          float gf = 0.0f;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            if (af[lc[s]] == s * N + n) {
              const float gz = tf[lc[s]];
              gf = fmaf(gz, w[s], gf);
              gw[s] = fmaf(gz, ft, gw[s]);
            }
          }
          if (pu) gf *= pd;
          out[po] = gf;
        }
      }
      if constexpr (MODE != GATHER_SLICE_FWD) {
        float part[D];
        weight_grad_to_key_grad<D>(p, gw, part);
#pragma unroll
        for (int a = 0; a < D; ++a) gk[k][a] += part[a];
      }
    }
  }
  if constexpr (MODE != GATHER_SLICE_FWD) {
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const int i = threadIdx.x + k * kTileThreads;
      if (i >= cnt) break;
      const int n = slabs > 1 ? (int)sel[i] : i;
#pragma unroll
      for (int a = 0; a < D; ++a) grad_keys[((size_t)unit * D + a) * N + n] = gk[k][a];
    }
  }
}

inline bool gather_config(const ctb_shape* s, int mode, TileConfig* out) {
  // worst case every point of a unit falls into one slab => the per-thread point bound covers N
  const int ppt = (s->N + kTileThreads - 1) / kTileThreads;
  if (ppt > 8) return false;
  if (!tile_config(s, mode == GATHER_SPLAT_BWD ? 8 : 4, 1, out)) return false;
  out->PPT = ppt <= 1 ? 1 : (ppt <= 2 ? 2 : (ppt <= 4 ? 4 : 8));
  return true;
}

template <int D, int MODE, int PPT>
cudaError_t launch_gather(const float* keys, const float* t1, const int* t2, const float* in, const float* pad,
                          float* out, float* grad_keys, const ctb_shape* s, const TileConfig& c, cudaStream_t stream) {
  const Grid<D> g = make_grid<D>(s->size);
  const bool vec4 = (g.C % 4 == 0) && (g.stride[0] % 4 == 0) && ((reinterpret_cast<uintptr_t>(t1) & 15) == 0) &&
                    (MODE != GATHER_SPLAT_BWD || (reinterpret_cast<uintptr_t>(t2) & 15) == 0);
  const long long blocks = (long long)s->B * s->H * c.slabs;
  if (blocks >= (1ll << 31)) return cudaErrorNotSupported;
  cudaError_t e;
  if (vec4) {
    e = cudaFuncSetAttribute(tile_gather_kernel<D, MODE, PPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)c.smem);
    if (e != cudaSuccess) return e;
    tile_gather_kernel<D, MODE, PPT, true><<<(unsigned)blocks, kTileThreads, c.smem, stream>>>(
        keys, t1, t2, in, pad, out, grad_keys, g, s->H, s->F, s->N, c.FG, c.R, c.slabs);
  } else {
    e = cudaFuncSetAttribute(tile_gather_kernel<D, MODE, PPT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)c.smem);
    if (e != cudaSuccess) return e;
    tile_gather_kernel<D, MODE, PPT, false><<<(unsigned)blocks, kTileThreads, c.smem, stream>>>(
        keys, t1, t2, in, pad, out, grad_keys, g, s->H, s->F, s->N, c.FG, c.R, c.slabs);
  }
  return cudaGetLastError();
}

template <int D, int MODE>
cudaError_t tile_gather(const float* keys, const float* t1, const int* t2, const float* in, const float* pad,
                        float* out, float* grad_keys, const ctb_shape* s, cudaStream_t stream) {
  TileConfig c;
  if (!gather_config(s, MODE, &c)) return cudaErrorNotSupported;
  switch (c.PPT) {
    case 1: return launch_gather<D, MODE, 1>(keys, t1, t2, in, pad, out, grad_keys, s, c, stream);
    case 2: return launch_gather<D, MODE, 2>(keys, t1, t2, in, pad, out, grad_keys, s, c, stream);
    case 4: return launch_gather<D, MODE, 4>(keys, t1, t2, in, pad, out, grad_keys, s, c, stream);
    default: return launch_gather<D, MODE, 8>(keys, t1, t2, in, pad, out, grad_keys, s, c, stream);
  }
}

}  // namespace ctb
