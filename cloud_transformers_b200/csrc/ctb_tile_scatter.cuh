// ctb_tile_scatter.cuh -- CTB_MODE_TILE scatters: point-stationary accumulation into a CTA-owned
// shared-memory tile with NATIVE shared-memory atomics, then one coalesced store of the tile.
//
// Measured on B200 (tools/microbench.cu, profiles/r01_microbench.txt): ATOMS.MAX.S32 on random addresses
// runs at ~11 lane-ops/clk/SM (the same rate as plain random LDS), shared float add is an
// ATOMS.CAST.SPIN loop at ~3 lane-ops/clk/SM, and L2 atomics (REDG) reach only ~0.67 lane-ops/clk/SM.
// So the scatter is done on chip:
//   * Splat forward, reduce = max (A2+A3): pass 1 atomicMax on the int view of the positive products
//     (the reference's zero floor makes non-positive products irrelevant), pass 2 resolves the winner as
//     atomicMin(e) among the entries equal to the cell's maximum == torch-scatter's "first strictly
//     greater in ascending e" rule.  Order independent => bit-reproducible without any sort.
//   * reduce = sum (Splat-sum, grad_grid of Slice backward A5): shared-memory float atomicAdd
//     (summation order depends on warp scheduling, like the reference's own atomicAdd scatter).
// The grid never sees a zero-fill pass or an L2 atomic: every cell of z / arg / grad_grid is written
// exactly once with 16-byte coalesced stores.  Work item = (unit, channel group, slab of grid rows).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ctb200.h"
#include "ctb_positions.cuh"

namespace ctb {

constexpr int kTileScatterThreads = 512;
constexpr int kTileSmemTwoCtas = 110 * 1024;
constexpr int kTileSmemMax = 220 * 1024;

struct TileScatterConfig {
  int FG;      // channels per work item
  int R;       // grid rows (axis 0) per slab
  int slabs;
  size_t smem;
};

inline bool tile_scatter_config(const ctb_shape* s, bool sum, bool want_arg, TileScatterConfig* out) {
  const int stride0 = s->dim == 2 ? s->size[1] : s->size[1] * s->size[2];
  const int W0 = s->size[0];
  const size_t per_cell = (sum || !want_arg) ? 4 : 8;
  const size_t plane = (size_t)W0 * stride0 * per_cell;
  if (plane <= (size_t)kTileSmemTwoCtas) {
    // whole planes: as many channels as fit next to a second CTA, split evenly
    int FG = (int)((size_t)kTileSmemTwoCtas / plane);
    if (FG > s->F) FG = s->F;
    if (FG > 16) FG = 16;
    const int groups = (s->F + FG - 1) / FG;
    FG = (s->F + groups - 1) / groups;
    out->FG = FG;
    out->R = W0;
    out->slabs = 1;
    out->smem = (size_t)FG * plane;
    return true;
  }
  for (int pass = 0; pass < 2; ++pass) {
    const size_t budget = pass == 0 ? kTileSmemTwoCtas : kTileSmemMax;
    int R = (int)(budget / ((size_t)stride0 * per_cell));
    if (R < 1) continue;
    if (R > W0) R = W0;
    const int slabs = (W0 + R - 1) / R;
    if (pass == 0 && slabs > 4) continue;  // too many re-scans of the points: take the big-tile budget
    R = (W0 + slabs - 1) / slabs;
    out->FG = 1;
    out->R = R;
    out->slabs = (W0 + R - 1) / R;
    out->smem = (size_t)R * stride0 * per_cell;
    return true;
  }
  return false;
}

template <int D, bool SUM, bool VEC4>
__global__ void __launch_bounds__(kTileScatterThreads)
tile_scatter_kernel(const float* __restrict__ keys, const float* __restrict__ feat, const float* __restrict__ pad,
                    float* __restrict__ z, int* __restrict__ arg, Grid<D> g, int H, int F, int N, int FG, int R,
                    int slabs, int groups) {
  constexpr int S = 1 << D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int stride0 = g.stride[0];
  const int W0 = g.W[0];
  const int tile_cells = R * stride0;
  float* tval = (float*)smem_raw;                          // [fg][tile_cells]
  int* targ = (int*)(tval + (size_t)FG * tile_cells);      // [fg][tile_cells]  (max with arg only)
  const bool want_arg = !SUM && arg != nullptr;

  int item = blockIdx.x;
  const int slab = item % slabs;
  item /= slabs;
  const int f0 = (item % groups) * FG;
  const int unit = item / groups;
  const int fg = min(FG, F - f0);
  const int x0 = slab * R, x1 = min(x0 + R, W0);
  const int cell0 = x0 * stride0;
  const int ncell = (x1 - x0) * stride0;

  for (int i = threadIdx.x; i < fg * tile_cells; i += kTileScatterThreads) {
    tval[i] = 0.0f;
    if (want_arg) targ[i] = 0x7FFFFFFF;
  }
  __syncthreads();

  const float* ku = keys + (size_t)unit * D * N;
  const float* pu = pad ? pad + (size_t)(unit / H) * N : nullptr;
  const float* fu = feat + ((size_t)unit * F + f0) * N;

#pragma unroll 1
  for (int pass = 0; pass < (want_arg ? 2 : 1); ++pass) {
    for (int n = threadIdx.x; n < N; n += kTileScatterThreads) {
      // cheap slab test on axis 0 first, full position only for points that touch this slab
      bool in_rng;
      float up0, dn0;
      int c0;
      axis_pos<D>(__ldg(ku + n), g.scale[0], g.W[0], up0, dn0, c0, in_rng);
      const bool in0 = (c0 >= x0) && (c0 < x1);
      const bool in1 = (c0 + 1 >= x0) && (c0 + 1 < x1);
      if (!in0 && !in1) continue;
      const Pos<D> p = point_pos<D>(ku, n, N, g);
      float w[S];
      int lc[S];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        w[s] = corner_weight<D>(p, s);
        lc[s] = p.base + corner_offset<D>(g, s) - cell0;
      }
      const float pd = pu ? __ldg(pu + n) : 1.0f;
      for (int f = 0; f < fg; ++f) {
        float ft = __ldg(fu + (size_t)f * N + n);
        if (pu) ft = CTB_FMUL(ft, pd);
        float* tf = tval + (size_t)f * tile_cells;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          if (!((s & 1) ? in1 : in0)) continue;
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
    for (int i = threadIdx.x; i < fg * n4; i += kTileScatterThreads) {
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
    for (int i = threadIdx.x; i < fg * ncell; i += kTileScatterThreads) {
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

template <int D>
cudaError_t tile_scatter(const float* keys, const float* feat, const float* pad, float* z, int* arg,
                         const ctb_shape* s, bool sum, cudaStream_t stream) {
  TileScatterConfig c;
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
    tile_scatter_kernel<D, SUMV, VECV><<<(unsigned)blocks, kTileScatterThreads, c.smem, stream>>>(                  \
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

}  // namespace ctb
