// ctb_sgather.cuh -- Slice forward of the dense shape classes with the points taken in the plan's SORTED order.
//
// tile_gather_kernel (ctb_tile.cuh) walks the points of a unit in index order: the 32 lanes of a warp read 32
// unrelated cells of the shared-memory tile per corner and channel -- a 3.5-way bank conflict on average, and the
// load/store unit is what that kernel waits for on the coarse grids (8^3 x F32, 16^2 x F16, 16^3 x F16; ncu: 66 %
// LSU busy).  The plan (ctb_plan.cuh) already holds the unit's points sorted by base cell.  Walking THAT order, the
// lanes of a warp sit in the same or in neighbouring cells: a corner read is a broadcast or a run of consecutive
// words.  A point's result per channel would be a scattered 4-byte global store now, so the results are written at
// slot n of a row staged in shared memory and leave as whole rows by the copy engine, like the tile arrives.
//
//   CTA = (unit, group of FG channels);  shared: tile [FG][C], rows out [FG][Np];  two CTAs per SM
//
// The arithmetic per (point, channel) is that of tile_gather_kernel in the same order: bit-identical results
// (layers/cloud_transform.py:204-211).  The backward gathers were tried in this form and stay with tile_gather_kernel
// (DESIGN.md section 7): they accumulate grad_keys over ALL channels of a point, so one CTA walks the channel groups
// one after the other and the exposed copy-engine latency per group costs more than the conflicts it removes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ctb200.h"
#include "ctb_plan.cuh"
#include "ctb_positions.cuh"
#include "ctb_tile.cuh"

namespace ctb {

constexpr int kSortThreads = 512;
constexpr size_t kSortSmemTwoCtas = 110 * 1024;

struct SortGatherConfig {
  int FG, groups;
  size_t smem;
};

inline bool sgather_config(const ctb_shape* s, SortGatherConfig* out) {
  if (!plan_supported(s) || !plan_dense(s)) return false;
  if (s->grid_dtype != CTB_DTYPE_F32) return false;
  if (s->N % 4 != 0 || shape_cells(s) % 4 != 0) return false;        // rows and planes move as 16-byte multiples
  static const int env_fg = getenv("CTB_SORT_FG") ? atoi(getenv("CTB_SORT_FG")) : 0;
  const size_t per = ((size_t)shape_cells(s) + (size_t)((s->N + 7) & ~7)) * 4;     // tile plane + staged row
  int FG = 1;
  while (FG * 2 <= s->F && (size_t)(FG * 2) * per + 64 <= kSortSmemTwoCtas) FG *= 2;
  if (env_fg) FG = env_fg;
  if (FG < 1 || (size_t)FG * per + 64 > kSortSmemTwoCtas) return false;
  out->FG = FG;
  out->groups = (s->F + FG - 1) / FG;
  out->smem = (size_t)FG * per + 64;
  return true;
}

template <int D>
__global__ void __launch_bounds__(kSortThreads, 2)
sorted_slice_fwd_kernel(const float* __restrict__ keys, const float* __restrict__ grid, const float* __restrict__ pad,
                        float* __restrict__ out, const uint16_t* __restrict__ perm, Grid<D> g, int H, int F, int N,
                        int Np, int FG, int groups) {
  constexpr int S = 1 << D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = g.C;
  uint64_t* bar = (uint64_t*)smem_raw;
  float* tile = (float*)(smem_raw + 64);            // [FG][C]
  float* rout = tile + (size_t)FG * C;              // [FG][Np]

  const int unit = blockIdx.x / groups;
  const int f0 = (blockIdx.x % groups) * FG;
  const int fg = min(FG, F - f0);
  const float* ku = keys + (size_t)unit * D * N;
  const float* pu = pad ? pad + (size_t)(unit / H) * N : nullptr;
  const uint16_t* pm = perm + (size_t)unit * Np;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    const uint32_t plane = (uint32_t)C * 4u;
    mbar_expect_tx(bar, plane * fg);
    for (int f = 0; f < fg; ++f) bulk_g2s(tile + (size_t)f * C, grid + ((size_t)unit * F + f0 + f) * C, plane, bar);
  }
  __syncthreads();
  bool waited = false;
  const unsigned fstep = (unsigned)C * 4u;
  // the slot -> point -> keys chain of a thread's NEXT point is fetched under the work on the current one (2-D only:
  // the 3-D point loop has no registers to spare at two 512-thread CTAs per SM)
  constexpr bool kAhead = true;
  int nn = 0;
  float nk[D], npd = 1.0f;
  auto prefetch = [&](int i) {
    nn = (int)__ldg(pm + i);
#pragma unroll
    for (int a2 = 0; a2 < D; ++a2) nk[a2] = __ldg(ku + (size_t)a2 * N + nn);
    if (pu) npd = __ldg(pu + nn);
  };
  if (kAhead && (int)threadIdx.x < N) prefetch(threadIdx.x);
#pragma unroll 1
  for (int i = threadIdx.x; i < N; i += kSortThreads) {
    int n;
    float pd, kv[D];
    if constexpr (kAhead) {
      n = nn;
      pd = npd;
#pragma unroll
      for (int a2 = 0; a2 < D; ++a2) kv[a2] = nk[a2];
      if (i + kSortThreads < N) prefetch(i + kSortThreads);
    } else {
      n = (int)__ldg(pm + i);
#pragma unroll
      for (int a2 = 0; a2 < D; ++a2) kv[a2] = __ldg(ku + (size_t)a2 * N + n);
      pd = pu ? __ldg(pu + n) : 1.0f;
    }
    const Pos<D> p = point_pos_from_values<D>(kv, g);
    float w[S];
    unsigned ab[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      w[s] = corner_weight<D>(p, s);
      ab[s] = smem_u32(tile) + (unsigned)(p.base + corner_offset<D>(g, s)) * 4u;
    }
    if (!waited) {                                  // the tile has been arriving under the position arithmetic
      mbar_wait(bar, 0);
      waited = true;
    }
#pragma unroll(D == 2 ? 4 : 1)
    for (int f = 0; f < fg; ++f) {
      const unsigned off = (unsigned)f * fstep;
      float acc = CTB_FMUL(lds_f32(ab[0] + off), w[0]);
#pragma unroll
      for (int s = 1; s < S; ++s) acc = fmaf(lds_f32(ab[s] + off), w[s], acc);
      if (pu) acc = CTB_FMUL(acc, pd);
      rout[(size_t)f * Np + n] = acc;
    }
  }
  if (!waited) mbar_wait(bar, 0);
  // the staged rows leave as whole rows
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int f = 0; f < fg; ++f) bulk_s2g(out + ((size_t)unit * F + f0 + f) * N, rout + (size_t)f * Np, (uint32_t)N * 4u);
    bulk_commit_and_wait_read();
  }
}

template <int D>
cudaError_t sorted_slice_fwd(const float* keys, const float* grid, const float* pad, float* out, const void* plan,
                             const ctb_shape* s, cudaStream_t stream) {
  SortGatherConfig c;
  if (!sgather_config(s, &c)) return cudaErrorNotSupported;
  const PlanView v = plan_view(plan, s);
  const Grid<D> g = make_grid<D>(s->size);
  const long long blocks = (long long)s->B * s->H * c.groups;
  if (blocks >= (1ll << 31)) return cudaErrorNotSupported;
  cudaError_t e = cudaFuncSetAttribute(sorted_slice_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
  if (e != cudaSuccess) return e;
  sorted_slice_fwd_kernel<D><<<(unsigned)blocks, kSortThreads, c.smem, stream>>>(keys, grid, pad, out, v.perm, g, s->H, s->F,
                                                                               s->N, v.Np, c.FG, c.groups);
  return cudaGetLastError();
}

}  // namespace ctb
