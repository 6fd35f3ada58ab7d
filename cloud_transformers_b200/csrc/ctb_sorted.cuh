// ctb_sorted.cuh -- cell-stationary (CTB_MODE_DETERMINISTIC) scatter: plan build + sorted-entry tile scatter.
// (The gathers of this mode are the tile gathers of ctb_tile.cuh, which are deterministic by construction.)
//
// Idea.  A (batch, head) unit owns its own grid slab, so the S*N (point, corner) "entries" of a unit are
// sorted ONCE by destination cell (plan_kernel, bitonic sort in shared memory).  After that
//   * every scatter (Splat forward = A2+A3, grad_grid of Slice backward = A5) walks the entries of a
//     slab of grid rows in cell order: the run of entries that hit one cell has a single owner thread,
//     which reduces it in ascending e = s*N + n order into a CTA-owned shared-memory tile.  No atomics,
//     no zero-fill pass over HBM, results bit-identical from run to run, arg = first maximum (the
//     torch-scatter CPU rule).  The tile leaves the SM once, with coalesced 16-byte stores.
//   * every gather (Slice forward = A4, grad_keys of Slice backward = A5/A7, Splat backward = A6/A7)
//     stages the grid slab in shared memory with coalesced 16-byte loads and resolves the 2^d random
//     corner reads per point against shared memory instead of L2 sectors.
// Grid layout stays the reference's NCHW / NCDHW; a slab of rows [x0, x1) of one channel plane is one
// contiguous byte range, which is what makes the tile moves coalesced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ctb200.h"
#include "ctb_positions.cuh"

namespace ctb {

constexpr int kPlanThreads = 1024;
constexpr int kScatterThreads = 512;
constexpr int kScatterChunk = 2048;          // entries staged per step
constexpr int kMaxSortEntries = 32768;       // S*N padded to a power of two must fit shared memory
constexpr int kSmemBudgetTwoCtas = 110 * 1024;
constexpr int kSmemBudgetMax = 220 * 1024;

inline int ceil_log2(unsigned long long v) {
  int b = 0;
  while ((1ull << b) < v) ++b;
  return b;
}

// ---- plan layout ------------------------------------------------------------------------------------
// per unit u:  ekey u32 [S*N]  (cell << ebits | s << nbits | n), ascending
//              ew   f32 [S*N]  corner weight of the entry (bit-exact lc value)
//              rstart i32 [W0 + 1]  first entry whose destination row (axis 0) is >= x
struct PlanView {
  uint32_t* ekey;
  float* ew;
  int* rstart;
  int E;        // S*N
  int nbits;    // bits of n
  int ebits;    // bits of (s, n)
  int rs;       // W0 + 1
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline bool plan_supported(const ctb_shape* s) {
  const long long S = 1ll << s->dim;
  const long long E = S * s->N;
  long long C = 1;
  for (int a = 0; a < s->dim; ++a) C *= s->size[a];
  if (E > kMaxSortEntries) return false;
  const int nbits = ceil_log2((unsigned long long)s->N);
  const int ebits = nbits + s->dim;
  const int cbits = ceil_log2((unsigned long long)C);
  return cbits + ebits <= 32;
}

inline PlanView plan_view(void* plan, const ctb_shape* s) {
  PlanView v;
  const size_t U = (size_t)s->B * s->H;
  v.E = (1 << s->dim) * s->N;
  v.nbits = ceil_log2((unsigned long long)s->N);
  v.ebits = v.nbits + s->dim;
  v.rs = s->size[0] + 1;
  unsigned char* p = (unsigned char*)plan;
  v.ekey = (uint32_t*)p;
  p += align_up(U * v.E * sizeof(uint32_t), 256);
  v.ew = (float*)p;
  p += align_up(U * v.E * sizeof(float), 256);
  v.rstart = (int*)p;
  return v;
}

inline size_t sorted_plan_bytes(const ctb_shape* s) {
  const size_t U = (size_t)s->B * s->H;
  const size_t E = ((size_t)1 << s->dim) * s->N;
  return align_up(U * E * 4, 256) * 2 + align_up(U * (s->size[0] + 1) * 4, 256);
}

// ---- plan build ---------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kPlanThreads)
plan_kernel(const float* __restrict__ keys, uint32_t* __restrict__ ekey, float* __restrict__ ew,
            int* __restrict__ rstart, Grid<D> g, int N, int P, int nbits) {
  constexpr int S = 1 << D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* kr = (float*)smem_raw;                 // [D][N] this unit's keys
  uint32_t* buf = (uint32_t*)(kr + (size_t)D * N);  // [P] sort buffer
  const int unit = blockIdx.x;
  const int E = S * N;
  const int ebits = nbits + D;
  const float* ku = keys + (size_t)unit * D * N;
  for (int i = threadIdx.x; i < D * N; i += kPlanThreads) kr[i] = __ldg(ku + i);
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += kPlanThreads) {
    const Pos<D> p = point_pos<D>(kr, n, N, g);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const uint32_t cell = (uint32_t)(p.base + corner_offset<D>(g, s));
      buf[s * N + n] = (cell << ebits) | ((uint32_t)s << nbits) | (uint32_t)n;
    }
  }
  for (int i = E + threadIdx.x; i < P; i += kPlanThreads) buf[i] = 0xFFFFFFFFu;
  __syncthreads();
  // bitonic sort, ascending; one compare-exchange per (thread, pair)
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += kPlanThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // insert a 0 bit at position log2(j)
        const int l = i | j;
        const uint32_t a = buf[i], b = buf[l];
        const bool up = (i & k) == 0;
        if ((a > b) == up) {
          buf[i] = b;
          buf[l] = a;
        }
      }
      __syncthreads();
    }
  }
  uint32_t* eo = ekey + (size_t)unit * E;
  float* wo = ew + (size_t)unit * E;
  const uint32_t nmask = (1u << nbits) - 1u;
  for (int i = threadIdx.x; i < E; i += kPlanThreads) {
    const uint32_t key = buf[i];
    const int n = (int)(key & nmask);
    const int s = (int)((key >> nbits) & (uint32_t)(S - 1));
    const Pos<D> p = point_pos<D>(kr, n, N, g);
    eo[i] = key;
    wo[i] = corner_weight<D>(p, s);
  }
  // rstart[x] = lower_bound(entries, cell >= x * stride0)
  const int rs = g.W[0] + 1;
  for (int x = threadIdx.x; x < rs; x += kPlanThreads) {
    const unsigned long long target = ((unsigned long long)x * (unsigned long long)g.stride[0]) << ebits;
    int lo = 0, hi = E;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((unsigned long long)buf[mid] < target) lo = mid + 1; else hi = mid;
    }
    rstart[(size_t)unit * rs + x] = lo;
  }
}

template <int D>
cudaError_t sorted_plan_build(const float* keys, void* plan, const ctb_shape* s, cudaStream_t stream) {
  if (!plan_supported(s)) return cudaErrorNotSupported;
  const PlanView v = plan_view(plan, s);
  const Grid<D> g = make_grid<D>(s->size);
  int P = 1;
  while (P < v.E) P <<= 1;
  const size_t smem = (size_t)D * s->N * 4 + (size_t)P * 4;
  if (smem > (size_t)kSmemBudgetMax) return cudaErrorNotSupported;
  cudaError_t e = cudaFuncSetAttribute(plan_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  plan_kernel<D><<<(unsigned)(s->B * s->H), kPlanThreads, smem, stream>>>(keys, v.ekey, v.ew, v.rstart, g, s->N, P,
                                                                         v.nbits);
  return cudaGetLastError();
}

// ---- tile scatter -------------------------------------------------------------------------------------
// One CTA = (unit, group of FG channels).  Walks the slabs of R grid rows; per slab the entries
// [rstart[x0], rstart[x1]) are staged in chunks and every run of equal cell is reduced by one thread.
template <int FG, bool SUM, bool VEC4>
__global__ void __launch_bounds__(kScatterThreads)
scatter_kernel(const uint32_t* __restrict__ ekey, const float* __restrict__ ew, const int* __restrict__ rstart,
               const float* __restrict__ feat, const float* __restrict__ pad, float* __restrict__ z,
               int* __restrict__ arg, int H, int F, int N, int C, int stride0, int W0, int R, int nbits, int dim,
               int groups) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tile_cells = R * stride0;
  float* rows = (float*)smem_raw;                          // [FG][N]   features * pad
  float* tval = rows + (((size_t)FG * N + 3) & ~(size_t)3);  // [FG][tile_cells], 16-byte aligned
  int* targ = (int*)(tval + (size_t)FG * tile_cells);      // [FG][tile_cells]  (MAX only)
  uint32_t* ck = (uint32_t*)(targ + (SUM ? (size_t)0 : (size_t)FG * tile_cells));  // [kScatterChunk]
  float* cw = (float*)(ck + kScatterChunk);

  const int unit = blockIdx.x / groups;
  const int f0 = (blockIdx.x % groups) * FG;
  const int E = (1 << dim) * N;
  const int ebits = nbits + dim;
  const uint32_t nmask = (1u << nbits) - 1u;
  const uint32_t smask = (1u << dim) - 1u;

  {  // stage the feature rows (pre-multiplied by the padding mask, cloud_transform.py:158-159)
    const float* fu = feat + ((size_t)unit * F + f0) * N;
    const float* pu = pad ? pad + (size_t)(unit / H) * N : nullptr;
    for (int i = threadIdx.x; i < FG * N; i += kScatterThreads) {
      float v = __ldg(fu + i);
      if (pu) v = CTB_FMUL(v, __ldg(pu + (i % N)));
      rows[i] = v;
    }
    for (int i = threadIdx.x; i < FG * tile_cells; i += kScatterThreads) {
      tval[i] = 0.0f;
      if constexpr (!SUM) targ[i] = -1;
    }
  }
  const uint32_t* eu = ekey + (size_t)unit * E;
  const float* wu = ew + (size_t)unit * E;
  const int* ru = rstart + (size_t)unit * (W0 + 1);
  __syncthreads();

  for (int x0 = 0; x0 < W0; x0 += R) {
    const int x1 = min(x0 + R, W0);
    const int j0 = __ldg(ru + x0), j1 = __ldg(ru + x1);
    const int cell0 = x0 * stride0;
    const int ncell = (x1 - x0) * stride0;
    for (int cb = j0; cb < j1; cb += kScatterChunk) {
      const int cn = min(kScatterChunk, j1 - cb);
      for (int i = threadIdx.x; i < cn; i += kScatterThreads) {
        ck[i] = __ldg(eu + cb + i);
        cw[i] = __ldg(wu + cb + i);
      }
      __syncthreads();
      for (int i = threadIdx.x; i < cn; i += kScatterThreads) {
        const uint32_t key = ck[i];
        const uint32_t cell = key >> ebits;
        const bool head = (i == 0) || ((ck[i - 1] >> ebits) != cell);
        if (!head) continue;
        const int lc = (int)cell - cell0;
        float best[FG];
        int barg[FG];
#pragma unroll
        for (int f = 0; f < FG; ++f) {
          best[f] = tval[f * tile_cells + lc];
          if constexpr (!SUM) barg[f] = targ[f * tile_cells + lc];
        }
        int k = i;
        uint32_t kk = key;
        do {
          const int n = (int)(kk & nmask);
          const int e = (int)((kk >> nbits) & smask) * N + n;
          const float w = cw[k];
#pragma unroll
          for (int f = 0; f < FG; ++f) {
            const float v = CTB_FMUL(rows[f * N + n], w);
            if constexpr (SUM) {
              best[f] = CTB_FADD(best[f], v);
            } else {
              if (v > best[f]) {
                best[f] = v;
                barg[f] = e;
              }
            }
          }
          ++k;
          if (k >= cn) break;
          kk = ck[k];
        } while ((kk >> ebits) == cell);
#pragma unroll
        for (int f = 0; f < FG; ++f) {
          tval[f * tile_cells + lc] = best[f];
          if constexpr (!SUM) targ[f * tile_cells + lc] = barg[f];
        }
      }
      __syncthreads();
    }
    // store the finished slab and clear the tile for the next one
    if constexpr (VEC4) {
      const int n4 = ncell >> 2;
      for (int i = threadIdx.x; i < FG * n4; i += kScatterThreads) {
        const int f = i / n4, r = i - f * n4;
        float4* ts = reinterpret_cast<float4*>(tval + (size_t)f * tile_cells) + r;
        float4* zo = reinterpret_cast<float4*>(z + ((size_t)unit * F + f0 + f) * C + cell0) + r;
        __stcs(zo, *ts);
        *ts = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (!SUM) {
          int4* as = reinterpret_cast<int4*>(targ + (size_t)f * tile_cells) + r;
          int4* ao = reinterpret_cast<int4*>(arg + ((size_t)unit * F + f0 + f) * C + cell0) + r;
          __stcs(ao, *as);
          *as = make_int4(-1, -1, -1, -1);
        }
      }
    } else {
      for (int i = threadIdx.x; i < FG * ncell; i += kScatterThreads) {
        const int f = i / ncell, r = i - f * ncell;
        z[((size_t)unit * F + f0 + f) * C + cell0 + r] = tval[f * tile_cells + r];
        tval[f * tile_cells + r] = 0.0f;
        if constexpr (!SUM) {
          arg[((size_t)unit * F + f0 + f) * C + cell0 + r] = targ[f * tile_cells + r];
          targ[f * tile_cells + r] = -1;
        }
      }
    }
    __syncthreads();
  }
}

struct ScatterConfig {
  int FG, R;
  size_t smem;
};

inline bool scatter_config(const ctb_shape* s, bool sum, ScatterConfig* out) {
  const int stride0 = s->dim == 2 ? s->size[1] : s->size[1] * s->size[2];
  const int W0 = s->size[0];
  const size_t per_cell = sum ? 4 : 8;
  static const int fgs[4] = {8, 4, 2, 1};
  for (int pass = 0; pass < 2; ++pass) {
    const size_t budget = pass == 0 ? kSmemBudgetTwoCtas : kSmemBudgetMax;
    for (int i = 0; i < 4; ++i) {
      const int FG = fgs[i];
      if (s->F % FG) continue;
      const size_t fixed = (((size_t)FG * s->N + 3) & ~(size_t)3) * 4 + (size_t)kScatterChunk * 8;
      if (fixed + (size_t)FG * stride0 * per_cell > budget) continue;
      int R = (int)((budget - fixed) / ((size_t)FG * stride0 * per_cell));
      if (R > W0) R = W0;
      // prefer at least 4 rows per slab (fewer barriers) unless the whole grid fits anyway
      if (pass == 0 && R < 4 && R < W0 && FG > 1) continue;
      out->FG = FG;
      out->R = R;
      out->smem = fixed + (size_t)FG * R * stride0 * per_cell;
      return true;
    }
  }
  return false;
}

template <int FG, bool SUM, bool VEC4>
cudaError_t launch_scatter(const PlanView& v, const float* feat, const float* pad, float* z, int* arg,
                           const ctb_shape* s, const ScatterConfig& c, int C, int stride0, cudaStream_t stream) {
  cudaError_t e =
      cudaFuncSetAttribute(scatter_kernel<FG, SUM, VEC4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
  if (e != cudaSuccess) return e;
  const int groups = s->F / FG;
  scatter_kernel<FG, SUM, VEC4><<<(unsigned)(s->B * s->H * groups), kScatterThreads, c.smem, stream>>>(
      v.ekey, v.ew, v.rstart, feat, pad, z, arg, s->H, s->F, s->N, C, stride0, s->size[0], c.R, v.nbits, s->dim,
      groups);
  return cudaGetLastError();
}

template <int D>
cudaError_t sorted_scatter(const void* plan, const float* feat, const float* pad, float* z, int* arg,
                           const ctb_shape* s, bool sum, cudaStream_t stream) {
  if (!plan_supported(s)) return cudaErrorNotSupported;
  ScatterConfig c;
  if (!scatter_config(s, sum, &c)) return cudaErrorNotSupported;
  const PlanView v = plan_view(const_cast<void*>(plan), s);
  const Grid<D> g = make_grid<D>(s->size);
  const int stride0 = g.stride[0];
  const bool vec4 = (g.C % 4 == 0) && (stride0 % 4 == 0) && ((reinterpret_cast<uintptr_t>(z) & 15) == 0) &&
                    (sum || (reinterpret_cast<uintptr_t>(arg) & 15) == 0);
#define CTB_SC(FGV)                                                                                          \
  case FGV:                                                                                                  \
    if (sum)                                                                                                 \
      return vec4 ? launch_scatter<FGV, true, true>(v, feat, pad, z, arg, s, c, g.C, stride0, stream)        \
                  : launch_scatter<FGV, true, false>(v, feat, pad, z, arg, s, c, g.C, stride0, stream);      \
    else                                                                                                     \
      return vec4 ? launch_scatter<FGV, false, true>(v, feat, pad, z, arg, s, c, g.C, stride0, stream)       \
                  : launch_scatter<FGV, false, false>(v, feat, pad, z, arg, s, c, g.C, stride0, stream);
  switch (c.FG) {
    CTB_SC(8)
    CTB_SC(4)
    CTB_SC(2)
    CTB_SC(1)
  }
#undef CTB_SC
  return cudaErrorNotSupported;
}

}  // namespace ctb
