// ctb_plan.cuh -- the per-unit plan: every (batch, head) unit's (point, corner) entries grouped by destination cell.
//
// Built ONCE per DifferentiablePositions call and shared by the operators of an MHCT block (layers/multihead_ct.py:99-107
// feeds the same local_coordinate / flattened_index to Splat and to Slice, so all passes see the same keys):
//
//   perm    u16   [U][Np]   points sorted by base cell (stable: ascending point index inside a cell); slot -> point
//   rstart  i32   [U][RS]   first slot whose base row (grid axis 0) is >= x,  x = 0 .. W0
//   cstart  u16   [U][CS]   first entry of destination cell c, c = 0 .. C                      (entry plans only)
//   ent     uint2 [U][E]    the S N entries of the unit ordered by (cell, s, n):                (entry plans only)
//                           .x = cell << ebits | s << nbits | n   (ascending over the whole list; inside a cell it is
//                                ascending e = s N + n, the order of torch-scatter's CPU loop),  ebits = nbits + dim
//                           .y = bits of the corner weight
//
// How: a stable LSD radix sort of the packed words (base cell << nbits | n) in shared memory, one CTA per unit --
// every warp owns a contiguous chunk, ranks its elements with __match_any_sync, and a (digit, warp) prefix turns the
// warp-private counts into destinations: no atomics, reproducible order.  The entries of a cell are the points of
// its 2^d neighbouring bins  bin(s) = cell - corner_offset(s)  taken in ascending s, so an entry's position follows
// from three prefix sums (cell start, earlier corners of the cell, rank inside its bin) without sorting S N items.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ctb200.h"
#include "ctb_positions.cuh"

namespace ctb {

constexpr int kPlanThreads = 512;
constexpr int kPlanWarps = kPlanThreads / 32;
constexpr int kPlanMaxPoints = 24576;     // two packed buffers of N words + histograms must fit shared memory
constexpr int kPlanMaxDigitBits = 8;
constexpr size_t kPlanSmemMax = 220 * 1024;

inline int ceil_log2(unsigned long long v) {
  int b = 0;
  while ((1ull << b) < v) ++b;
  return b;
}
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline long long shape_cells(const ctb_shape* s) {
  long long C = 1;
  for (int a = 0; a < s->dim; ++a) C *= s->size[a];
  return C;
}

struct PlanLayout {
  int N, Np, RS, CS, E;    // points, padded points, rstart / cstart entries per unit, entries per unit (0 = none)
  int nbits, cbits, passes, dbits;
  size_t perm_off, rstart_off, cstart_off, ent_off, bytes;
  size_t build_smem;
};

inline bool plan_supported(const ctb_shape* s) {
  if (s->N > kPlanMaxPoints) return false;
  const long long C = shape_cells(s);
  return ceil_log2((unsigned long long)C) + ceil_log2((unsigned long long)s->N) <= 32;
}

// entry plans: cell starts and entry counts are u16
inline bool plan_has_entries(const ctb_shape* s) {
  if (!plan_supported(s)) return false;
  const long long C = shape_cells(s), E = (long long)s->N << s->dim;
  if (C > 65535 || E > 65535) return false;
  const int cbits = ceil_log2((unsigned long long)C);
  if (cbits + ceil_log2((unsigned long long)s->N) + s->dim > 32) return false;     // cell | s | n in one word
  const int passes = (cbits + kPlanMaxDigitBits - 1) / kPlanMaxDigitBits;
  const int dbits = (cbits + passes - 1) / passes;
  const size_t Np = (size_t)((s->N + 7) & ~7);
  return Np * 8 + (size_t)(kPlanWarps + 1) * ((size_t)1 << dbits) * 4 + 2 * (size_t)((C + 2 + 7) & ~7ll) * 2 <= kPlanSmemMax;
}
// dense grid: at least four (point, corner) entries per cell on average -- where the plan pays for itself
// (measured on B200, profiles/r02_ops_*.txt: 64^2 x F16 at N = 2048, two entries per cell, loses the plan build)
inline bool plan_dense(const ctb_shape* s) { return ((long long)s->N << s->dim) >= 4 * shape_cells(s); }

inline PlanLayout plan_layout(const ctb_shape* s) {
  PlanLayout L;
  const size_t U = (size_t)s->B * s->H;
  const long long C = shape_cells(s);
  const bool ent = plan_has_entries(s);
  L.N = s->N;
  L.Np = (s->N + 7) & ~7;
  L.RS = (s->size[0] + 1 + 3) & ~3;
  L.CS = ent ? (int)((C + 2 + 7) & ~7ll) : 0;
  L.E = ent ? (((s->N << s->dim) + 1) & ~1) : 0;
  L.nbits = ceil_log2((unsigned long long)s->N);
  L.cbits = ceil_log2((unsigned long long)C);
  L.passes = (L.cbits + kPlanMaxDigitBits - 1) / kPlanMaxDigitBits;
  if (L.passes < 1) L.passes = 1;
  L.dbits = (L.cbits + L.passes - 1) / L.passes;
  if (L.dbits < 1) L.dbits = 1;
  size_t o = 0;
  L.perm_off = o;   o += align_up(U * L.Np * 2, 256);
  L.rstart_off = o; o += align_up(U * L.RS * 4, 256);
  L.cstart_off = o; o += align_up(U * (size_t)L.CS * 2, 256);
  L.ent_off = o;    o += align_up(U * (size_t)L.E * 8, 256);
  L.bytes = o;
  L.build_smem = (size_t)L.Np * 8 + (size_t)(kPlanWarps + 1) * ((size_t)1 << L.dbits) * 4 + 2 * (size_t)L.CS * 2;
  return L;
}

struct PlanView {
  const uint16_t* perm;
  const int* rstart;
  const uint16_t* cstart;   // nullptr if the plan has no entries
  const uint2* ent;
  int Np, RS, CS, E, nbits;
};

inline PlanView plan_view(const void* plan, const ctb_shape* s) {
  const PlanLayout L = plan_layout(s);
  const unsigned char* p = (const unsigned char*)plan;
  PlanView v;
  v.perm = (const uint16_t*)(p + L.perm_off);
  v.rstart = (const int*)(p + L.rstart_off);
  v.cstart = L.CS ? (const uint16_t*)(p + L.cstart_off) : nullptr;
  v.ent = L.E ? (const uint2*)(p + L.ent_off) : nullptr;
  v.Np = L.Np;
  v.RS = L.RS;
  v.CS = L.CS;
  v.E = L.E;
  v.nbits = L.nbits;
  return v;
}

// One stable counting pass on `dbits` bits at `shift`: src -> dst.  hist is u32 [kPlanWarps][R], tot u32 [R].
__device__ __forceinline__ void plan_radix_pass(const uint32_t* src, uint32_t* dst, uint32_t* hist, uint32_t* tot,
                                                int N, int shift, int dbits) {
  const int R = 1 << dbits;
  const uint32_t mask = (uint32_t)R - 1u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (((N + kPlanWarps - 1) / kPlanWarps) + 31) & ~31;     // chunk of a warp, a multiple of 32
  const int beg = min(N, warp * per), end = min(N, beg + per);
  for (int i = threadIdx.x; i < kPlanWarps * R; i += kPlanThreads) hist[i] = 0;
  __syncthreads();
  uint32_t* myh = hist + warp * R;
  // A. warp-private digit counts
  for (int b0 = beg; b0 < end; b0 += 32) {
    const int i = b0 + lane;
    const bool ok = i < end;
    const unsigned act = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const uint32_t d = (src[i] >> shift) & mask;
      const unsigned peers = __match_any_sync(act, d);
      if (lane == __ffs(peers) - 1) myh[d] += __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  // B. per digit: exclusive prefix over the warps; then an exclusive scan of the digit totals (warp 0)
  for (int d = threadIdx.x; d < R; d += kPlanThreads) {
    uint32_t run = 0;
    for (int w = 0; w < kPlanWarps; ++w) {
      const uint32_t t = hist[w * R + d];
      hist[w * R + d] = run;
      run += t;
    }
    tot[d] = run;
  }
  __syncthreads();
  if (warp == 0) {
    const int per_lane = (R + 31) / 32;
    uint32_t local = 0;
    for (int k = 0; k < per_lane; ++k) {
      const int d = lane * per_lane + k;
      if (d < R) local += tot[d];
    }
    uint32_t incl = local;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    uint32_t run = incl - local;
    for (int k = 0; k < per_lane; ++k) {
      const int d = lane * per_lane + k;
      if (d < R) {
        const uint32_t t = tot[d];
        tot[d] = run;
        run += t;
      }
    }
  }
  __syncthreads();
  // C. stable scatter: destination = digit base + earlier warps + earlier elements of this warp
  for (int b0 = beg; b0 < end; b0 += 32) {
    const int i = b0 + lane;
    const bool ok = i < end;
    const unsigned act = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const uint32_t v = src[i];
      const uint32_t d = (v >> shift) & mask;
      const unsigned peers = __match_any_sync(act, d);
      const uint32_t base = tot[d] + myh[d];
      dst[base + __popc(peers & ((1u << lane) - 1u))] = v;
      __syncwarp(act);
      if (lane == __ffs(peers) - 1) myh[d] = myh[d] + __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
}

// exclusive scan of a u16 array in shared memory (values and total < 65536), in place; all threads call it
__device__ __forceinline__ void plan_scan_u16(uint16_t* a, int n, uint32_t* warp_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (n + kPlanThreads - 1) / kPlanThreads;
  const int beg = min(n, (int)threadIdx.x * per), end = min(n, beg + per);
  uint32_t local = 0;
  for (int i = beg; i < end; ++i) local += a[i];
  uint32_t incl = local;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t t = lane < kPlanWarps ? warp_tot[lane] : 0, inc2 = t;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, inc2, o);
      if (lane >= o) inc2 += up;
    }
    if (lane < kPlanWarps) warp_tot[lane] = inc2 - t;
  }
  __syncthreads();
  uint32_t run = warp_tot[warp] + incl - local;
  for (int i = beg; i < end; ++i) {
    const uint32_t t = a[i];
    a[i] = (uint16_t)run;
    run += t;
  }
  __syncthreads();
}

template <int D>
__global__ void __launch_bounds__(kPlanThreads)
plan_build_kernel(const float* __restrict__ keys, uint16_t* __restrict__ perm, int* __restrict__ rstart,
                  uint16_t* __restrict__ cstart, uint2* __restrict__ ent, Grid<D> g, int N, int Np, int RS, int CS,
                  int E, int nbits, int passes, int dbits) {
  constexpr int S = 1 << D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* bufA = (uint32_t*)smem_raw;
  uint32_t* bufB = bufA + Np;
  uint32_t* hist = bufB + Np;                                  // [kPlanWarps][1 << dbits]
  uint32_t* tot = hist + kPlanWarps * (1 << dbits);            // [1 << dbits]
  uint16_t* bs = (uint16_t*)(tot + (1 << dbits));              // [CS] first slot of every base cell (bin)
  uint16_t* cs = bs + CS;                                      // [CS] first entry of every destination cell
  const int unit = blockIdx.x;
  const float* ku = keys + (size_t)unit * D * N;
  for (int n = threadIdx.x; n < N; n += kPlanThreads) {
    const Pos<D> p = point_pos<D>(ku, n, N, g);
    bufA[n] = ((uint32_t)p.base << nbits) | (uint32_t)n;
  }
  __syncthreads();
  uint32_t* src = bufA;
  uint32_t* dst = bufB;
  for (int pass = 0; pass < passes; ++pass) {
    plan_radix_pass(src, dst, hist, tot, N, nbits + pass * dbits, dbits);
    uint32_t* t = src;
    src = dst;
    dst = t;
  }
  const uint32_t nmask = (1u << nbits) - 1u;
  uint16_t* pu = perm + (size_t)unit * Np;
  for (int i = threadIdx.x; i < N; i += kPlanThreads) pu[i] = (uint16_t)(src[i] & nmask);
  // row starts: first slot whose base cell is >= x * stride0
  int* rs = rstart + (size_t)unit * RS;
  for (int x = threadIdx.x; x <= g.W[0]; x += kPlanThreads) {
    const uint32_t cell = (uint32_t)x * (uint32_t)g.stride[0];
    int lo = 0, hi = N;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((src[mid] >> nbits) < cell) lo = mid + 1; else hi = mid;
    }
    rs[x] = lo;
  }
  if (ent == nullptr) return;

  // bin starts: first slot whose base cell is >= c, for c = 0 .. C + 1 (binary search: bounded work per thread,
  // whatever the clustering)
  const int C = g.C;
  for (int c = threadIdx.x; c < CS; c += kPlanThreads) {
    int lo = 0, hi = N;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((int)(src[mid] >> nbits) < c) lo = mid + 1; else hi = mid;
    }
    bs[c] = (uint16_t)lo;
  }
  __syncthreads();
  // entries per destination cell: the points of the 2^d bins  cell - corner_offset(s).  A bin index that leaves the
  // grid through a lower face wraps onto a cell with coordinate W - 1 on that axis, which is never a base cell
  // (always empty), so only negative indices need a guard.
  auto bin_count = [&](int b) { return b >= 0 ? (int)bs[b + 1] - (int)bs[b] : 0; };
  for (int c = threadIdx.x; c < CS; c += kPlanThreads) {
    int t = 0;
    if (c < C) {
#pragma unroll
      for (int s = 0; s < S; ++s) t += bin_count(c - corner_offset<D>(g, s));
    }
    cs[c] = (uint16_t)t;
  }
  __syncthreads();
  plan_scan_u16(cs, C + 1, hist);            // cs[c] = first entry of cell c, cs[C] = S N
  uint16_t* cu = cstart + (size_t)unit * CS;
  for (int c = threadIdx.x; c < CS; c += kPlanThreads) cu[c] = c <= C ? cs[c] : cs[C];
  // placement: entry (slot j of bin b, corner s) -> cell c = b + offset(s), position = start of c + entries of the
  // earlier corners of c + rank of the point inside its bin
  uint2* eu = ent + (size_t)unit * E;
  for (int j = threadIdx.x; j < N; j += kPlanThreads) {
    const uint32_t pk = src[j];
    const int n = (int)(pk & nmask), b = (int)(pk >> nbits);
    const Pos<D> p = point_pos<D>(ku, n, N, g);
    const int rank_in_bin = j - (int)bs[b];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int c = b + corner_offset<D>(g, s);
      int before = 0;
#pragma unroll
      for (int s2 = 0; s2 < s; ++s2) before += bin_count(c - corner_offset<D>(g, s2));
      eu[(int)cs[c] + before + rank_in_bin] =
          make_uint2(((uint32_t)c << (nbits + D)) | ((uint32_t)s << nbits) | (uint32_t)n,
                     (uint32_t)__float_as_int(corner_weight<D>(p, s)));
    }
  }
}

template <int D>
cudaError_t plan_build(const float* keys, void* plan, const ctb_shape* s, cudaStream_t stream) {
  if (!plan_supported(s)) return cudaErrorNotSupported;
  const PlanLayout L = plan_layout(s);
  const Grid<D> g = make_grid<D>(s->size);
  cudaError_t e = cudaFuncSetAttribute(plan_build_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)L.build_smem);
  if (e != cudaSuccess) return e;
  unsigned char* p = (unsigned char*)plan;
  plan_build_kernel<D><<<(unsigned)(s->B * s->H), kPlanThreads, L.build_smem, stream>>>(
      keys, (uint16_t*)(p + L.perm_off), (int*)(p + L.rstart_off), L.CS ? (uint16_t*)(p + L.cstart_off) : nullptr,
      L.E ? (uint2*)(p + L.ent_off) : nullptr, g, s->N, L.Np, L.RS, L.CS, L.E, L.nbits, L.passes, L.dbits);
  return cudaGetLastError();
}

}  // namespace ctb
