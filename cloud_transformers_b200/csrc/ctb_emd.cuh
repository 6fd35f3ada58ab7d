// ctb_emd.cuh -- approximate Earth Mover's Distance by the auction algorithm (the completion
// loss of train_inpainter.py:187-192 / train_image_reconstruction.py:173-175).
//
// Reference: emd_linear/emd_cuda.cu.  Per iteration it launches seven kernels (:249-262): count / scan / list the
// unassigned sources (:30-93), Bid (:95-173: every unassigned source finds the target with the best and second-best
// value  3 - |p1 - p2| - price, bids on the best with increment best - better + eps), GetMax (:175-188: the highest
// increment per target wins), Assign (:190-210: the winner takes the target and evicts its owner, price += increment;
// in the LAST iteration every remaining source takes the target it bid on), then CalcDist (:212-221).  With the
// training setting (eps 0.005, 50 iterations) that is 351 launches per loss, 21 001 with the validation setting
// (0.004, 3000) -- almost all of them over a handful of unassigned points.
//
// Here a cluster of 1 .. 8 CTAs runs the whole auction of one cloud in ONE launch: prices, assignments, bids and the
// unassigned lists live in shared memory (28 bytes per point; clouds above 4096 points keep them in an L2-resident
// workspace instead, up to 32768 points -- the inpainting decoder emits 16384), iterations are separated by
// cluster barriers / __syncthreads(), and the loop ends as soon as nothing is unassigned.  An unassigned source is scanned by T = 1 .. 32 lanes (as many
// as 1024 threads allow), targets arrive in shared-memory tiles, and the per-target winner is ONE 64-bit shared
// atomicMax on (increment bits << 32 | ~source): the highest increment wins, exact ties go to the smallest source
// index -- deterministic, where the reference lets any bidder within 1e-6 of the maximum win by a race (:181-185).
// The value is computed as the reference's expression evaluates: the literal 3.0 makes it a double subtraction
// rounded once to float (:131).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ctb {

constexpr int kEmdThreads = 1024;
constexpr int kEmdTile = 4096;          // targets per shared-memory tile (48 KB); clouds up to here keep all targets resident
constexpr int kEmdSharedPoints = 4096;  // up to here the whole auction state lives in the shared memory of CTA 0
constexpr int kEmdMaxPoints = 32768;    // beyond, the state lives in a global workspace (L2), prices cached per CTA

__host__ __device__ inline size_t emd_state_bytes(int n) {                     // auction state of one cloud
  const size_t np = (size_t)((n + 3) & ~3);
  return np * (8 + 4 + 4 + 4 + 2 * 4) + 16 + (n > kEmdSharedPoints ? np * 8 : 0);      // + the price changes of an iteration
}
inline size_t emd_smem_bytes(int n) {
  const size_t np = (size_t)((n + 3) & ~3);
  return (n <= kEmdSharedPoints ? emd_state_bytes(n) : np * 4 + 16) + (size_t)(n < kEmdTile ? np : kEmdTile) * 3 * 4;
}
__host__ __device__ inline size_t emd_header_bytes(int B) { return ((size_t)B * 8 + 255) / 256 * 256; }   // barriers
inline size_t emd_workspace_bytes(int B, int n) {
  return n <= kEmdSharedPoints ? 0 : emd_header_bytes(B) + (size_t)B * ((emd_state_bytes(n) + 255) / 256 * 256);
}

struct EmdTop2 {
  float best, better;
  int i;
};
// merge two partial scans; the result is the one of a single scan in ascending k (first maximum wins, :132-139)
__device__ __forceinline__ void emd_merge(EmdTop2& a, float bbest, float bbetter, int bi) {
  if (bbest > a.best || (bbest == a.best && bi >= 0 && (a.i < 0 || bi < a.i))) {
    a.better = fmaxf(a.best, bbetter);
    a.best = bbest;
    a.i = bi;
  } else {
    a.better = fmaxf(a.better, bbest);
  }
}

// A cluster of CS CTAs works on one cloud while many sources are unassigned (the bids are O(unassigned x n)): the
// auction state lives with CTA 0 (its shared memory, or the global workspace for large clouds); per iteration every
// CTA refreshes its copy of the prices, bids for its share of the unassigned list and posts (target, increment) per
// source to CTA 0 (distributed shared memory / global stores), CTA 0 picks the winners, assigns and writes the NEXT
// unassigned list on the way (losers stay, evicted owners join: no O(n) rescan).  Two cluster barriers per iteration.
// The number of unassigned sources never grows (a winner evicts at most one owner), so once it is <= kEmdSolo the
// other CTAs leave and CTA 0 finishes alone with __syncthreads() only -- the long tail of the 3000-iteration
// validation setting; with few sources left, up to all 1024 lanes scan one source.
constexpr int kEmdSolo = 64;

// GLOBAL: the state arrays live in `workspace` (global memory, served by L2) instead of CTA 0's shared memory.
// What another CTA (or an atomic) wrote is read with ld.global.cg.
template <bool G, typename T>
__device__ __forceinline__ T emd_xload(const T* p) {
  if constexpr (G) return __ldcg(p);
  else return *p;
}

// One (source, target) pair.  The value as the reference's expression evaluates (:131: the literal 3.0 makes it a
// double subtraction rounded once to float), behind a cheap NECESSARY condition for "value > better": with
// c2 = 3 - better + margin, the pair can only matter if sqrt(sq) < c2 - price, i.e. sq <= (c2 - price)^2.  The margin
// (1e-5 absolute and relative) is far above the float roundings of the test, so no candidate is ever dropped and the
// result is bit-identical to evaluating every pair.
struct EmdScan {
  EmdTop2 t;
  float c2;
  __device__ __forceinline__ void reset() {
    t.best = -1e9f;
    t.better = -1e9f;
    t.i = -1;
    refresh();
  }
  __device__ __forceinline__ void refresh() {
    const float c = 3.0f - t.better;
    c2 = c + 1e-5f * (fabsf(c) + 1.0f);
  }
  __device__ __forceinline__ void consider(float dx, float dy, float dz, float p, int k) {
    const float sq = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    const float r = fmaf(-p, 0.99999f, c2);                      // (prices are >= 0)
    if (r > 0.0f && sq <= r * r) {
      const float d = (float)(3.0 - (double)__fsqrt_rn(sq) - (double)p);
      if (d > t.best) {
        t.better = t.best;
        t.best = d;
        t.i = k;
        refresh();
      } else if (d > t.better) {
        t.better = d;
        refresh();
      }
    }
  }
};

template <bool GLOBAL>
__global__ void __launch_bounds__(kEmdThreads, 1)
emd_auction_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, float* __restrict__ dist,
                   int* __restrict__ assignment, unsigned char* __restrict__ workspace, int n, float eps, int iters, int K) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  // the CTAs of a cloud: a thread-block cluster (shared state), or K co-resident CTAs of a cooperative launch that
  // meet at a barrier in the workspace (global state: small batches of large clouds then fill the GPU)
  const unsigned rank = GLOBAL ? blockIdx.x % (unsigned)K : cluster.block_rank();
  const unsigned CS = GLOBAL ? (unsigned)K : cluster.num_blocks();
  const int b = blockIdx.x / CS;
  unsigned* gbar = (unsigned*)workspace + 2 * (size_t)b;       // global state: {arrivals, generation} of this cloud
  int Kc = (int)CS;                                              // CTAs still working on this cloud
  auto cloud_sync = [&]() {
    if constexpr (GLOBAL) {
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned g;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(gbar + 1) : "memory");
        __threadfence();
        if (atomicAdd(gbar, 1u) == (unsigned)Kc - 1u) {
          atomicExch(gbar, 0u);
          __threadfence();
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(gbar + 1), "r"(g + 1u) : "memory");
        } else {
          unsigned now;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(gbar + 1) : "memory");
          } while (now == g);
        }
        __threadfence();
      }
      __syncthreads();
    } else {
      cluster.sync();
    }
  };
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int np = (n + 3) & ~3;
  unsigned char* state = GLOBAL ? workspace + emd_header_bytes(gridDim.x / CS) + (size_t)b * ((emd_state_bytes(n) + 255) / 256 * 256)
                                : smem_raw;
  unsigned long long* top = (unsigned long long*)state;         // [n] highest bid on a target: inc bits << 32 | ~source
  float* gprice = (float*)(top + np);                           // [n] the prices (CTA 0 updates them)
  float* binc = gprice + np;                                    // [n] bid increment of a source
  int* bid = (int*)(binc + np);                                 // [n] target a source bids on (written by every CTA of
                                                                //     the cluster: whole words, no sub-word remote stores)
  uint16_t* asg = (uint16_t*)(bid + np);                        // [n] source -> target, 0xffff = unassigned
  uint16_t* inv = asg + np;                                     // [n] target -> source
  uint16_t* list0 = inv + np;                                   // [n] unassigned sources, two lists taking turns
  uint16_t* list1 = list0 + np;
  int* cnts = (int*)(list1 + np);                               // [2] their lengths, [2] number of price changes
  int* chg_t = cnts + 4;                                        // [n] global state only: targets whose price changed in
  float* chg_p = (float*)(chg_t + np);                          // [n] this iteration and their new prices
  unsigned char* rest = GLOBAL ? smem_raw : (unsigned char*)(cnts + 4);
  // the prices the bid loop reads: shared memory.  Shared state: CTA 0 reads its own array, the others a copy of it;
  // global state: every CTA a copy (CTA 0 keeps its copy current while it assigns).
  float* price = GLOBAL ? (float*)rest : gprice;
  float* tile = GLOBAL ? price + np : (float*)rest;             // [min(n, kEmdTile)][3] targets
  __shared__ float mrg[32 * 3];                                 // partial scans of the warps of a source
  const int tile_pts = min(n, kEmdTile);
  const bool resident = n <= kEmdTile;                          // all targets stay in the tile
  // CTA 0's arrays as the other CTAs of the cluster see them
  const float* m_price = GLOBAL ? gprice : cluster.map_shared_rank(gprice, 0);
  float* m_binc = GLOBAL ? binc : cluster.map_shared_rank(binc, 0);
  int* m_bid = GLOBAL ? bid : cluster.map_shared_rank(bid, 0);
  const uint16_t* m_list0 = GLOBAL ? list0 : cluster.map_shared_rank(list0, 0);
  const uint16_t* m_list1 = GLOBAL ? list1 : cluster.map_shared_rank(list1, 0);
  const int* m_cnts = GLOBAL ? cnts : cluster.map_shared_rank(cnts, 0);

  const float* x1 = xyz1 + (size_t)b * n * 3;
  const float* x2 = xyz2 + (size_t)b * n * 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (rank == 0) {
    for (int j = threadIdx.x; j < n; j += kEmdThreads) {
      top[j] = 0ull;
      gprice[j] = 0.0f;
      if (GLOBAL) price[j] = 0.0f;
      asg[j] = 0xffffu;
      inv[j] = 0xffffu;
      list0[j] = (uint16_t)j;                                   // everybody starts unassigned (:30-93)
    }
    if (threadIdx.x == 0) {
      cnts[0] = n;
      cnts[1] = 0;
      cnts[2] = 0;
    }
    __syncthreads();
  }
  if (GLOBAL && rank != 0)
    for (int j = threadIdx.x; j < n; j += kEmdThreads) price[j] = 0.0f;
  if (resident)
    for (int i = threadIdx.x; i < n * 3; i += kEmdThreads) tile[i] = __ldg(x2 + i);
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    const bool last = it == iters - 1;
    const int cur = it & 1;
    const bool was_together = Kc > 1;
    if (was_together) cloud_sync();
    const int U = emd_xload<GLOBAL>((was_together ? m_cnts : cnts) + cur);
    if (U == 0) {
      if (!GLOBAL && was_together) cluster.sync();       // CTA 0 must not leave before everyone has read its count
      break;
    }
    // The number of unassigned sources never grows, so CTAs only ever leave.  A cluster stays whole until CTA 0 can
    // finish alone; the co-resident CTAs of the global variant thin out so that everyone left has > 4 sources.
    if (was_together) {
      int want = Kc;
      if (GLOBAL) {
        while (want > 1 && U <= (want / 2) * 8) want /= 2;
      } else if (U <= kEmdSolo) {
        want = 1;
      }
      if (GLOBAL && rank != 0 && (int)rank < want) {
        // the prices CTA 0 changed in the previous iteration
        const int nc = __ldcg(cnts + 2);
        for (int i = threadIdx.x; i < nc; i += kEmdThreads) price[__ldcg(chg_t + i)] = __ldcg(chg_p + i);
        __syncthreads();
      }
      if (!GLOBAL && want != Kc) cluster.sync();         // (everyone has read CTA 0's count before anyone moves on)
      Kc = want;
      if ((int)rank >= Kc) return;       // (nothing reads this CTA's shared memory)
    }
    const bool together = Kc > 1;
    const uint16_t* m_una = together ? (cur ? m_list1 : m_list0) : (cur ? list1 : list0);
    int* w_bid = together ? m_bid : bid;                     // (alone, CTA 0 uses its own addresses)
    float* w_binc = together ? m_binc : binc;
    // ---- bids (:95-173) ----------------------------------------------------------------------------------------
    // T lanes scan one source (T = 1 .. 1024: as many as the CTA's share of the list allows); targets come from the
    // shared-memory tile -- resident for the whole kernel where the cloud fits it, else refilled per tile.
    const int per = (U + Kc - 1) / Kc;
    const int s_lo = min(U, (int)rank * per), s_hi = min(U, s_lo + per);
    const int mine = s_hi - s_lo;
    if (!GLOBAL && together && rank != 0) {
      for (int j = threadIdx.x; j < n; j += kEmdThreads) price[j] = m_price[j];
      __syncthreads();
    }
    int T = 1;
    while (T < kEmdThreads && max(mine, 1) * (T * 2) <= kEmdThreads) T *= 2;
    const int per_pass = kEmdThreads / T;            // sources per pass over the targets
    const int sub = threadIdx.x % T;                 // my position among the lanes of my source
    for (int s0 = s_lo; s0 < s_hi; s0 += per_pass) {
      const int sidx = s0 + threadIdx.x / T;
      const bool active = sidx < s_hi;
      const int j = active ? (int)emd_xload<GLOBAL>(m_una + sidx) : 0;
      const float px = __ldg(x1 + (size_t)j * 3), py = __ldg(x1 + (size_t)j * 3 + 1), pz = __ldg(x1 + (size_t)j * 3 + 2);
      EmdScan sc;
      sc.reset();
      for (int k0 = 0; k0 < n; k0 += tile_pts) {
        const int ck = min(tile_pts, n - k0);
        if (!resident) {
          __syncthreads();
          for (int i = threadIdx.x; i < ck * 3; i += kEmdThreads) tile[i] = __ldg(x2 + (size_t)k0 * 3 + i);
          __syncthreads();
        }
        if (active) {
#pragma unroll 2
          for (int k = sub; k < ck; k += T)
            sc.consider(__fsub_rn(tile[k * 3], px), __fsub_rn(tile[k * 3 + 1], py), __fsub_rn(tile[k * 3 + 2], pz),
                        price[k0 + k], k0 + k);
        }
      }
      // merge the T partial scans: inside a warp by shuffles, then (T > 32) across the warps of the source
      for (int o = min(T, 32) >> 1; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, sc.t.best, o), obb = __shfl_xor_sync(0xffffffffu, sc.t.better, o);
        const int oi = __shfl_xor_sync(0xffffffffu, sc.t.i, o);
        emd_merge(sc.t, ob, obb, oi);
      }
      if (T > 32) {
        const int W = T >> 5;                          // warps per source (aligned groups of W warps)
        if (lane == 0) {
          mrg[warp * 3] = sc.t.best;
          mrg[warp * 3 + 1] = sc.t.better;
          mrg[warp * 3 + 2] = __int_as_float(sc.t.i);
        }
        __syncthreads();
        if (lane == 0 && (warp % W) == 0)
          for (int w = 1; w < W; ++w)
            emd_merge(sc.t, mrg[(warp + w) * 3], mrg[(warp + w) * 3 + 1], __float_as_int(mrg[(warp + w) * 3 + 2]));
        __syncthreads();
      }
      if (active && sub == 0) {
        w_bid[j] = sc.t.i;
        w_binc[j] = __fadd_rn(__fsub_rn(sc.t.best, sc.t.better), eps);
      }
    }
    if (together) cloud_sync(); else __syncthreads();
    // ---- winners take their targets (:175-210) -----------------------------------------------------------------
    if (rank == 0) {
      const uint16_t* una = cur ? list1 : list0;
      uint16_t* nxt = cur ? list0 : list1;
      int* ncnt = cnts + (cur ^ 1);
      // the highest increment per target, ties to the smallest source (CTA 0's own atomics: the other CTAs only post
      // (target, increment) per source)
      if (!last) {
        for (int s = threadIdx.x; s < U; s += kEmdThreads) {
          const int j = (int)una[s];
          atomicMax(top + emd_xload<GLOBAL>(bid + j),
                    ((unsigned long long)__float_as_uint(emd_xload<GLOBAL>(binc + j)) << 32) | (unsigned long long)(0xffffffffu - (unsigned)j));
        }
        if (threadIdx.x == 0) {
          *ncnt = 0;
          cnts[2] = 0;
        }
        __syncthreads();
      }
      for (int s0 = 0; s0 < U; s0 += kEmdThreads) {
        const int s = s0 + threadIdx.x;
        int push = -1;                                           // joins the next unassigned list
        if (s < U) {
          const int j = (int)una[s];
          const int tg = emd_xload<GLOBAL>(bid + j);
          if (last) {
            asg[j] = (uint16_t)tg;                             // (:196-199: everybody left takes what it bid on)
          } else if ((int)(0xffffffffu - (unsigned)(emd_xload<GLOBAL>(top + tg) & 0xffffffffull)) == j) {
            const unsigned prev = inv[tg];
            if (prev != 0xffffu) {
              asg[prev] = 0xffffu;
              push = (int)prev;
            }
            inv[tg] = (uint16_t)j;
            asg[j] = (uint16_t)tg;
            const float np_ = __fadd_rn(gprice[tg], emd_xload<GLOBAL>(binc + j));
            gprice[tg] = np_;
            if (GLOBAL) {
              price[tg] = np_;
              if (together) {
                const int ci = atomicAdd(cnts + 2, 1);
                chg_t[ci] = tg;
                chg_p[ci] = np_;
              }
            }
          } else {
            push = j;
          }
        }
        if (!last) {
          const unsigned m = __ballot_sync(0xffffffffu, push >= 0);
          int base = 0;
          if (lane == 0 && m) base = atomicAdd(ncnt, __popc(m));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (push >= 0) nxt[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)push;
        }
      }
      __syncthreads();
      // the targets that were bid on are open for bids again (after every loser has read the winning bid)
      if (!last) {
        for (int s = threadIdx.x; s < U; s += kEmdThreads) {
          const int j = (int)una[s];
          const int tg = emd_xload<GLOBAL>(bid + j);
          if (asg[j] == (uint16_t)tg && inv[tg] == (uint16_t)j) top[tg] = 0ull;
        }
        __syncthreads();
      }
    }
  }
  if (rank != 0) return;
  __syncthreads();
  // ---- squared distance to the assigned target (:212-221) --------------------------------------------------------
  for (int j = threadIdx.x; j < n; j += kEmdThreads) {
    const int k = asg[j] == 0xffffu ? -1 : (int)asg[j];
    float d = 0.0f;
    if (k >= 0) {
      const float dx = __fsub_rn(__ldg(x1 + (size_t)j * 3), __ldg(x2 + (size_t)k * 3));
      const float dy = __fsub_rn(__ldg(x1 + (size_t)j * 3 + 1), __ldg(x2 + (size_t)k * 3 + 1));
      const float dz = __fsub_rn(__ldg(x1 + (size_t)j * 3 + 2), __ldg(x2 + (size_t)k * 3 + 2));
      d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    }
    dist[(size_t)b * n + j] = d;
    assignment[(size_t)b * n + j] = k;
  }
}

// grad_xyz1[j] = 2 grad_dist[j] (p1[j] - p2[assignment[j]]); the reference gives xyz2 no gradient (emd_module.py:73-80)
__global__ void emd_grad_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                const float* __restrict__ grad_dist, const int* __restrict__ assignment,
                                float* __restrict__ grad_xyz1, long long total, int n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / n;
    const int k = assignment[i];
    const float g = __fmul_rn(grad_dist[i], 2.0f);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float p2 = k >= 0 ? xyz2[((size_t)b * n + k) * 3 + a] : xyz1[(size_t)i * 3 + a];
      grad_xyz1[(size_t)i * 3 + a] = __fmul_rn(g, __fsub_rn(xyz1[(size_t)i * 3 + a], p2));
    }
  }
}

inline cudaError_t emd_forward(const float* xyz1, const float* xyz2, float* dist, int* assignment, void* workspace, int B, int n,
                               float eps, int iters, cudaStream_t stream) {
  const size_t smem = emd_smem_bytes(n);
  const bool global = n > kEmdSharedPoints;
  unsigned char* ws = (unsigned char*)workspace;
  static const int env_cs = getenv("CTB_EMD_CLUSTER") ? atoi(getenv("CTB_EMD_CLUSTER")) : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kEmdThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  cfg.attrs = at;
  if (!global) {
    cudaError_t e = cudaFuncSetAttribute(emd_auction_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // CTAs per cloud: one per 512 points (the first iterations bid for ~n sources against n targets; measured:
    // 32 x 2048 points 1.68 / 1.16 / 0.89 ms with 1 / 2 / 4 CTAs), at most 8 (the portable cluster size)
    int cs = 1;
    while (cs < 8 && cs * 2 * 512 <= n) cs *= 2;
    if (env_cs >= 1 && env_cs <= 8 && (env_cs & (env_cs - 1)) == 0) cs = env_cs;
    cfg.gridDim = dim3((unsigned)B * cs);
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, emd_auction_kernel<false>, xyz1, xyz2, dist, assignment, ws, n, eps, iters, cs);
  }
  // global state: K CTAs per cloud, one per 256 points, as many as are co-resident (they spin on a barrier)
  cudaError_t e = cudaFuncSetAttribute(emd_auction_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, emd_auction_kernel<true>, kEmdThreads, smem);
  if (e != cudaSuccess) return e;
  const int room = sms * per_sm;
  int K = 1;
  while (K < 64 && K * 2 * 256 <= n && (long long)B * (K * 2) <= room) K *= 2;
  if (env_cs >= 1 && (env_cs & (env_cs - 1)) == 0 && (long long)B * env_cs <= room) K = env_cs;
  cfg.gridDim = dim3((unsigned)B * K);
  if (K > 1) {
    e = cudaMemsetAsync(ws, 0, emd_header_bytes(B), stream);
    if (e != cudaSuccess) return e;
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.numAttrs = 1;
  } else {
    cfg.numAttrs = 0;
  }
  return cudaLaunchKernelEx(&cfg, emd_auction_kernel<true>, xyz1, xyz2, dist, assignment, ws, n, eps, iters, K);
}

inline cudaError_t emd_backward(const float* xyz1, const float* xyz2, const float* grad_dist, const int* assignment,
                                float* grad_xyz1, int B, int n, cudaStream_t stream) {
  const long long total = (long long)B * n;
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (blocks > 148u * 8u) blocks = 148u * 8u;
  emd_grad_kernel<<<blocks, 256, 0, stream>>>(xyz1, xyz2, grad_dist, assignment, grad_xyz1, total, n);
  return cudaGetLastError();
}

}  // namespace ctb
