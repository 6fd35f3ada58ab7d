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
// unassigned list live in shared memory (26 bytes per point: clouds up to 8192 points), iterations are separated by
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
constexpr int kEmdTile = 1024;          // targets per shared-memory tile (12 KB)
constexpr int kEmdMaxPoints = 8192;

inline size_t emd_smem_bytes(int n) {
  const size_t np = (size_t)((n + 3) & ~3);
  return np * (8 + 4 + 4 + 4 + 2 * 3) + (size_t)kEmdTile * 3 * 4 + 16;
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
// auction state lives in the shared memory of CTA 0; per iteration CTA 0 lists the unassigned sources, every CTA
// copies the prices, bids for its share of the list and posts (target, increment) per source into CTA 0's shared memory
// (distributed shared memory stores), CTA 0 picks the winners and assigns.  Two cluster barriers per iteration.  The number of
// unassigned sources never grows (a winner evicts at most one owner), so once it is <= kEmdSolo the other CTAs leave
// and CTA 0 finishes alone with __syncthreads() only -- the long tail of the 3000-iteration validation setting.
constexpr int kEmdSolo = 64;

__global__ void __launch_bounds__(kEmdThreads, 1)
emd_auction_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, float* __restrict__ dist,
                   int* __restrict__ assignment, int n, float eps, int iters) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank(), CS = cluster.num_blocks();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int np = (n + 3) & ~3;
  unsigned long long* top = (unsigned long long*)smem_raw;      // [n] highest bid on a target: inc bits << 32 | ~source
  float* price = (float*)(top + np);                            // [n]  (CTA 0: the prices; others: their copy)
  float* binc = price + np;                                     // [n] bid increment of a source
  int* bid = (int*)(binc + np);                                 // [n] target a source bids on (written by every CTA of
                                                                //     the cluster: whole words, no sub-word remote stores)
  uint16_t* asg = (uint16_t*)(bid + np);                        // [n] source -> target, 0xffff = unassigned
  uint16_t* inv = asg + np;                                     // [n] target -> source
  uint16_t* una = inv + np;                                     // [n] list of unassigned sources
  float* tile = (float*)(una + np);                             // [kEmdTile][3]
  int* cnt = (int*)(tile + kEmdTile * 3);
  // the master copies (CTA 0 of the cluster); for CTA 0 these are its own arrays
  const float* m_price = cluster.map_shared_rank(price, 0);
  float* m_binc = cluster.map_shared_rank(binc, 0);
  int* m_bid = cluster.map_shared_rank(bid, 0);
  const uint16_t* m_una = cluster.map_shared_rank(una, 0);
  const int* m_cnt = cluster.map_shared_rank(cnt, 0);

  const int b = blockIdx.x / CS;
  const float* x1 = xyz1 + (size_t)b * n * 3;
  const float* x2 = xyz2 + (size_t)b * n * 3;
  const int lane = threadIdx.x & 31;

  if (rank == 0) {
    for (int j = threadIdx.x; j < n; j += kEmdThreads) {
      top[j] = 0ull;
      price[j] = 0.0f;
      asg[j] = 0xffffu;
      inv[j] = 0xffffu;
    }
  }
  bool together = CS > 1;                // the cluster still works as one
  for (int it = 0; it < iters; ++it) {
    const bool last = it == iters - 1;
    // ---- list the unassigned sources (:30-93) ------------------------------------------------------------------
    if (rank == 0) {
      if (threadIdx.x == 0) *cnt = 0;
      __syncthreads();
      for (int j0 = 0; j0 < n; j0 += kEmdThreads) {
        const int j = j0 + threadIdx.x;
        const bool un = j < n && asg[j] == 0xffffu;
        const unsigned m = __ballot_sync(0xffffffffu, un);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(cnt, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (un) una[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)j;
      }
      __syncthreads();
    }
    if (together) cluster.sync();
    const int U = together ? *m_cnt : *cnt;
    if (U == 0) break;
    if (together && U <= kEmdSolo) {
      together = false;
      if (rank != 0) return;             // (nothing reads this CTA's shared memory; CTA 0 goes on alone)
    }
    // ---- bids (:95-173) ----------------------------------------------------------------------------------------
    const int parts = together ? (int)CS : 1;
    const int per = (U + parts - 1) / parts;
    const int s_lo = min(U, (together ? (int)rank : 0) * per), s_hi = min(U, s_lo + per);
    if (together && rank != 0)
      for (int j = threadIdx.x; j < n; j += kEmdThreads) price[j] = m_price[j];
    int T = 1;
    while (T < 32 && max(s_hi - s_lo, 1) * (T * 2) <= kEmdThreads) T *= 2;         // lanes per source
    const int per_pass = kEmdThreads / T;
    const int sub = threadIdx.x % T;
    for (int s0 = s_lo; s0 < s_hi; s0 += per_pass) {
      const int sidx = s0 + threadIdx.x / T;
      const bool active = sidx < s_hi;
      const int j = active ? (int)m_una[sidx] : 0;
      const float px = __ldg(x1 + (size_t)j * 3), py = __ldg(x1 + (size_t)j * 3 + 1), pz = __ldg(x1 + (size_t)j * 3 + 2);
      EmdTop2 t;
      t.best = -1e9f;
      t.better = -1e9f;
      t.i = -1;
      for (int k0 = 0; k0 < n; k0 += kEmdTile) {
        const int ck = min(kEmdTile, n - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < ck * 3; i += kEmdThreads) tile[i] = __ldg(x2 + (size_t)k0 * 3 + i);
        __syncthreads();
        if (active) {
#pragma unroll 2
          for (int k = sub; k < ck; k += T) {
            const float dx = __fsub_rn(tile[k * 3], px), dy = __fsub_rn(tile[k * 3 + 1], py), dz = __fsub_rn(tile[k * 3 + 2], pz);
            const float sq = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const float d = (float)(3.0 - (double)__fsqrt_rn(sq) - (double)price[k0 + k]);
            if (d > t.best) {
              t.better = t.best;
              t.best = d;
              t.i = k0 + k;
            } else if (d > t.better) {
              t.better = d;
            }
          }
        }
      }
      for (int o = T >> 1; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, t.best, o), obb = __shfl_xor_sync(0xffffffffu, t.better, o);
        const int oi = __shfl_xor_sync(0xffffffffu, t.i, o);
        emd_merge(t, ob, obb, oi);
      }
      if (active && sub == 0) {
        const float inc = __fadd_rn(__fsub_rn(t.best, t.better), eps);
        m_bid[j] = t.i;
        m_binc[j] = inc;
      }
    }
    if (together) cluster.sync(); else __syncthreads();
    // ---- winners take their targets (:175-210) -----------------------------------------------------------------
    if (rank == 0) {
      // the highest increment per target, ties to the smallest source (CTA 0's own shared-memory atomics: the other
      // CTAs only post (target, increment) per source)
      if (!last) {
        for (int s = threadIdx.x; s < U; s += kEmdThreads) {
          const int j = (int)una[s];
          atomicMax(top + bid[j], ((unsigned long long)__float_as_uint(binc[j]) << 32) | (unsigned long long)(0xffffffffu - (unsigned)j));
        }
        __syncthreads();
      }
      for (int s = threadIdx.x; s < U; s += kEmdThreads) {
        const int j = (int)una[s];
        const int tg = bid[j];
        if (last || (int)(0xffffffffu - (unsigned)(top[tg] & 0xffffffffull)) == j) {
          const unsigned prev = inv[tg];
          if (!last && prev != 0xffffu) asg[prev] = 0xffffu;
          inv[tg] = (uint16_t)j;
          asg[j] = (uint16_t)tg;
          if (!last) {
            price[tg] = __fadd_rn(price[tg], binc[j]);
            top[tg] = 0ull;
          }
        }
      }
      __syncthreads();
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

inline cudaError_t emd_forward(const float* xyz1, const float* xyz2, float* dist, int* assignment, int B, int n, float eps,
                               int iters, cudaStream_t stream) {
  const size_t smem = emd_smem_bytes(n);
  cudaError_t e = cudaFuncSetAttribute(emd_auction_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // CTAs per cloud: one per 512 points (the first iterations bid for ~n sources against n targets; measured:
  // 32 x 2048 points 1.68 / 1.16 / 0.89 ms with 1 / 2 / 4 CTAs), at most 8 (the portable cluster size)
  static const int env_cs = getenv("CTB_EMD_CLUSTER") ? atoi(getenv("CTB_EMD_CLUSTER")) : 0;
  int cs = 1;
  while (cs < 8 && cs * 2 * 512 <= n) cs *= 2;
  if (env_cs >= 1 && env_cs <= 8 && (env_cs & (env_cs - 1)) == 0) cs = env_cs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)B * cs);
  cfg.blockDim = dim3(kEmdThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, emd_auction_kernel, xyz1, xyz2, dist, assignment, n, eps, iters);
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
