// ctb_syncbn.cuh -- SyncBatchNorm whose cross-GPU exchange is fused into its two kernels.
//
// train_classification.py:107-109 wraps the model in SyncBatchNorm + DDP.  torch's SyncBatchNorm runs, per layer and
// direction, statistics kernel -> NCCL all_gather (forward) / all_reduce (backward) of a few KB -> combine kernel ->
// elementwise kernel; model_zoo/scanobject/classifier.py has 106 such layers, i.e. 212 latency-bound collectives of
// ~40 us each at 8 ranks inside every step (profiles/r02_train_step_kernels_2gpu_ddp_syncbn.txt) -- the largest part of
// what separates 8-GPU training from 8 x one GPU once the step itself is a CUDA graph.
//
// Here the exchange is a ONE-SHOT all-gather over NVLink peer memory, written by the statistics kernel itself:
//   kernel 1 (one CTA per channel): local per-channel sums; every CTA stores its two floats straight into the exchange
//            buffer of EVERY rank (peer pointers from torch's symmetric-memory rendezvous; NVSwitch gives each peer full
//            bandwidth), the last CTA to finish publishes an epoch flag on every peer (release at system scope);
//   kernel 2 (elementwise): waits for the flags of all ranks (acquire at system scope), sums the W partials of its
//            channel in rank order -- identical bits on every rank -- and applies normalisation (forward) or the
//            input gradient (backward).
// Two slots per layer alternate, so a fast rank's next exchange never overwrites data a slow rank is still reading
// (a rank cannot start exchange k + 2 before every rank has written k + 1).  Epochs live in device memory and all
// pointers are fixed, so the kernels replay inside a CUDA graph.  Equal per-rank batch sizes are assumed (weak scaling).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ctb {

constexpr int kBnThreads = 512;

// per-layer state in symmetric memory, identical layout on every rank:
//   data  f32 [2 slots][W][2 C]     partial (sum_a, sum_b) of every rank
//   flag  u32 [2 slots][W]          epoch published by each source rank
// local (ordinary device memory): epoch u32 [1] = exchanges completed so far (the slot of exchange e is e & 1, so the
// alternation also holds when the same captured launch is replayed by a CUDA graph), done u32 [2] (CTA counters of
// kernel 1 / kernel 2)
struct BnExchange {
  float* const* peer_data;      // device array [W]: this layer's data block on every rank
  unsigned* const* peer_flag;   // device array [W]: this layer's flag block on every rank
  unsigned* epoch;              // local [1]
  unsigned* done;               // local [2]
  int rank, world, C;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ float2 block_sum2(float a, float b, float2* scratch) {
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) scratch[warp] = make_float2(a, b);
  __syncthreads();
  if (warp == 0) {
    float2 v = lane < (int)(blockDim.x >> 5) ? scratch[lane] : make_float2(0.f, 0.f);
    for (int o = 16; o > 0; o >>= 1) {
      v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
      v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    if (lane == 0) scratch[0] = v;
  }
  __syncthreads();
  return scratch[0];
}

// publish (pa, pb) of channel c to every rank; the last CTA of the grid raises this rank's flag on every peer
__device__ __forceinline__ void bn_push(const BnExchange& ex, int c, float pa, float pb) {
  const unsigned e = *ex.epoch + 1u;
  const int slot = (int)(e & 1u);
  if (threadIdx.x < ex.world) {
    float* dst = ex.peer_data[threadIdx.x] + ((size_t)(slot * ex.world + ex.rank) * ex.C + c) * 2;
    reinterpret_cast<float2*>(dst)[0] = make_float2(pa, pb);
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(ex.done + 0, 1u);
    if (prev == (unsigned)gridDim.x - 1u) {
      ex.done[0] = 0;
      __threadfence_system();
      for (int r = 0; r < ex.world; ++r) st_release_sys(ex.peer_flag[r] + slot * ex.world + ex.rank, e);
    }
  }
}

// wait until every rank has published epoch e for this slot, then fetch channel c's partial of every rank
// (part[r], r = 0 .. world - 1; the caller combines them in rank order: identical bits on every rank)
__device__ __forceinline__ void bn_pull(const BnExchange& ex, int c, float2* part) {
  const unsigned e = *ex.epoch + 1u;
  const int slot = (int)(e & 1u);
  if (threadIdx.x < ex.world) {
    const unsigned* f = ex.peer_flag[ex.rank] + slot * ex.world + threadIdx.x;
    while ((int)(ld_acquire_sys(f) - e) < 0) {
    }
    const float* pv = ex.peer_data[ex.rank] + ((size_t)(slot * ex.world + threadIdx.x) * ex.C + c) * 2;
    part[threadIdx.x] = make_float2(__ldcg(pv), __ldcg(pv + 1));     // (written by a peer: bypass L1)
  }
  __syncthreads();
}

// the last CTA of kernel 2 closes the epoch of the slot
__device__ __forceinline__ void bn_close(const BnExchange& ex, unsigned total_ctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(ex.done + 1, 1u);
    if (prev == total_ctas - 1u) {
      ex.done[1] = 0;
      *ex.epoch = *ex.epoch + 1u;
    }
  }
}

// ---- forward ------------------------------------------------------------------------------------------------
// x f32 [B, C, L].  One CTA per channel: (local mean, local sum of squared deviations) -> all ranks.  The sums run on
// x - K with K = the channel's first element, so E[d^2] - E[d]^2 does not cancel when |mean| >> std.
__global__ void __launch_bounds__(kBnThreads)
syncbn_fwd_stats_kernel(const float* __restrict__ x, BnExchange ex, int B, int L) {
  __shared__ float2 scratch[32];
  const int c = blockIdx.x;
  const float K = __ldg(x + (size_t)c * L);
  float s = 0.f, q = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* p = x + ((size_t)b * ex.C + c) * L;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      const float d = __ldg(p + l) - K;
      s += d;
      q = fmaf(d, d, q);
    }
  }
  const float2 t = block_sum2(s, q, scratch);
  const float n = (float)B * (float)L;
  bn_push(ex, c, K + t.x / n, fmaxf(t.y - t.x * t.x / n, 0.0f));
}

// grid (C, chunks): y = (x - mean) * invstd * w + b with the statistics of ALL ranks; chunk 0 saves mean / invstd and
// updates the running statistics (momentum, unbiased variance) like nn.SyncBatchNorm
__global__ void __launch_bounds__(kBnThreads)
syncbn_fwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ weight, const float* __restrict__ bias,
                        float* __restrict__ y, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                        float* __restrict__ running_mean, float* __restrict__ running_var, BnExchange ex, int B, int L,
                        float eps, float momentum) {
  __shared__ float2 part[32];
  const int c = blockIdx.x;
  bn_pull(ex, c, part);
  // merge the ranks' (mean, M2) with equal counts n: mean = avg of means, M2 = sum M2_r + n sum (mean_r - mean)^2
  const float n = (float)B * (float)L, count = (float)ex.world * n;
  float mean = 0.0f;
  for (int r = 0; r < ex.world; ++r) mean += part[r].x;
  mean /= (float)ex.world;
  float m2 = 0.0f;
  for (int r = 0; r < ex.world; ++r) {
    const float dm = part[r].x - mean;
    m2 += part[r].y + n * dm * dm;
  }
  const float var = m2 / count;
  const float invstd = rsqrtf(var + eps);
  const float w = weight ? __ldg(weight + c) : 1.0f, bb = bias ? __ldg(bias + c) : 0.0f;
  const float scale = invstd * w, shift = bb - mean * scale;
  const long long per = (long long)B * L;
  for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.y * blockDim.x) {
    const int b = (int)(i / L), l = (int)(i - (long long)b * L);
    const size_t o = ((size_t)b * ex.C + c) * L + l;
    y[o] = fmaf(__ldg(x + o), scale, shift);
  }
  if (blockIdx.y == 0 && threadIdx.x == 0) {
    save_mean[c] = mean;
    save_invstd[c] = invstd;
    if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
    if (running_var) running_var[c] = (1.0f - momentum) * running_var[c] + momentum * var * (count / fmaxf(count - 1.0f, 1.0f));
  }
  bn_close(ex, gridDim.x * gridDim.y);
}

// ---- backward -----------------------------------------------------------------------------------------------
// One CTA per channel: local (sum dy, sum dy * xhat) = (grad_bias, grad_weight) of this rank -> all ranks
__global__ void __launch_bounds__(kBnThreads)
syncbn_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ save_mean,
                        const float* __restrict__ save_invstd, float* __restrict__ grad_weight,
                        float* __restrict__ grad_bias, BnExchange ex, int B, int L) {
  __shared__ float2 scratch[32];
  const int c = blockIdx.x;
  const float mean = __ldg(save_mean + c), invstd = __ldg(save_invstd + c);
  float s = 0.f, q = 0.f;
  for (int b = 0; b < B; ++b) {
    const size_t o = ((size_t)b * ex.C + c) * L;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      const float g = __ldg(dy + o + l);
      s += g;
      q = fmaf(g, (__ldg(x + o + l) - mean) * invstd, q);
    }
  }
  const float2 t = block_sum2(s, q, scratch);
  if (threadIdx.x == 0) {
    if (grad_bias) grad_bias[c] = t.x;
    if (grad_weight) grad_weight[c] = t.y;
  }
  bn_push(ex, c, t.x, t.y);
}

// dx = w * invstd * (dy - mean(dy) - xhat * mean(dy * xhat)), means over ALL ranks
__global__ void __launch_bounds__(kBnThreads)
syncbn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ weight,
                        const float* __restrict__ save_mean, const float* __restrict__ save_invstd, float* __restrict__ dx,
                        BnExchange ex, int B, int L) {
  __shared__ float2 part[32];
  const int c = blockIdx.x;
  bn_pull(ex, c, part);
  float2 tot = make_float2(0.f, 0.f);
  for (int r = 0; r < ex.world; ++r) {
    tot.x += part[r].x;
    tot.y += part[r].y;
  }
  const float count = (float)ex.world * (float)B * (float)L;
  const float mean = __ldg(save_mean + c), invstd = __ldg(save_invstd + c);
  const float w = weight ? __ldg(weight + c) : 1.0f;
  const float m_dy = tot.x / count, m_dyx = tot.y / count;
  const float k = w * invstd;
  const long long per = (long long)B * L;
  for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.y * blockDim.x) {
    const int b = (int)(i / L), l = (int)(i - (long long)b * L);
    const size_t o = ((size_t)b * ex.C + c) * L + l;
    const float xhat = (__ldg(x + o) - mean) * invstd;
    dx[o] = k * (__ldg(dy + o) - m_dy - xhat * m_dyx);
  }
  bn_close(ex, gridDim.x * gridDim.y);
}

inline int bn_chunks(int C, long long per) {
  // enough CTAs for ~4 waves of 148 SMs, at least 4096 elements per CTA
  long long want = (4ll * 148 * 2 + C - 1) / C;
  const long long cap = (per + 4095) / 4096;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  return (int)want;
}

}  // namespace ctb
