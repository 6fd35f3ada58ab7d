// ctb_syncbn.cuh -- SyncBatchNorm whose cross-GPU exchange is fused into its two kernels.
//
// train_classification.py:107-109 wraps the model in SyncBatchNorm + DDP.  torch's SyncBatchNorm runs, per layer and
// direction, statistics kernel -> NCCL all_gather (forward) / all_reduce (backward) of a few KB -> combine kernel ->
// elementwise kernel; model_zoo/scanobject/classifier.py has 106 such layers, i.e. 212 latency-bound collectives of
// ~40 us each at 8 ranks inside every step (profiles/r02_train_step_kernels_2gpu_ddp_syncbn.txt) -- the largest part of
// what separates 8-GPU training from 8 x one GPU once the step itself is a CUDA graph.
//
// Here the exchange is a ONE-SHOT all-gather over NVLink peer memory, written by the statistics kernel itself:
//   kernel 1 (grid channels x chunks): local per-channel sums, float4 loads; the last chunk of a channel to finish folds
//            the chunk partials in chunk order and stores the channel's two floats straight into the exchange buffer
//            of EVERY rank (peer pointers from torch's symmetric-memory rendezvous; NVSwitch gives each peer full
//            bandwidth); the last channel to finish publishes an epoch flag on every peer (release at system scope);
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
constexpr int kBnMaxChunks = 16;      // CTAs of the statistics kernels per channel

// per-layer state in symmetric memory, identical layout on every rank:
//   data  f32 [2 slots][W][2 C]     partial (sum_a, sum_b) of every rank
//   flag  u32 [2 slots][W]          epoch published by each source rank
// local (ordinary device memory): epoch u32 [1] = exchanges completed so far (the slot of exchange e is e & 1, so the
// alternation also holds when the same captured launch is replayed by a CUDA graph), done u32 [2] (channel counter of
// kernel 1 / CTA counter of kernel 2), scratch = float2 [C][kBnMaxChunks] chunk partials + u32 [C] chunk counters
struct BnExchange {
  float* const* peer_data;      // device array [W]: this layer's data block on every rank
  unsigned* const* peer_flag;   // device array [W]: this layer's flag block on every rank
  unsigned* epoch;              // local [1]
  unsigned* done;               // local [2]
  float2* scratch;              // local [C][kBnMaxChunks], then unsigned [C]
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

// publish (pa, pb) of channel c to every rank; the last channel of the grid raises this rank's flag on every peer
// (called by ONE CTA per channel, all its threads)
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

// ---- iteration over one channel of x f32 [B, C, L] ------------------------------------------------------------
// The channel's B rows of L floats are walked as one flat range of V-wide vectors (V = 4 when L % 4 == 0 and the
// tensors are 16-byte aligned, else 1); a CTA takes the contiguous part [lo, hi) of it and its threads stride through
// it, carrying (row, column) along instead of dividing per element.
template <int V>
struct ChannelWalk {
  int b, l, db, dl, LV;
  long long i, hi;
  size_t row_stride, base;
  __device__ __forceinline__ ChannelWalk(int B, int C, int L, int c, int part, int parts) {
    LV = L / V;
    const long long per = (long long)B * LV;
    const long long lo = per * part / parts;
    hi = per * (part + 1) / parts;
    i = lo + threadIdx.x;
    b = (int)(i / LV);
    l = (int)(i - (long long)b * LV);
    db = (int)blockDim.x / LV;
    dl = (int)blockDim.x - db * LV;
    row_stride = (size_t)C * L;
    base = (size_t)c * L;
  }
  __device__ __forceinline__ bool valid() const { return i < hi; }
  __device__ __forceinline__ size_t offset() const { return (size_t)b * row_stride + base + (size_t)l * V; }
  __device__ __forceinline__ void next() {
    i += blockDim.x;
    b += db;
    l += dl;
    if (l >= LV) {
      l -= LV;
      ++b;
    }
  }
};

template <int V> struct VecOf;
template <> struct VecOf<4> { using T = float4; };
template <> struct VecOf<1> { using T = float; };
__device__ __forceinline__ float4 ldv(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float ldv(const float* p) { return __ldg(p); }
__device__ __forceinline__ float comp(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }
__device__ __forceinline__ float comp(const float& v, int) { return v; }
__device__ __forceinline__ float& comp(float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }
__device__ __forceinline__ float& comp(float& v, int) { return v; }

// fold the chunk partials of channel c: every chunk CTA deposits (a, b); the last one to arrive returns true with the
// sum over chunks in chunk order (deterministic) and re-arms the counter
__device__ __forceinline__ bool bn_fold_chunks(const BnExchange& ex, int c, float2& t, float2* sh) {
  const int S = gridDim.y;
  if (S == 1) return true;
  unsigned* counter = reinterpret_cast<unsigned*>(ex.scratch + (size_t)ex.C * kBnMaxChunks) + c;
  float2* mine = ex.scratch + (size_t)c * kBnMaxChunks;
  __shared__ int last;
  if (threadIdx.x == 0) {
    __stcg(reinterpret_cast<float*>(mine + blockIdx.y), t.x);
    __stcg(reinterpret_cast<float*>(mine + blockIdx.y) + 1, t.y);
    __threadfence();
    const unsigned prev = atomicAdd(counter, 1u);
    last = prev == (unsigned)S - 1u;
    if (last) {
      *counter = 0;
      __threadfence();
      float a = 0.f, b = 0.f;
      for (int j = 0; j < S; ++j) {
        a += __ldcg(reinterpret_cast<const float*>(mine + j));
        b += __ldcg(reinterpret_cast<const float*>(mine + j) + 1);
      }
      sh[1] = make_float2(a, b);       // (slot 0 may still be being read by block_sum2's callers)
    }
  }
  __syncthreads();
  if (last) t = sh[1];
  return last != 0;
}

// ---- forward ------------------------------------------------------------------------------------------------
// grid (C, chunks).  (local mean, local sum of squared deviations) of every channel -> all ranks.  The sums run on
// x - K with K = the channel's first element, so E[d^2] - E[d]^2 does not cancel when |mean| >> std.
template <int V>
__global__ void __launch_bounds__(kBnThreads)
syncbn_fwd_stats_kernel(const float* __restrict__ x, BnExchange ex, int B, int L) {
  using T = typename VecOf<V>::T;
  __shared__ float2 scratch[32];
  const int c = blockIdx.x;
  const float K = __ldg(x + (size_t)c * L);
  float s = 0.f, q = 0.f;
  ChannelWalk<V> w(B, ex.C, L, c, blockIdx.y, gridDim.y);
#pragma unroll 4
  for (; w.valid(); w.next()) {
    const T v = ldv(reinterpret_cast<const T*>(x + w.offset()));
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float d = comp(v, k) - K;
      s += d;
      q = fmaf(d, d, q);
    }
  }
  float2 t = block_sum2(s, q, scratch);
  if (!bn_fold_chunks(ex, c, t, scratch)) return;
  const float n = (float)B * (float)L;
  bn_push(ex, c, K + t.x / n, fmaxf(t.y - t.x * t.x / n, 0.0f));
}

// grid (C, chunks): y = (x - mean) * invstd * w + b with the statistics of ALL ranks; chunk 0 saves mean / invstd and
// updates the running statistics (momentum, unbiased variance) like nn.SyncBatchNorm
template <int V>
__global__ void __launch_bounds__(kBnThreads)
syncbn_fwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ weight, const float* __restrict__ bias,
                        float* __restrict__ y, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                        float* __restrict__ running_mean, float* __restrict__ running_var, BnExchange ex, int B, int L,
                        float eps, float momentum) {
  using T = typename VecOf<V>::T;
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
  const float wt = weight ? __ldg(weight + c) : 1.0f, bb = bias ? __ldg(bias + c) : 0.0f;
  const float scale = invstd * wt, shift = bb - mean * scale;
  ChannelWalk<V> w(B, ex.C, L, c, blockIdx.y, gridDim.y);
#pragma unroll 4
  for (; w.valid(); w.next()) {
    const size_t o = w.offset();
    T v = ldv(reinterpret_cast<const T*>(x + o));
#pragma unroll
    for (int k = 0; k < V; ++k) comp(v, k) = fmaf(comp(v, k), scale, shift);
    *reinterpret_cast<T*>(y + o) = v;
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
// grid (C, chunks): local (sum dy, sum dy * xhat) = (grad_bias, grad_weight) of this rank -> all ranks
template <int V>
__global__ void __launch_bounds__(kBnThreads)
syncbn_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ save_mean,
                        const float* __restrict__ save_invstd, float* __restrict__ grad_weight,
                        float* __restrict__ grad_bias, BnExchange ex, int B, int L) {
  using T = typename VecOf<V>::T;
  __shared__ float2 scratch[32];
  const int c = blockIdx.x;
  const float mean = __ldg(save_mean + c), invstd = __ldg(save_invstd + c);
  float s = 0.f, q = 0.f;
  ChannelWalk<V> w(B, ex.C, L, c, blockIdx.y, gridDim.y);
#pragma unroll 4
  for (; w.valid(); w.next()) {
    const size_t o = w.offset();
    const T g = ldv(reinterpret_cast<const T*>(dy + o));
    const T v = ldv(reinterpret_cast<const T*>(x + o));
#pragma unroll
    for (int k = 0; k < V; ++k) {
      s += comp(g, k);
      q = fmaf(comp(g, k), (comp(v, k) - mean) * invstd, q);
    }
  }
  float2 t = block_sum2(s, q, scratch);
  if (!bn_fold_chunks(ex, c, t, scratch)) return;
  if (threadIdx.x == 0) {
    if (grad_bias) grad_bias[c] = t.x;
    if (grad_weight) grad_weight[c] = t.y;
  }
  bn_push(ex, c, t.x, t.y);
}

// dx = w * invstd * (dy - mean(dy) - xhat * mean(dy * xhat)), means over ALL ranks
template <int V>
__global__ void __launch_bounds__(kBnThreads)
syncbn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ weight,
                        const float* __restrict__ save_mean, const float* __restrict__ save_invstd, float* __restrict__ dx,
                        BnExchange ex, int B, int L) {
  using T = typename VecOf<V>::T;
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
  const float wt = weight ? __ldg(weight + c) : 1.0f;
  const float m_dy = tot.x / count, m_dyx = tot.y / count;
  const float k0 = wt * invstd;
  ChannelWalk<V> w(B, ex.C, L, c, blockIdx.y, gridDim.y);
#pragma unroll 4
  for (; w.valid(); w.next()) {
    const size_t o = w.offset();
    const T g = ldv(reinterpret_cast<const T*>(dy + o));
    const T v = ldv(reinterpret_cast<const T*>(x + o));
    T r;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float xhat = (comp(v, k) - mean) * invstd;
      comp(r, k) = k0 * (comp(g, k) - m_dy - xhat * m_dyx);
    }
    *reinterpret_cast<T*>(dx + o) = r;
  }
  bn_close(ex, gridDim.x * gridDim.y);
}

// CTAs per channel: enough for `waves` x 148 SMs x 4 resident CTAs, at least `min_per_cta` elements each
inline int bn_chunks(int C, long long per, int limit, int min_per_cta) {
  long long want = (2ll * 148 * 4 + C - 1) / C;
  const long long cap = (per + min_per_cta - 1) / min_per_cta;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  if (want > limit) want = limit;
  return (int)want;
}

}  // namespace ctb
