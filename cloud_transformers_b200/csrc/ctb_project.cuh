// ctb_project.cuh -- A8: the callers' glue in front of DifferentiablePositions, fused into one kernel each way.
//
// Reference (layers/utils.py:25-34 VolTransformer, :53-61 PlaneTransformer; layers/multihead_ct.py:93-97,
// multihead_ct_adain.py:112-115, multihead_ct_pool.py:61-66):
//     p      = orig_pcd[b, :, n] + res_scale * keys_res[b, h, :, n] + shift[h]          (3-vector)
//     q_j    = sum_c p_c * R[h][c][j]                    einsum 'bhcp,hcn->bhnp', R = so3_exponential_map(log_R)
//     q_j   *= scales[h][j]                              (optional)
//     keys[b, h*d + j, n] = tanh(q_j),  j < d            (Plane keeps the first two coordinates)
// The reference runs this as ~6 elementwise / bmm kernels over [B,H,3,N]; here it is one pass forward and one
// pass backward (grad_pcd summed over heads in registers, parameter gradients warp-reduced then accumulated).
// R itself (12 floats per head) is computed by the caller in PyTorch so autograd carries d R / d log_R.
#pragma once
#include <cuda_runtime.h>

namespace ctb {

constexpr int kProjBlock = 256;
constexpr int kProjAcc = 16;   // per-head accumulator row: shift[3] R[9] scales[3] res_scale[1]

// key_stats (optional, f64 [2], accumulated): sum and sum of squares of the pre-tanh keys q -- what the reference
// logs per block as mean(keys) / var(keys) (layers/multihead_ct.py:109-113) -- so that the pre-tanh tensor never has to
// be materialised for the statistics.
template <int D>
__global__ void __launch_bounds__(kProjBlock)
project_fwd_kernel(const float* __restrict__ pcd, const float* __restrict__ res, float res_scale,
                   const float* __restrict__ shift, const float* __restrict__ rot, const float* __restrict__ scales,
                   float* __restrict__ keys, double* __restrict__ key_stats, int H, int N, int chunks) {
  const int unit = blockIdx.x / chunks;
  const int n = (blockIdx.x % chunks) * kProjBlock + threadIdx.x;
  const bool live = n < N;
  const int b = unit / H, h = unit % H;
  float s1 = 0.0f, s2 = 0.0f;
  if (live) {
    float p[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      p[c] = __ldg(pcd + ((size_t)b * 3 + c) * N + n);
      if (res) p[c] += res_scale * __ldg(res + ((size_t)unit * 3 + c) * N + n);
      p[c] += __ldg(shift + h * 3 + c);
    }
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float q = p[0] * __ldg(rot + h * 9 + j);
      q = fmaf(p[1], __ldg(rot + h * 9 + 3 + j), q);
      q = fmaf(p[2], __ldg(rot + h * 9 + 6 + j), q);
      if (scales) q *= __ldg(scales + h * D + j);
      keys[((size_t)unit * D + j) * N + n] = tanhf(q);
      s1 += q;
      s2 = fmaf(q, q, s2);
    }
  }
  if (key_stats != nullptr) {
    double d1 = (double)s1, d2 = (double)s2;
    for (int o = 16; o > 0; o >>= 1) {
      d1 += __shfl_xor_sync(0xffffffffu, d1, o);
      d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(key_stats, d1);
      atomicAdd(key_stats + 1, d2);
    }
  }
}

// Backward.  CTA = 32 consecutive points x up to 32 heads (one warp per head): the per-head parameter gradients are
// warp-reduced over the 32 points and written as per-CTA partials (summed in a fixed order by
// project_param_reduce_kernel), grad_pcd is summed over the heads through shared memory in a fixed order --
// everything is deterministic.  partial is [chunks * B][H][kProjAcc].
template <int D>
__global__ void __launch_bounds__(1024)
project_bwd_kernel(const float* __restrict__ pcd, const float* __restrict__ res, float res_scale,
                   const float* __restrict__ shift, const float* __restrict__ rot, const float* __restrict__ scales,
                   const float* __restrict__ keys, const float* __restrict__ grad_keys,
                   float* __restrict__ grad_pcd, float* __restrict__ grad_res, float* __restrict__ partial, int H, int N,
                   int chunks) {
  __shared__ float gp_s[32][3][33];
  const int b = blockIdx.x / chunks;
  const int n = (blockIdx.x % chunks) * 32 + threadIdx.x;
  const bool live = n < N;
  const int hw = threadIdx.y;                 // warp index = head slot
  float pc[3] = {0.f, 0.f, 0.f}, gsum[3] = {0.f, 0.f, 0.f};
  if (live)
#pragma unroll
    for (int c = 0; c < 3; ++c) pc[c] = __ldg(pcd + ((size_t)b * 3 + c) * N + n);
  for (int h0 = 0; h0 < H; h0 += blockDim.y) {
    const int h = h0 + hw;
    float v[kProjAcc];
#pragma unroll
    for (int i = 0; i < kProjAcc; ++i) v[i] = 0.0f;
    float gp[3] = {0.f, 0.f, 0.f};
    if (live && h < H) {
      const size_t unit = (size_t)b * H + h;
      float p[3], r[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (res) r[c] = __ldg(res + (unit * 3 + c) * N + n);
        p[c] = pc[c] + res_scale * r[c] + __ldg(shift + h * 3 + c);
      }
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const float t = __ldg(keys + (unit * D + j) * N + n);
        float gq = __ldg(grad_keys + (unit * D + j) * N + n) * (1.0f - t * t);   // through tanh
        const float r0 = __ldg(rot + h * 9 + j), r1 = __ldg(rot + h * 9 + 3 + j), r2 = __ldg(rot + h * 9 + 6 + j);
        if (scales) {
          const float q = fmaf(p[2], r2, fmaf(p[1], r1, p[0] * r0));
          v[12 + j] = gq * q;                       // d / d scales[h][j]
          gq *= __ldg(scales + h * D + j);
        }
        v[3 + 0 + j] = p[0] * gq;                    // d / d R[h][c][j]
        v[3 + 3 + j] = p[1] * gq;
        v[3 + 6 + j] = p[2] * gq;
        gp[0] = fmaf(gq, r0, gp[0]);
        gp[1] = fmaf(gq, r1, gp[1]);
        gp[2] = fmaf(gq, r2, gp[2]);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        v[c] = gp[c];                                // d / d shift[h][c]
        if (grad_res) grad_res[(unit * 3 + c) * N + n] = res_scale * gp[c];
        v[15] = fmaf(gp[c], r[c], v[15]);            // d / d res_scale
      }
    }
    // parameter gradients: reduce over the 32 points of this warp, one partial row per (CTA, head)
#pragma unroll
    for (int i = 0; i < kProjAcc; ++i) {
      float x = v[i];
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (threadIdx.x == 0 && h < H) partial[((size_t)blockIdx.x * H + h) * kProjAcc + i] = x;
    }
    // grad_pcd: sum over heads in a fixed order through shared memory
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; ++c) gp_s[hw][c][threadIdx.x] = gp[c];
    __syncthreads();
    if (hw == 0)
      for (int w = 0; w < (int)blockDim.y; ++w)
#pragma unroll
        for (int c = 0; c < 3; ++c) gsum[c] += gp_s[w][c][threadIdx.x];
  }
  if (hw == 0 && live)
#pragma unroll
    for (int c = 0; c < 3; ++c) grad_pcd[((size_t)b * 3 + c) * N + n] = gsum[c];
}

// acc[h][i] += sum over CTAs of partial[cta][h][i]: one 128-thread block per (h, i), fixed reduction tree
__global__ void __launch_bounds__(128)
project_param_reduce_kernel(const float* __restrict__ partial, float* __restrict__ acc, int ctas, int H) {
  __shared__ float ws[4];
  const int t = blockIdx.x;
  const int stride = H * kProjAcc;
  float s = 0.0f;
  for (int c = threadIdx.x; c < ctas; c += 128) s += partial[(size_t)c * stride + t];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) acc[t] += (ws[0] + ws[1]) + (ws[2] + ws[3]);
}

}  // namespace ctb
