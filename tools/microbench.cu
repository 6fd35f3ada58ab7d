// microbench.cu -- B200 micro-measurements that drive the kernel design (DESIGN.md section "measured costs"):
// shared-memory atomics vs plain shared RMW, L2 atomics (scalar / v4), scattered global loads vs shared loads,
// and a float4 stream copy.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

constexpr int TILE = 8192;   // 32 KB of 4-byte cells (static shared limit is 48 KB)
constexpr int ITERS = 4096;

template <int MODE>  // 0 atomicMax s32, 1 atomicAdd f32, 2 plain RMW max, 3 plain load only
__global__ void smem_kernel(float* out) {
  __shared__ float tile[TILE];
  for (int i = threadIdx.x; i < TILE; i += blockDim.x) tile[i] = 0.f;
  __syncthreads();
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  float acc = 0.f;
  for (int it = 0; it < ITERS; ++it) {
    const int a = lcg(s) & (TILE - 1);
    const float v = (float)(s & 1023u);
    if (MODE == 0) atomicMax((int*)&tile[a], __float_as_int(v));
    else if (MODE == 1) atomicAdd(&tile[a], v);
    else if (MODE == 2) { float o = tile[a]; if (v > o) tile[a] = v; }
    else acc += tile[a];
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = tile[5] + acc;
  if (MODE == 3 && acc == 12345.f) out[0] = acc;
}

// address patterns of the channel-lane kernels: lanes = channels of one cell (consecutive words)
// PAT 0: 32 lanes on one random 32-word row     1: two half-warps on two random 16-word cells (pitch 16, any parity)
// PAT 2: as 1 but opposite parities (no conflict) 3: two half-warps, pitch 17     4: all lanes one word
// OP 0 atomicAdd s32, 1 atomicMax s32, 2 load, 3 two atomicAdd s32 into two arrays (the two-limb sum)
template <int OP, int PAT>
__global__ void smem_lane_kernel(float* out) {
  __shared__ int tile[TILE];
  for (int i = threadIdx.x; i < TILE; i += blockDim.x) tile[i] = 0;
  __syncthreads();
  uint32_t s = blockIdx.x * 7919u + (threadIdx.x >> 5) * 104729u + 1u;   // warp-uniform stream
  const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
  int acc = 0;
  for (int it = 0; it < ITERS; ++it) {
    const uint32_t r0 = lcg(s), r1 = lcg(s);
    int a;
    if (PAT == 0) a = (r0 & (TILE / 2 / 32 - 1)) * 32 + lane;
    else if (PAT == 1) a = ((half ? r1 : r0) & (TILE / 2 / 16 - 1)) * 16 + l16;
    else if (PAT == 2) a = ((((half ? r1 : r0) & (TILE / 2 / 32 - 1)) << 1) | half) * 16 + l16;
    else if (PAT == 3) a = ((half ? r1 : r0) & 127) * 17 + l16;
    else a = r0 & (TILE / 2 - 1);
    const int v = (int)(s & 1023u) + lane;
    if (OP == 0) atomicAdd(&tile[a], v);
    else if (OP == 1) atomicMax(&tile[a], v);
    else if (OP == 2) acc += tile[a];
    else { atomicAdd(&tile[a], v); atomicAdd(&tile[a + TILE / 2], v >> 3); }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (float)(tile[5] + acc);
  if (OP == 2 && acc == 12345) out[0] = (float)acc;
}

template <int MODE>  // 0 red.max.s32, 1 red.add.f32, 2 red.add.v4.f32, 3 scattered 4B load, 4 scattered 4B store
__global__ void gmem_kernel(float* buf, size_t n_elems, float* out, int iters) {
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
  float acc = 0.f;
  const size_t mask = n_elems - 1;
  for (int it = 0; it < iters; ++it) {
    size_t a = ((size_t)lcg(s) * 1u + ((size_t)lcg(s) << 20)) & mask;
    const float v = (float)(s & 1023u);
    if (MODE == 0) atomicMax((int*)&buf[a], __float_as_int(v));
    else if (MODE == 1) atomicAdd(&buf[a], v);
    else if (MODE == 2) {
      a &= ~(size_t)3;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(buf + a), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
    } else if (MODE == 3) acc += __ldg(&buf[a]);
    else buf[a] = v;
  }
  if (acc == 12345.f) out[0] = acc;
}

__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <typename F>
float time_ms(F f, int reps = 3) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
  float* out; CK(cudaMalloc(&out, 1 << 20));
  {
    const int blocks = sms * 2, threads = 512;
    const double lane_ops = (double)blocks * threads * ITERS;
    auto rep = [&](const char* name, float ms, double mult) {
      printf("%-58s %8.3f ms  %7.3f lane-ops/clk/SM\n", name, ms, lane_ops * mult / (ms * 1e-3) / sms / (p.clockRate * 1e3));
    };
    rep("smem lanes atomicAdd.s32  1 row of 32 words", time_ms([&] { smem_lane_kernel<0, 0><<<blocks, threads>>>(out); }), 1);
    rep("smem lanes atomicAdd.s32  2 cells pitch 16 random", time_ms([&] { smem_lane_kernel<0, 1><<<blocks, threads>>>(out); }), 1);
    rep("smem lanes atomicAdd.s32  2 cells pitch 16 opposite parity", time_ms([&] { smem_lane_kernel<0, 2><<<blocks, threads>>>(out); }), 1);
    rep("smem lanes atomicAdd.s32  2 cells pitch 17", time_ms([&] { smem_lane_kernel<0, 3><<<blocks, threads>>>(out); }), 1);
    rep("smem lanes atomicAdd.s32  all lanes one word", time_ms([&] { smem_lane_kernel<0, 4><<<blocks, threads>>>(out); }), 1);
    rep("smem lanes atomicMax.s32  1 row of 32 words", time_ms([&] { smem_lane_kernel<1, 0><<<blocks, threads>>>(out); }), 1);
    rep("smem lanes atomicMax.s32  2 cells pitch 16 opposite parity", time_ms([&] { smem_lane_kernel<1, 2><<<blocks, threads>>>(out); }), 1);
    rep("smem lanes load           1 row of 32 words", time_ms([&] { smem_lane_kernel<2, 0><<<blocks, threads>>>(out); }), 1);
    rep("smem lanes load           2 cells pitch 17", time_ms([&] { smem_lane_kernel<2, 3><<<blocks, threads>>>(out); }), 1);
    rep("smem lanes 2x atomicAdd   1 row of 32 words (per atomic)", time_ms([&] { smem_lane_kernel<3, 0><<<blocks, threads>>>(out); }), 2);
    rep("smem lanes 2x atomicAdd   2 cells opposite parity (per atomic)", time_ms([&] { smem_lane_kernel<3, 2><<<blocks, threads>>>(out); }), 2);
  }
  const char* sn[4] = {"smem atomicMax.s32 random", "smem atomicAdd.f32 random", "smem plain RMW max random", "smem plain load random"};
  for (int ctas = 1; ctas <= 2; ++ctas) {
    const int blocks = sms * ctas, threads = 512;
    float ms[4];
    ms[0] = time_ms([&] { smem_kernel<0><<<blocks, threads>>>(out); });
    ms[1] = time_ms([&] { smem_kernel<1><<<blocks, threads>>>(out); });
    ms[2] = time_ms([&] { smem_kernel<2><<<blocks, threads>>>(out); });
    ms[3] = time_ms([&] { smem_kernel<3><<<blocks, threads>>>(out); });
    for (int m = 0; m < 4; ++m) {
      const double ops = (double)blocks * threads * ITERS;
      printf("%-28s ctas/SM=%d: %8.3f ms  %8.2f Glane-ops/s  %6.3f lane-ops/clk/SM (at %.2f GHz)\n", sn[m], ctas, ms[m],
             ops / ms[m] / 1e6, ops / ms[m] / 1e6 / sms / (p.clockRate / 1e6), p.clockRate / 1e6);
    }
  }
  const char* gn[5] = {"gmem red.max.s32 scattered", "gmem red.add.f32 scattered", "gmem red.add.v4.f32 scattered", "gmem ld 4B scattered", "gmem st 4B scattered"};
  for (size_t mb : {8, 64, 1024}) {
    const size_t n = mb * 1024 * 1024 / 4;
    float* buf; CK(cudaMalloc(&buf, n * 4)); CK(cudaMemset(buf, 0, n * 4));
    const int blocks = sms * 8, threads = 256, iters = 512;
    float ms[5];
    ms[0] = time_ms([&] { gmem_kernel<0><<<blocks, threads>>>(buf, n, out, iters); });
    ms[1] = time_ms([&] { gmem_kernel<1><<<blocks, threads>>>(buf, n, out, iters); });
    ms[2] = time_ms([&] { gmem_kernel<2><<<blocks, threads>>>(buf, n, out, iters); });
    ms[3] = time_ms([&] { gmem_kernel<3><<<blocks, threads>>>(buf, n, out, iters); });
    ms[4] = time_ms([&] { gmem_kernel<4><<<blocks, threads>>>(buf, n, out, iters); });
    for (int m = 0; m < 5; ++m) {
      const double ops = (double)blocks * threads * iters;
      printf("%-30s region %4zu MB: %8.3f ms  %8.2f Glane-ops/s  %6.3f lane-ops/clk/SM\n", gn[m], mb, ms[m], ops / ms[m] / 1e6,
             ops / ms[m] / 1e6 / sms / (p.clockRate / 1e6));
    }
    CK(cudaFree(buf));
  }
  {
    const size_t n = (size_t)1 << 28;  // 1 GiB each
    float4 *a, *b; CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMemset(a, 1, n * 4));
    float ms = time_ms([&] { copy_kernel<<<sms * 16, 512>>>(a, b, n / 4); }, 5);
    printf("float4 copy 1 GiB: %.3f ms  %.1f GB/s (read+write)\n", ms, 2.0 * n * 4 / ms / 1e6);
    float msm = time_ms([&] { cudaMemsetAsync(b, 0, n * 4); }, 5);
    printf("memset 1 GiB: %.3f ms  %.1f GB/s (write)\n", msm, 1.0 * n * 4 / msm / 1e6);
  }
  return 0;
}
