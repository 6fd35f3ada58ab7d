// ctb_tile_cl.cuh -- channel-lane scatter for coarse, dense grids (the channel-last tile layout).
//
// On 8^3 / 16^2 grids real clouds put tens to hundreds of points into one cell.  With one lane per POINT
// (ctb_tile.cuh) the lanes of a warp then hit the same shared-memory word: atomics serialise and ncu shows
// 75-85 % of the shared wavefronts as bank conflicts (profiles/r01_*).  Here one lane owns one CHANNEL of a
// point instead: the 2^d atomics of a point go to consecutive words of the channel-last tile [cell][fg|1] --
// conflict-free inside the warp -- and only different warps can still meet in a cell.
//
// Work item = (unit, channel group <= 32), whole grid in the tile.  Points are processed in chunks:
//   A. all threads: features of the chunk -> shared [fg][chunk+1] with coalesced loads (the transpose that
//      turns "lane = point" loads into "lane = channel" reads), positions of the chunk (lane = point) -> shared
//   B. every warp walks points of the chunk, lanes = channels (32 / LP points at a time for LP-lane groups)
// Arithmetic and results are identical to tile_scatter_kernel (same products, integer max / min / add).
#pragma once
#include "ctb_tile.cuh"

namespace ctb {

constexpr int kClChunk = 128;   // points per chunk (power of two)

__host__ __device__ inline int cl_stage_words(int FG) { return (FG * (kClChunk + 1) + 3) & ~3; }
inline size_t cl_extra_bytes(int FG, int dim) {
  return (size_t)cl_stage_words(FG) * 4 + (size_t)kClChunk * (1 << dim) * 8 + 16;
}

template <int D, bool SUM, typename GT>
__global__ void __launch_bounds__(kTileThreads, 2)
cl_scatter_kernel(const float* __restrict__ keys, const float* __restrict__ feat, const float* __restrict__ pad,
                  GT* __restrict__ z, int* __restrict__ arg, Grid<D> g, int H, int F, int N, int FG, int groups,
                  int tw, int LP) {
  constexpr int S = 1 << D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int fgp = cl_pitch(FG, TILE_CL);
  const bool want_arg = !SUM && arg != nullptr;
  float* tval = (float*)smem_raw;                                    // max: value bits   | sum: low limb
  int* targ = (int*)(tval + tw);                                     // max: arg (if any) | sum: high limb
  float* xs = (float*)(targ + ((SUM || want_arg) ? tw : 0));         // [FG][kClChunk + 1]
  int* pa = (int*)(xs + cl_stage_words(FG));                         // [kClChunk][S] word offset of the corner cell
  float* pw = (float*)(pa + kClChunk * S);                           // [kClChunk][S] corner weight
  int* counter = (int*)(pw + kClChunk * S);                          // [1] max|v| bits, [2] non-finite

  const int f0 = (blockIdx.x % groups) * FG;
  const int unit = blockIdx.x / groups;
  const int fg = min(FG, F - f0);
  {
    float4* t4 = reinterpret_cast<float4*>(tval);
    int4* a4 = reinterpret_cast<int4*>(targ);
    for (int i = threadIdx.x; i < (tw >> 2); i += kTileThreads) {
      t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (SUM) a4[i] = make_int4(0, 0, 0, 0);
      else if (want_arg) a4[i] = make_int4(-1, -1, -1, -1);
    }
    if (threadIdx.x == 0) counter[1] = counter[2] = 0;
  }
  const float* ku = keys + (size_t)unit * D * N;
  const float* pu = pad ? pad + (size_t)(unit / H) * N : nullptr;
  const float* fu = feat + ((size_t)unit * F + f0) * N;
  __syncthreads();

  // fixed-point scale of the sum (see tile_scatter_kernel)
  bool fixed_point = false;
  int limb_bits = 0;
  float scale = 1.0f, inv_scale = 1.0f;
  if constexpr (SUM) {
    float m = 0.0f;
    bool bad = false;
    for (int i = threadIdx.x; i < fg * N; i += kTileThreads) {
      const int n = i % N;
      const float v = fabsf(__ldg(fu + i) * (pu ? __ldg(pu + n) : 1.0f));
      bad |= !(v <= 3.0e38f);
      m = fmaxf(m, v);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    bad = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
      atomicMax(counter + 1, __float_as_int(m));
      if (bad) counter[2] = 1;
    }
    __syncthreads();
    const float M = __int_as_float(counter[1]);
    fixed_point = counter[2] == 0;
    const int cnt_bits = 32 - __clz(N > 1 ? N - 1 : 1);
    if (cnt_bits <= 11) limb_bits = 32 - cnt_bits;
    if (fixed_point && M > 0.0f) {
      int k = limb_bits > 0 ? (2 * limb_bits - 1) - (ilogbf(M) + 1) : 62 - (ilogbf(M) + 1) - (cnt_bits + 1);
      k = k > 120 ? 120 : k;
      scale = ldexpf(1.0f, k);
      inv_scale = ldexpf(1.0f, -k);
    }
  }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ch = lane % LP;                 // my channel inside the group
  const int sub = lane / LP;                // which of the 32 / LP points of this warp step
  const int ppw = 32 / LP;                  // points per warp step
  const bool ch_ok = ch < fg;
  const unsigned lmask = limb_bits > 0 ? (1u << limb_bits) - 1u : 0u;

#pragma unroll 1
  for (int pass = 0; pass < (want_arg ? 2 : 1); ++pass) {
#pragma unroll 1
    for (int c0 = 0; c0 < N; c0 += kClChunk) {
      const int pcn = min(kClChunk, N - c0);
      // A. stage features (coalesced along the points) and positions of the chunk
      for (int i = threadIdx.x; i < fg * kClChunk; i += kTileThreads) {
        const int f = i / kClChunk, j = i % kClChunk;
        float v = 0.0f;
        if (j < pcn) {
          v = __ldg(fu + (size_t)f * N + c0 + j);
          if (pu) v = CTB_FMUL(v, __ldg(pu + c0 + j));
        }
        xs[f * (kClChunk + 1) + j] = v;
      }
      if (threadIdx.x < pcn) {
        const Pos<D> p = point_pos<D>(ku, c0 + threadIdx.x, N, g);
#pragma unroll
        for (int s = 0; s < S; ++s) {
          pa[threadIdx.x * S + s] = (p.base + corner_offset<D>(g, s)) * fgp;
          pw[threadIdx.x * S + s] = corner_weight<D>(p, s);
        }
      }
      __syncthreads();
      // B. lanes = channels
      for (int j = warp * ppw + sub; j < pcn; j += (kTileThreads / 32) * ppw) {
        if (!ch_ok) continue;
        const float x = xs[ch * (kClChunk + 1) + j];
        int a[S];
        float w[S];
        if constexpr (S == 8) {
          const int4 a0 = reinterpret_cast<const int4*>(pa)[j * 2], a1 = reinterpret_cast<const int4*>(pa)[j * 2 + 1];
          const float4 w0 = reinterpret_cast<const float4*>(pw)[j * 2], w1 = reinterpret_cast<const float4*>(pw)[j * 2 + 1];
          a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
          w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
        } else {
          const int4 a0 = reinterpret_cast<const int4*>(pa)[j];
          const float4 w0 = reinterpret_cast<const float4*>(pw)[j];
          a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
          w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
        }
        const int n = c0 + j;
        if constexpr (SUM) {
          if (fixed_point && limb_bits > 0) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const long long q = __float2ll_rn(CTB_FMUL(CTB_FMUL(x, w[s]), scale));
              atomicAdd((unsigned*)tval + a[s] + ch, (unsigned)q & lmask);
              atomicAdd(targ + a[s] + ch, (int)(q >> limb_bits));
            }
          } else if (fixed_point) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const long long q = __float2ll_rn(CTB_FMUL(CTB_FMUL(x, w[s]), scale));
              const unsigned lo = (unsigned)q;
              const unsigned old = atomicAdd((unsigned*)tval + a[s] + ch, lo);
              const int hi = (int)(q >> 32) + ((unsigned)(old + lo) < old ? 1 : 0);
              if (hi != 0) atomicAdd(targ + a[s] + ch, hi);
            }
          } else {
#pragma unroll
            for (int s = 0; s < S; ++s) atomicAdd(tval + a[s] + ch, CTB_FMUL(x, w[s]));
          }
        } else if (pass == 0) {
#pragma unroll
          for (int s = 0; s < S; ++s)
            atomicMax((int*)tval + a[s] + ch, __float_as_int(fmaxf(CTB_FMUL(x, w[s]), 0.0f)));
        } else {
          bool hit[S];
          bool any = false;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const int t = ((const int*)tval)[a[s] + ch];
            hit[s] = (__float_as_int(CTB_FMUL(x, w[s])) == t) & (t != 0);
            any |= hit[s];
          }
          if (any) {
#pragma unroll
            for (int s = 0; s < S; ++s)
              if (hit[s]) atomicMin((unsigned*)targ + a[s] + ch, (unsigned)(s * N + n));
          }
        }
      }
      __syncthreads();
    }
  }

  GT* zu = z + ((size_t)unit * F + f0) * g.C;
  int* au = want_arg ? arg + ((size_t)unit * F + f0) * g.C : nullptr;
  for_each_plane_element(fg, g.C, [&](int f, int r) {
    float v1 = tval[r * fgp + f];
    if (SUM && fixed_point) {
      const long long lo = (long long)(unsigned)__float_as_int(v1);
      const int hi = targ[r * fgp + f];
      const long long q = limb_bits > 0 ? ((long long)hi << limb_bits) + lo : (((long long)hi << 32) | lo);
      v1 = __ll2float_rn(q) * inv_scale;
    }
    grid_store(zu + (size_t)f * g.C + r, v1);
    if (want_arg) __stcs(au + (size_t)f * g.C + r, targ[r * fgp + f]);
  });
}

// host: use the channel-lane kernel when the shape is channel-last, fits whole, and a group has <= 32 channels
template <int D, typename GT>
bool cl_scatter_try(const float* keys, const float* feat, const float* pad, GT* z, int* arg, const ctb_shape* s,
                    bool sum, cudaStream_t stream, cudaError_t* err) {
  TileConfig c;
  if (!tile_scatter_config(s, sum, arg != nullptr, &c)) return false;
  if ((c.layout != TILE_CL && c.layout != TILE_CLQ) || c.slabs != 1) return false;
  const int arrays = (sum || arg != nullptr) ? 2 : 1;
  // re-fit the channel group with the staging buffers included
  int FG = c.FG > 32 ? 32 : c.FG;
  const int cells = s->size[0] * (s->dim == 2 ? s->size[1] : s->size[1] * s->size[2]);
  auto bytes = [&](int fg) { return (size_t)tile_array_words(cells, fg, TILE_CL) * 4 * arrays + cl_extra_bytes(fg, s->dim); };
  while (FG > 1 && bytes(FG) > (size_t)kTileSmemTwoCtas) --FG;
  if (bytes(FG) > (size_t)kTileSmemTwoCtas || FG < 4) return false;
  int groups = (s->F + FG - 1) / FG;
  FG = (s->F + groups - 1) / groups;
  groups = (s->F + FG - 1) / FG;
  int LP = 1;
  while (LP < FG) LP <<= 1;
  const int tw = tile_array_words(cells, FG, TILE_CL);
  const size_t smem = bytes(FG);
  const Grid<D> g = make_grid<D>(s->size);
  const long long blocks = (long long)s->B * s->H * groups;
  if (blocks >= (1ll << 31)) return false;
  if (sum) {
    *err = cudaFuncSetAttribute(cl_scatter_kernel<D, true, GT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (*err != cudaSuccess) return true;
    cl_scatter_kernel<D, true, GT><<<(unsigned)blocks, kTileThreads, smem, stream>>>(keys, feat, pad, z, arg, g, s->H, s->F,
                                                                               s->N, FG, groups, tw, LP);
  } else {
    *err = cudaFuncSetAttribute(cl_scatter_kernel<D, false, GT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (*err != cudaSuccess) return true;
    cl_scatter_kernel<D, false, GT><<<(unsigned)blocks, kTileThreads, smem, stream>>>(keys, feat, pad, z, arg, g, s->H, s->F,
                                                                                s->N, FG, groups, tw, LP);
  }
  *err = cudaGetLastError();
  return true;
}

}  // namespace ctb
