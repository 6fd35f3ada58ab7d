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
// The chunks are double-buffered: while B runs on chunk c, the global loads of chunk c+1 are in flight into
// registers and are staged into the other buffer afterwards (one barrier per chunk).
// Arithmetic and results are identical to tile_scatter_kernel (same products, integer max / min / add); the sum
// adds the RAW limb words (rounding constant included) and a per-cell contribution count removes the constant at
// read-out, which saves two integer instructions per element.
//
// Paired mode (PAR: groups of 9..16 channels, last grid axis even).  Two points share a warp step, 16 lanes each.
// The cell pitch is 16 words, so a cell lives in banks 0-15 (even cell) or 16-31 (odd cell), and the parity of a
// cell is the parity of its last-axis coordinate.  Each point has 2^(d-1) even and 2^(d-1) odd corners: the first
// half-warp walks its point's even corners first, the second half-warp its odd corners first, so at every step the
// two half-warps are in opposite bank halves -- no bank conflict on any atomic (ncu before: 40 % of the shared
// wavefronts of the c3d sum were conflicts, r01_ncu_full_summary.csv).
#pragma once
#include "ctb_tile.cuh"

namespace ctb {

constexpr int kClThreads = 512; // CTA size of the channel-lane kernel (its staging index math is compile time)
constexpr int kClChunk = 128;   // points per chunk (power of two)

__host__ __device__ inline int cl_stage_words(int FG) { return (FG * (kClChunk + 1) + 3) & ~3; }
// words of one staging buffer: features [FG][chunk+1], corner offsets and weights [chunk][2^d], flips [chunk]
__host__ __device__ inline int cl_buffer_words(int FG, int dim) {
  return cl_stage_words(FG) + kClChunk * (1 << dim) * 2 + kClChunk;
}
inline size_t cl_extra_bytes(int FG, int dim, int cells) {          // + per-cell contribution counts (sum)
  return (size_t)cl_buffer_words(FG, dim) * 4 * 2 + 16 + (size_t)((cells + 3) & ~3) * 4;
}

template <int D, bool SUM, bool PAR, typename GT>
__global__ void __launch_bounds__(kClThreads, 2)
cl_scatter_kernel(const float* __restrict__ keys, const float* __restrict__ feat, const float* __restrict__ pad,
                  GT* __restrict__ z, int* __restrict__ arg, Grid<D> g, int H, int F, int N, int FG, int groups,
                  int tw, int LP) {
  constexpr int S = 1 << D;
  constexpr int HALF = S / 2;            // corner bit of the last (fastest) grid axis
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int fgp = PAR ? 16 : cl_pitch(FG, TILE_CL);
  const bool want_arg = !SUM && arg != nullptr;
  float* tval = (float*)smem_raw;                                    // max: value bits   | sum: low limb
  int* targ = (int*)(tval + tw);                                     // max: arg (if any) | sum: high limb
  float* stage = (float*)(targ + ((SUM || want_arg) ? tw : 0));      // two staging buffers (double buffering)
  const int bw = cl_buffer_words(FG, D);
  int* counter = (int*)(stage + 2 * bw);                             // [1] max|v| bits, [2] non-finite
  int* ccnt = counter + 4;                                           // sum: contributions per cell

  asm volatile("griddepcontrol.launch_dependents;");   // see tile_scatter_kernel
  const int f0 = (blockIdx.x % groups) * FG;
  const int unit = blockIdx.x / groups;
  const int fg = min(FG, F - f0);
  {
    float4* t4 = reinterpret_cast<float4*>(tval);
    int4* a4 = reinterpret_cast<int4*>(targ);
    for (int i = threadIdx.x; i < (tw >> 2); i += kClThreads) {
      t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (SUM) a4[i] = make_int4(0, 0, 0, 0);
      else if (want_arg) a4[i] = make_int4(-1, -1, -1, -1);
    }
    if (SUM) for (int i = threadIdx.x; i < g.C; i += kClThreads) ccnt[i] = 0;
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
    for (int n = threadIdx.x; n < N; n += kClThreads) {
      const float pd = pu ? __ldg(pu + n) : 1.0f;
      for (int f = 0; f < fg; f += 8) {       // eight independent loads in flight per thread
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = f + q < fg ? __ldg(fu + (size_t)(f + q) * N + n) : 0.0f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float av = fabsf(v[q] * pd);
          bad |= !(av <= 3.0e38f);
          m = fmaxf(m, av);
        }
      }
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
    if (cnt_bits <= kLimbMaxCountBits) limb_bits = kLimbBits;
    if (fixed_point && M > 0.0f) {
      int k = limb_bits > 0 ? fixed_split_exponent(M) : 62 - (ilogbf(M) + 1) - (cnt_bits + 1);
      k = k > 120 ? 120 : k;
      scale = ldexpf(1.0f, k);
      inv_scale = ldexpf(1.0f, -k);
    }
  }

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (PAR) LP = 16;                          // (compile-time constant in the paired kernels)
  const int lp_sh = PAR ? 4 : __ffs(LP) - 1; // LP is a power of two
  const int ch = lane & (LP - 1);           // my channel inside the group
  const int sub = lane >> lp_sh;            // which of the 32 / LP points of a warp step
  const bool ch_ok = ch < fg;
  const unsigned lo_base = smem_u32(tval) + (unsigned)ch * 4u, hi_off = (unsigned)tw * 4u;

  // Chunk pipeline: while the warps work through chunk c (shared atomics), the features and keys of chunk c+1 are
  // already in flight into registers; they are staged into the other buffer afterwards -- one barrier per chunk
  // and no exposed global-load latency.
  constexpr int XR = PAR ? (16 * kClChunk / kClThreads) : (32 * kClChunk / kClThreads);   // feature regs / thread
  const int nchunks = (N + kClChunk - 1) / kClChunk;
  float nx[XR], nk[D], npd = 1.0f;          // (the chunk index j of a thread is the same for all its XR rows)
  auto fetch = [&](int c) {                     // global -> registers
    const int c0 = c * kClChunk, pcn = min(kClChunk, N - c0);
    // thread -> (row f = tid / chunk + r * rows_per_sweep, point j = tid % chunk): j is the same for all r
    const int j = threadIdx.x % kClChunk, fb = threadIdx.x / kClChunk;
    const float* src = fu + (size_t)fb * N + c0 + j;
    const bool jok = j < pcn;
#pragma unroll
    for (int r = 0; r < XR; ++r)
      nx[r] = (jok && fb + r * (kClThreads / kClChunk) < fg) ? __ldg(src + (size_t)r * (kClThreads / kClChunk) * N) : 0.0f;
    if (pu) npd = jok ? __ldg(pu + c0 + j) : 0.0f;
    if ((int)threadIdx.x < pcn) {
#pragma unroll
      for (int a2 = 0; a2 < D; ++a2) nk[a2] = __ldg(ku + (size_t)a2 * N + c0 + threadIdx.x);
    }
  };
  auto stage_chunk = [&](int c) {               // registers -> staging buffer c & 1
    const int pcn = min(kClChunk, N - c * kClChunk);
    float* xs = stage + (c & 1) * bw;
    int* pa = (int*)(xs + cl_stage_words(FG));
    float* pw = (float*)(pa + kClChunk * S);
    int* pfl = (int*)(pw + kClChunk * S);
#pragma unroll
    for (int r = 0; r < XR; ++r) {
      const int i = threadIdx.x + r * kClThreads;
      const int f = i / kClChunk, j = i % kClChunk;
      if (f < fg) xs[f * (kClChunk + 1) + j] = pu ? CTB_FMUL(nx[r], npd) : nx[r];
    }
    if ((int)threadIdx.x < pcn) {
      const Pos<D> p = point_pos_from_values<D>(nk, g);
      // PAR: slots 0..HALF-1 hold the corners in even cells, HALF..S-1 those in odd cells
      const int fl = PAR ? ((p.base & 1) ? HALF : 0) : 0;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        pa[threadIdx.x * S + (s ^ fl)] = (p.base + corner_offset<D>(g, s)) * fgp;
        pw[threadIdx.x * S + (s ^ fl)] = corner_weight<D>(p, s);
        if (SUM) atomicAdd(ccnt + p.base + corner_offset<D>(g, s), 1);
      }
      if (PAR) pfl[threadIdx.x] = fl;
    }
  };

#pragma unroll 1
  for (int pass = 0; pass < (want_arg ? 2 : 1); ++pass) {
    fetch(0);
    stage_chunk(0);
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      const int c0 = c * kClChunk;
      const int pcn = min(kClChunk, N - c0);
      if (c + 1 < nchunks) fetch(c + 1);
      const float* xs = stage + (c & 1) * bw;
      const int* pa = (const int*)(xs + cl_stage_words(FG));
      const float* pw = (const float*)(pa + kClChunk * S);
      const int* pfl = (const int*)(pw + kClChunk * S);
      // lanes = channels.  A warp step takes the points j, j + LP, j + 2 LP .. so that the staged feature reads of
      // its point groups fall into different banks.
      for (int it = warp; it < (kClChunk / 32) * LP; it += kClThreads / 32) {
        const int j = ((it >> lp_sh) << 5) + (it & (LP - 1)) + (sub << lp_sh);
        if (!ch_ok || j >= pcn) continue;
        const float x = xs[ch * (kClChunk + 1) + j];
        int a[S];
        float w[S];
        const int hs = PAR ? sub : 0;      // second half-warp starts with the odd-cell corners
        if constexpr (S == 8) {
          const int4 a0 = reinterpret_cast<const int4*>(pa)[j * 2 + hs], a1 = reinterpret_cast<const int4*>(pa)[j * 2 + (hs ^ 1)];
          const float4 w0 = reinterpret_cast<const float4*>(pw)[j * 2 + hs], w1 = reinterpret_cast<const float4*>(pw)[j * 2 + (hs ^ 1)];
          a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
          w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
        } else {
          const int2 a0 = reinterpret_cast<const int2*>(pa)[j * 2 + hs], a1 = reinterpret_cast<const int2*>(pa)[j * 2 + (hs ^ 1)];
          const float2 w0 = reinterpret_cast<const float2*>(pw)[j * 2 + hs], w1 = reinterpret_cast<const float2*>(pw)[j * 2 + (hs ^ 1)];
          a[0] = a0.x; a[1] = a0.y; a[2] = a1.x; a[3] = a1.y;
          w[0] = w0.x; w[1] = w0.y; w[2] = w1.x; w[3] = w1.y;
        }
        const int n = c0 + j;
        if constexpr (SUM) {
          if (fixed_point && limb_bits > 0) {
            // raw limb words (magic + limb): the per-cell counts take the magic back out at read-out
            const float xsc = CTB_FMUL(x, scale);      // power-of-two scale: same product bits as (x * w) * scale
#pragma unroll
            for (int s = 0; s < S; s += 2) {
              unsigned long long u, r;
              fixed_split2_raw(xsc, w[s], w[s + 1], u, r);
              const unsigned a0 = lo_base + ((unsigned)a[s] << 2), a1 = lo_base + ((unsigned)a[s + 1] << 2);
              red_shared_add_u32(a0, (unsigned)r);
              red_shared_add_u32(a0 + hi_off, (unsigned)u);
              red_shared_add_u32(a1, (unsigned)(r >> 32));
              red_shared_add_u32(a1 + hi_off, (unsigned)(u >> 32));
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
            const int sx = PAR ? (pfl[j] ^ (hs ? HALF : 0)) : 0;      // slot -> corner number
#pragma unroll
            for (int s = 0; s < S; ++s)
              if (hit[s]) atomicMin((unsigned*)targ + a[s] + ch, (unsigned)((s ^ sx) * N + n));
          }
        }
      }
      if (c + 1 < nchunks) stage_chunk(c + 1);
      __syncthreads();
    }
  }

  GT* zu = z + ((size_t)unit * F + f0) * g.C;
  int* au = want_arg ? arg + ((size_t)unit * F + f0) * g.C : nullptr;
  if constexpr (PAR) {
    // pitch 16: lanes along the cells read four channels at a time (16-byte shared loads), stores stay coalesced
    const int C = g.C;
    for (int i = threadIdx.x; i < C * 4; i += kClThreads) {
      const int r = i % C, qd = i / C;
      if (qd * 4 >= fg) continue;
      const float4 v4 = reinterpret_cast<const float4*>(tval)[r * 4 + qd];
      const int4 h4 = (SUM || want_arg) ? reinterpret_cast<const int4*>(targ)[r * 4 + qd] : make_int4(0, 0, 0, 0);
      const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
      const int hh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int f = qd * 4 + k;
        if (f >= fg) break;
        float v1 = vv[k];
        if (SUM && fixed_point) {
          const unsigned back = (unsigned)ccnt[r] * (unsigned)kRoundMagicBits;
          if (limb_bits > 0) v1 = fixed_join((int)((unsigned)__float_as_int(v1) - back), (int)((unsigned)hh[k] - back), inv_scale);
          else v1 = __ll2float_rn(((long long)hh[k] << 32) | (long long)(unsigned)__float_as_int(v1)) * inv_scale;
        }
        grid_store(zu + (size_t)f * C + r, v1);
        if (want_arg) __stcs(au + (size_t)f * C + r, hh[k]);
      }
    }
    return;
  }
  for_each_plane_element(fg, g.C, [&](int f, int r) {
    float v1 = tval[r * fgp + f];
    if (SUM && fixed_point) {
      const int hi = targ[r * fgp + f];
      const unsigned back = (unsigned)ccnt[r] * (unsigned)kRoundMagicBits;
      if (limb_bits > 0) v1 = fixed_join((int)((unsigned)__float_as_int(v1) - back), (int)((unsigned)hi - back), inv_scale);
      else v1 = __ll2float_rn(((long long)hi << 32) | (long long)(unsigned)__float_as_int(v1)) * inv_scale;
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
  // re-fit the channel group with the staging buffers included; 9..16 channels on an even last axis run paired
  // (cell pitch 16), everything else with the odd pitch FG | 1
  static const bool no_pairing = getenv("CTB_CL_NO_PAIRING") != nullptr;
  const bool even_last = (s->size[s->dim - 1] % 2) == 0;
  const int cells = s->size[0] * (s->dim == 2 ? s->size[1] : s->size[1] * s->size[2]);
  auto lanes = [](int fg) { int lp = 1; while (lp < fg) lp <<= 1; return lp; };
  auto paired = [&](int fg) { return lanes(fg) == 16 && even_last && !no_pairing; };
  auto tile_words = [&](int fg) { return paired(fg) ? cells * 16 : tile_array_words(cells, fg, TILE_CL); };
  auto bytes = [&](int fg) { return (size_t)tile_words(fg) * 4 * arrays + cl_extra_bytes(fg, s->dim, cells); };
  int FG = c.FG > 32 ? 32 : c.FG;
  while (FG > 1 && bytes(FG) > (size_t)kTileSmemTwoCtas) --FG;
  if (bytes(FG) > (size_t)kTileSmemTwoCtas || FG < 4) return false;
  int groups = (s->F + FG - 1) / FG;
  FG = (s->F + groups - 1) / groups;
  groups = (s->F + FG - 1) / FG;
  const int LP = lanes(FG);
  const bool par = paired(FG);
  const int tw = tile_words(FG);
  const size_t smem = bytes(FG);
  if (smem > (size_t)kTileSmemTwoCtas) return false;
  const Grid<D> g = make_grid<D>(s->size);
  const long long blocks = (long long)s->B * s->H * groups;
  if (blocks >= (1ll << 31)) return false;
  auto launch = [&](auto kernel) {
    *err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (*err != cudaSuccess) return;
    kernel<<<(unsigned)blocks, kClThreads, smem, stream>>>(keys, feat, pad, z, arg, g, s->H, s->F, s->N, FG, groups, tw, LP);
  };
  if (sum) {
    if (par) launch(cl_scatter_kernel<D, true, true, GT>);
    else launch(cl_scatter_kernel<D, true, false, GT>);
  } else {
    if (par) launch(cl_scatter_kernel<D, false, true, GT>);
    else launch(cl_scatter_kernel<D, false, false, GT>);
  }
  if (*err != cudaSuccess) return true;
  *err = cudaGetLastError();
  return true;
}

}  // namespace ctb
