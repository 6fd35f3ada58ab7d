// ctb_tile.cuh -- CTB_MODE_TILE kernels: the grid lives in CTA-owned shared-memory tiles.
//
// Measured on B200 (tools/microbench.cu, profiles/r01_microbench.txt): ATOMS.MAX.S32 on random shared
// addresses runs at ~11 lane-ops/clk/SM (= plain random LDS), shared float add is an ATOMS.CAST.SPIN loop
// at ~3, L2 atomics (REDG) reach ~0.67 and scattered 4-byte LDG ~1.0 lane-ops/clk/SM.  So both the
// scatters and the gathers resolve their 2^d random corner accesses per point ON CHIP and touch HBM only
// with coalesced tile moves:
//
//   work item = (unit (b,h), channel group, slab of R grid rows along axis 0)
//   1. (slabs > 1) compact the indices of the points whose rows fall into the slab into shared memory --
//      one cheap axis-0 test per point -- so the main loops run with full warps;
//   2. scatter: accumulate into the tile with native shared atomics, store the slab once
//      gather : load the slab (+1 halo row) once, read corners from the tile;
//   every cell of z / arg / grad_grid is written exactly once, no zero-fill pass, no L2 atomics.
//
// Two tile layouts (template LAYOUT), chosen per shape on the host:
//   PM4 / PM1  plane-major [channel][cell] -- big sparse grids (cells > entries): the tile move dominates, so it
//              is done with 16-byte (PM4) or 4-byte (PM1, odd shapes) accesses and no address arithmetic;
//   CL         channel-last [cell][fg|1]   -- small dense grids (entries >= cells): the per-entry work
//              dominates, channel offsets become immediates and the odd pitch keeps plane moves conflict-free.
//
// tile_scatter_kernel  A2+A3 Splat forward (reduce = max: pass 1 atomicMax on the int view of max(v, +0) -- the
//                      reference's zero floor makes non-positive products irrelevant --, pass 2 atomicMin(e)
//                      among the entries equal to the maximum == torch-scatter's "first strictly greater in
//                      ascending e" rule; order independent => reproducible without a sort) and the grad_grid
//                      half of A5 Slice backward (reduce = sum, shared float atomics).
// tile_gather_kernel   A4 Slice forward, the grad_keys half of A5, A6 Splat backward (+A7 folded in).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <type_traits>
#include "../../include/ctb200.h"
#include "ctb_positions.cuh"

namespace ctb {

// Developer tool: build with -DCTB_PHASE_TIMERS to accumulate, per CTA (thread 0), the clock64 cycles between the
// phase boundaries of tile_scatter_kernel into g_phase[] (read with ctb_debug_phase()).
#ifdef CTB_PHASE_TIMERS
__device__ unsigned long long g_phase[16];
#define CTB_STAMP(i)                                                         \
  do {                                                                       \
    if (threadIdx.x == 0) {                                                  \
      const long long t_now = clock64();                                     \
      atomicAdd(&g_phase[i], (unsigned long long)(t_now - t_prev));          \
      t_prev = t_now;                                                        \
    }                                                                        \
  } while (0)
#define CTB_STAMP_INIT long long t_prev = clock64()
#else
#define CTB_STAMP(i) do { } while (0)
#define CTB_STAMP_INIT do { } while (0)
#endif

#ifndef CTB_TILE_THREADS_DEFAULT
#define CTB_TILE_THREADS_DEFAULT 512
#endif
#ifndef CTB_TILE_MIN_CTAS
#define CTB_TILE_MIN_CTAS 2
#endif
constexpr int kTileThreads = CTB_TILE_THREADS_DEFAULT;   // launch bound; the actual CTA size is blockDim.x
inline int tile_threads() {
  static const int t = getenv("CTB_TILE_THREADS") ? atoi(getenv("CTB_TILE_THREADS")) : kTileThreads;
  return (t >= 128 && t <= kTileThreads && t % 32 == 0) ? t : kTileThreads;
}
constexpr int kTileSmemTwoCtas = 110 * 1024;
constexpr int kTileSmemMax = 220 * 1024;
constexpr int kTileMaxPoints = 65535;  // compacted point lists are uint16

enum TileLayout { TILE_PM4 = 0, TILE_PM1 = 1, TILE_CL = 2, TILE_CLQ = 3 };

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier: a plane slab is one contiguous byte range both in
// HBM and in a plane-major tile, so one elected thread moves a whole tile with a handful of instructions and
// the copy engine overlaps it with whatever the CTA does next (point compaction, position arithmetic).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_and_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// grid element type GT: float, or __nv_bfloat16 for the bf16 STORAGE mode (grids in bf16, all arithmetic in fp32)
__device__ __forceinline__ float grid_load(const float* p) { return __ldcs(p); }
__device__ __forceinline__ float grid_load(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void grid_store(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void grid_store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

struct TileConfig {
  int FG;      // channels per tile
  int R;       // base rows (axis 0) per slab
  int slabs;
  int layout;  // TileLayout
  int words;   // 4-byte words of one tile array
  size_t smem;
};

// channel-last pitch in words: CL odd (conflict-free scalar plane moves); CLQ = FG + 4, a multiple of 4 that is not a
// power of two, so that the 4-channel "quads" of a cell are 16-byte aligned
// ---- exact, order-independent sums: two carry-free 32-bit limbs ---------------------------------------
// A contribution v is taken as the integer q = rint(v * 2^k), |q| < 2^40, and stored as q = hi * 2^20 + lo with
// hi = rint(q / 2^20) (|hi| < 2^20) and lo = q - hi * 2^20 (|lo| <= 2^19).  A point meets a cell at most once, so with
// at most 2048 points per work item neither limb sum can leave int32: integer adds commute, the result does not
// depend on the order of the atomics.  The split is done on the FMA pipe with the 1.5 * 2^23 rounding constant
// (float -> int64 conversions run on the quarter-rate XU pipe and were the top cost of the sum kernels, ncu r01).
constexpr int kLimbBits = 20;
constexpr int kLimbMaxCountBits = 11;       // <= 2048 contributions per cell
constexpr float kRoundMagic = 12582912.0f;  // 1.5 * 2^23
constexpr int kRoundMagicBits = 0x4B400000;

__device__ __forceinline__ void fixed_split(float t, int& lo, int& hi) {
  const float u = __fmaf_rn(t, 0x1p-20f, kRoundMagic);      // magic + rint(t / 2^20), one rounding
  const float hf = __fsub_rn(u, kRoundMagic);
  const float rem = __fmaf_rn(-hf, 0x1p20f, t);             // exact
  const float r = __fadd_rn(rem, kRoundMagic);              // magic + rint(rem)
  hi = __float_as_int(u) - kRoundMagicBits;
  lo = __float_as_int(r) - kRoundMagicBits;
}
// Two contributions at once with the packed fp32 instructions of sm_100 (FMUL2 / FFMA2 / FADD2): same IEEE
// roundings per element as fixed_split, half the issue slots.  t0 = x * w0, t1 = x * w1.
__device__ __forceinline__ unsigned long long pack_f32x2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void fixed_split2(float x, float w0, float w1, int& lo0, int& hi0, int& lo1, int& hi1) {
  const unsigned long long xx = pack_f32x2(x, x), ww = pack_f32x2(w0, w1);
  const unsigned long long magic = pack_f32x2(kRoundMagic, kRoundMagic), nmagic = pack_f32x2(-kRoundMagic, -kRoundMagic);
  const unsigned long long dn = pack_f32x2(0x1p-20f, 0x1p-20f), nup = pack_f32x2(-0x1p20f, -0x1p20f);
  unsigned long long t, u, hf, rem, r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(xx), "l"(ww));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(u) : "l"(t), "l"(dn), "l"(magic));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(hf) : "l"(u), "l"(nmagic));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rem) : "l"(hf), "l"(nup), "l"(t));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(rem), "l"(magic));
  hi0 = (int)(unsigned)(u & 0xffffffffull) - kRoundMagicBits;
  hi1 = (int)(unsigned)(u >> 32) - kRoundMagicBits;
  lo0 = (int)(unsigned)(r & 0xffffffffull) - kRoundMagicBits;
  lo1 = (int)(unsigned)(r >> 32) - kRoundMagicBits;
}
// Raw variant: returns the packed float pairs u = magic + hi, r = magic + lo.  Their bit patterns are
// kRoundMagicBits + hi / + lo, so a kernel that also counts the contributions per cell can add the raw words
// (wrap-around arithmetic) and subtract count * kRoundMagicBits once at read-out -- no per-element integer op.
__device__ __forceinline__ void fixed_split2_raw(float x, float w0, float w1, unsigned long long& u, unsigned long long& r) {
  const unsigned long long xx = pack_f32x2(x, x), ww = pack_f32x2(w0, w1);
  const unsigned long long magic = pack_f32x2(kRoundMagic, kRoundMagic), nmagic = pack_f32x2(-kRoundMagic, -kRoundMagic);
  const unsigned long long dn = pack_f32x2(0x1p-20f, 0x1p-20f), nup = pack_f32x2(-0x1p20f, -0x1p20f);
  unsigned long long t, hf, rem;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(xx), "l"(ww));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(u) : "l"(t), "l"(dn), "l"(magic));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(hf) : "l"(u), "l"(nmagic));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rem) : "l"(hf), "l"(nup), "l"(t));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(rem), "l"(magic));
}
// shared-memory loads through 32-bit shared-window addresses: keeps nvcc from rebuilding the generic address
// (S2R SR_CgaCtaId + shifts) next to every predicated tile access
__device__ __forceinline__ float lds_f32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
// raw bf16 tile element widened to fp32 (a bf16 is the high half of the fp32 with the same value)
__device__ __forceinline__ float lds_bf16(unsigned addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return __uint_as_float((unsigned)v << 16);
}
__device__ __forceinline__ int lds_s32(unsigned addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// predicated shared min: one setp + one predicated RED, no branch around the (rare) winner update
__device__ __forceinline__ void red_shared_min_u32_if(bool p, unsigned addr, unsigned v) {
  asm volatile(
      "{\n"
      " .reg .pred q;\n"
      " setp.ne.u32 q, %0, 0;\n"
      " @q red.shared.min.u32 [%1], %2;\n"
      "}" ::"r"((unsigned)p), "r"(addr), "r"(v)
      : "memory");
}
__device__ __forceinline__ void red_shared_add_u32(unsigned addr, unsigned v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// scale exponent for the two-limb format: max|v| * 2^k <= 2^40 / 1.001, so that |hi| stays below 2^20
__device__ __forceinline__ int fixed_split_exponent(float M) { return 2 * kLimbBits - (ilogbf(M * 1.001f) + 1); }
__device__ __forceinline__ float fixed_join(int lo, int hi, float inv_scale) {
  return __ll2float_rn(((long long)hi << kLimbBits) + (long long)lo) * inv_scale;
}

__host__ __device__ inline int cl_pitch(int FG, int layout) { return layout == TILE_CLQ ? FG + 4 : (FG | 1); }

inline int tile_array_words(int cells, int FG, int layout) {
  const long long w = (layout == TILE_CL || layout == TILE_CLQ) ? (long long)cells * cl_pitch(FG, layout)
                                                                : (long long)cells * FG;
  return (int)((w + 3) & ~3ll);
}

// Pick layout, FG and R.  Two ways to make a unit's F x C grid fit the shared-memory budget:
//   channel groups  (whole grid, FG channels): positions are recomputed once per group  (~110 instr / point)
//   row slabs       (all channels, R rows)   : boundary rows are visited twice, points are compacted per slab and
//                                              their features are read with gaps
// Whole-grid groups win as soon as a group holds >= 3 channels (or all of them); otherwise all channels in
// balanced slabs of >= 3 rows; otherwise fewer channels in slabs.  halo = 1 for gathers (the +1 corner row).
inline bool tile_config(const ctb_shape* s, int arrays, int halo, TileConfig* out, bool allow_quad = false,
                        int cell_bytes = 0) {
  // cell_bytes: shared-memory bytes per (cell, channel) over all tile arrays; default 4 per array.  The gathers of
  // the bf16 storage mode keep the grid tile as raw bf16: 2 bytes (+ 4 for the int32 arg tile of Splat backward).
  if (cell_bytes == 0) cell_bytes = 4 * arrays;
  const int stride0 = s->dim == 2 ? s->size[1] : s->size[1] * s->size[2];
  const int W0 = s->size[0];
  const long long C = (long long)W0 * stride0;
  const int rows = W0 - halo;                  // base rows (gather) / destination rows (scatter)
  if (s->N > kTileMaxPoints) return false;
  const long long entries = (long long)s->N << s->dim;
  // channel-last only pays when the per-entry work dwarfs the tile move (coarse, dense grids); its quad-lane form
  // (CLQ: one lane = one point x 4 channels, 16-byte tile accesses) needs channel counts that are multiples of 4
  const bool dense = entries >= 8 * C;
  if (allow_quad && dense && s->F % 4 == 0 && s->F >= 8) {
    const size_t lb = (((size_t)s->N + 7) & ~(size_t)7) * 2 + 32;
    for (int FG = s->F; FG >= 8; FG -= 4) {
      const size_t b = (size_t)tile_array_words((int)C, FG, TILE_CLQ) * cell_bytes + lb;
      if (b <= (size_t)kTileSmemTwoCtas) {
        const int groups = (s->F + FG - 1) / FG;
        FG = ((s->F + groups - 1) / groups + 3) & ~3;     // balanced groups, still multiples of 4
        out->FG = FG;
        out->R = rows;
        out->slabs = 1;
        out->layout = TILE_CLQ;
        out->words = tile_array_words((int)C, FG, TILE_CLQ);
        out->smem = (size_t)out->words * cell_bytes + lb;
        return true;
      }
    }
  }
  const int layout = dense ? TILE_CL : ((stride0 % 4 == 0) ? TILE_PM4 : TILE_PM1);
  const size_t list_bytes = (((size_t)s->N + 7) & ~(size_t)7) * 2 + 32;  // uint16 list + counters + mbarrier
  auto bytes = [&](long long cells, int FG) {
    return (size_t)tile_array_words((int)cells, FG, layout) * cell_bytes + list_bytes;
  };
  auto fill = [&](int FG, int R, int slabs) {
    out->FG = FG;
    out->R = R;
    out->slabs = slabs;
    out->layout = layout;
    out->words = tile_array_words((slabs == 1 ? W0 : R + halo) * stride0, FG, layout);
    out->smem = (size_t)out->words * cell_bytes + list_bytes;
  };
  // tuning knobs for experiments (not part of the ABI): CTB_TILE_BUDGET_KB, CTB_TILE_MIN_GROUP
  static const int env_budget = getenv("CTB_TILE_BUDGET_KB") ? atoi(getenv("CTB_TILE_BUDGET_KB")) : 0;
  static const int env_min_group = getenv("CTB_TILE_MIN_GROUP") ? atoi(getenv("CTB_TILE_MIN_GROUP")) : 3;
  for (int pass = 0; pass < 2; ++pass) {
    size_t budget = pass == 0 ? kTileSmemTwoCtas : kTileSmemMax;
    if (pass == 0 && env_budget > 0) budget = (size_t)env_budget * 1024;
    // (1) whole grid, channel groups
    int FG = s->F;
    while (FG > 1 && bytes(C, FG) > budget) --FG;
    if (bytes(C, FG) <= budget && (FG >= env_min_group || FG == s->F)) {
      const int groups = (s->F + FG - 1) / FG;
      fill((s->F + groups - 1) / groups, rows, 1);
      return true;
    }
    // (2) row slabs, as many channels as give >= 3 rows per slab
    for (int fg = s->F; fg >= 1; fg = (fg + 1) / 2) {
      int R = rows;
      while (R > 1 && bytes((long long)(R + halo) * stride0, fg) > budget) --R;
      const int min_rows = rows < 3 ? rows : 3;
      if (bytes((long long)(R + halo) * stride0, fg) <= budget && R >= min_rows) {
        const int groups = (s->F + fg - 1) / fg;
        const int slabs = (rows + R - 1) / R;
        R = (rows + slabs - 1) / slabs;
        fill((s->F + groups - 1) / groups, R, (rows + R - 1) / R);
        return true;
      }
      if (fg == 1) break;
    }
  }
  return false;
}

// Shared-memory list of the points of this slab.  Returns the count; sel[i] is the point index.
// A point is selected if one of its two corner rows c0, c0+1 (scatter, BOTH_ROWS) or its base row c0 (gather)
// lies in [x0, x1).
template <int D, bool BOTH_ROWS>
__device__ __forceinline__ int compact_slab_points(const float* __restrict__ ku, int N, const Grid<D>& g, int x0, int x1,
                                                   unsigned short* sel, int* counter, int n_begin = 0) {
  // points n_begin <= n < N are scanned (a CTA that shares its work item with others passes its own range)
  if (threadIdx.x == 0) *counter = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  constexpr int kBatch = 4;   // keys of 4 rounds are fetched together: one L2 round trip instead of four
  for (int n0 = n_begin; n0 < N; n0 += (int)blockDim.x * kBatch) {
    float kx[kBatch];
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
      const int n = n0 + b * (int)blockDim.x + threadIdx.x;
      kx[b] = n < N ? __ldg(ku + n) : 0.0f;
    }
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
      const int n = n0 + b * (int)blockDim.x + threadIdx.x;
      bool take = false;
      if (n < N) {
        bool in_rng;
        float up0, dn0;
        int c0;
        axis_pos<D>(kx[b], g.scale[0], g.W[0], up0, dn0, c0, in_rng);
        take = (c0 >= x0 && c0 < x1) || (BOTH_ROWS && (c0 + 1 >= x0 && c0 + 1 < x1));
      }
      const unsigned m = __ballot_sync(0xffffffffu, take);
      int base = 0;
      if (lane == 0 && m) base = atomicAdd(counter, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (take) sel[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)n;
    }
  }
  __syncthreads();
  return *counter;
}

// Iterate the [fg][count] index space with kTileThreads threads and no per-element division:
// fn(f, r) is called for every (plane f, element r).
template <typename Fn>
__device__ __forceinline__ void for_each_plane_element(int fg, int count, Fn fn) {
  if (count <= 0) return;
  int f = threadIdx.x / count;
  int r = threadIdx.x - f * count;
  while (f < fg) {
    fn(f, r);
    r += (int)blockDim.x;
    while (r >= count) {
      r -= count;
      ++f;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
template <int D, bool SUM, int LAYOUT, typename GT>
__global__ void __launch_bounds__(kTileThreads, CTB_TILE_MIN_CTAS)
tile_scatter_kernel(const float* __restrict__ keys, const float* __restrict__ feat, const float* __restrict__ pad,
                    GT* __restrict__ z, int* __restrict__ arg, Grid<D> g, int H, int F, int N, int FG, int R,
                    int slabs, int groups, int tw) {
  constexpr int S = 1 << D;
  constexpr bool CL = LAYOUT == TILE_CL || LAYOUT == TILE_CLQ;
  constexpr bool F32 = std::is_same<GT, float>::value;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int stride0 = g.stride[0];
  const int W0 = g.W[0];
  const int tile_cells = R * stride0;
  const int cs = CL ? cl_pitch(FG, LAYOUT) : 1; // word stride between cells
  const int fs = CL ? 1 : tile_cells;           // word stride between channels
  const bool want_arg = !SUM && arg != nullptr;
  float* tval = (float*)smem_raw;                                   // max: value bits   | sum: low limb
  int* targ = (int*)(tval + tw);                                    // max: arg (if any) | sum: high limb
  unsigned short* sel = (unsigned short*)(targ + ((SUM || want_arg) ? tw : 0));
  int* counter = (int*)(sel + ((N + 7) & ~7));                      // [0] compaction, [1] max|v| bits, [2] non-finite
  const unsigned targ_base = smem_u32(targ);                        // shared-window address (predicated arg reds)

  CTB_STAMP_INIT;
  // a kernel launched behind this one with programmatic stream serialization (the key-gradient gather of Slice
  // backward, which reads nothing written here) may start filling the SMs this grid's last wave leaves idle
  asm volatile("griddepcontrol.launch_dependents;");
  int item = blockIdx.x;
  const int slab = item % slabs;
  item /= slabs;
  const int f0 = (item % groups) * FG;
  const int unit = item / groups;
  const int fg = min(FG, F - f0);
  const int x0 = slab * R, x1 = min(x0 + R, W0);
  const int cell0 = x0 * stride0;
  const int ncell = (x1 - x0) * stride0;

  {
    float4* t4 = reinterpret_cast<float4*>(tval);
    int4* a4 = reinterpret_cast<int4*>(targ);
    for (int i = threadIdx.x; i < (tw >> 2); i += (int)blockDim.x) {
      t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (SUM) a4[i] = make_int4(0, 0, 0, 0);
      else if (want_arg) a4[i] = make_int4(-1, -1, -1, -1);   // "no winner" == max unsigned
    }
    if (threadIdx.x == 0) counter[1] = counter[2] = 0;
  }
  const float* ku = keys + (size_t)unit * D * N;
  const float* pu = pad ? pad + (size_t)(unit / H) * N : nullptr;
  const float* fu = feat + ((size_t)unit * F + f0) * N;
  int cnt = N;
  CTB_STAMP(0);                                  // tile init (issue only)
  if (slabs > 1) cnt = compact_slab_points<D, true>(ku, N, g, x0, x1, sel, counter);
  else __syncthreads();
  CTB_STAMP(1);                                  // compaction (+ init drained)

  // sum: pick the fixed-point scale 2^k from M = max |feature * pad| of this item, so that cnt * M * 2^k < 2^62
  // (|w| <= 1, a point hits a cell at most once).  Any k gives the same exact integer sum, so the result does not
  // depend on k as long as nothing overflows; non-finite inputs fall back to float atomics.
  bool fixed_point = false;
  int limb_bits = 0;      // > 0: two carry-free limbs of this many bits; 0: 32-bit limbs with carry
  float scale = 1.0f, inv_scale = 1.0f;
  if constexpr (SUM) {
    float m = 0.0f;
    bool bad = false;
    // (loads in batches of four channels, two points per trip: eight independent loads in flight per thread)
#pragma unroll 2
    for (int i = threadIdx.x; i < cnt; i += (int)blockDim.x) {
      const int n = slabs > 1 ? (int)sel[i] : i;
      const float pd = pu ? __ldg(pu + n) : 1.0f;
      for (int f = 0; f < fg; f += 4) {
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = f + q < fg ? __ldg(fu + (size_t)(f + q) * N + n) : 0.0f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
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
    const int cnt_bits = 32 - __clz(cnt > 1 ? cnt - 1 : 1);   // cnt <= 2^cnt_bits
    if (cnt_bits <= kLimbMaxCountBits) limb_bits = kLimbBits;  // 40 magnitude bits: far below fp32 resolution
    if (fixed_point && M > 0.0f) {
      // carry-free limbs: see fixed_split; carried limbs: the whole sum must stay below 2^62
      int k = limb_bits > 0 ? fixed_split_exponent(M) : 62 - (ilogbf(M) + 1) - (cnt_bits + 1);
      k = k > 120 ? 120 : k;
      scale = ldexpf(1.0f, k);
      inv_scale = ldexpf(1.0f, -k);
    }
  }

  CTB_STAMP(2);                                  // sum prepass
  if (!CL && slabs > 1 && cnt <= (int)blockDim.x && fg <= 4) {
    // Sparse-grid slab (class a): at most one point per thread.  All global loads (keys, features) are issued up
    // front in one batch, and the arg pass reuses the registers of the max pass -- no second trip to L2.
    const bool has = (int)threadIdx.x < cnt;
    const int n = has ? (int)sel[threadIdx.x] : 0;
    float kv[D], ft[4];
#pragma unroll
    for (int a2 = 0; a2 < D; ++a2) kv[a2] = __ldg(ku + (size_t)a2 * N + n);
    const float pd = pu ? __ldg(pu + n) : 1.0f;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      ft[f] = f < fg ? __ldg(fu + (size_t)f * N + n) : 0.0f;
      if (pu) ft[f] = CTB_FMUL(ft[f], pd);
    }
    const Pos<D> p = point_pos_from_values<D>(kv, g);
    CTB_STAMP(3);                                // global loads of the point (thread 0's)
    const bool in0 = has && (p.c0 >= x0) && (p.c0 < x1);
    const bool in1 = has && (p.c0 + 1 >= x0) && (p.c0 + 1 < x1);
    float w[S];
    int a[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const bool ins = (s & 1) ? in1 : in0;
      const int lc = p.base + corner_offset<D>(g, ins ? s : (s ^ 1)) - cell0;
      w[s] = ins ? corner_weight<D>(p, s) : 0.0f;
      a[s] = has ? lc : 0;
    }
    if constexpr (SUM) {
      if (!has) {
        // idle lanes must not touch the tile (they would all hammer word 0)
      } else if (fixed_point && limb_bits > 0) {
#pragma unroll
        for (int f = 0; f < 4; ++f)
          if (f < fg)
#pragma unroll
            for (int s = 0; s < S; ++s)
              if (w[s] != 0.0f) {
                int qlo, qhi;
                fixed_split(CTB_FMUL(CTB_FMUL(ft[f], w[s]), scale), qlo, qhi);
                atomicAdd((int*)tval + a[s] + f * fs, qlo);
                atomicAdd(targ + a[s] + f * fs, qhi);
              }
      } else {
#pragma unroll
        for (int f = 0; f < 4; ++f)
          if (f < fg)
#pragma unroll
            for (int s = 0; s < S; ++s)
              if (w[s] != 0.0f) {
                if (fixed_point) {
                  const long long q = __float2ll_rn(CTB_FMUL(CTB_FMUL(ft[f], w[s]), scale));
                  const unsigned lo = (unsigned)q;
                  const unsigned old = atomicAdd((unsigned*)tval + a[s] + f * fs, lo);
                  const int hi = (int)(q >> 32) + ((unsigned)(old + lo) < old ? 1 : 0);
                  if (hi != 0) atomicAdd(targ + a[s] + f * fs, hi);
                } else {
                  atomicAdd(tval + a[s] + f * fs, CTB_FMUL(ft[f], w[s]));
                }
              }
      }
      __syncthreads();
    } else {
      if (has) {
#pragma unroll
        for (int f = 0; f < 4; ++f)
          if (f < fg)
#pragma unroll
            for (int s = 0; s < S; ++s)
              atomicMax((int*)tval + a[s] + f * fs, __float_as_int(fmaxf(CTB_FMUL(ft[f], w[s]), 0.0f)));
      }
      __syncthreads();
      CTB_STAMP(4);                              // max pass
      if (want_arg) {
#pragma unroll
        for (int f = 0; f < 4; ++f)
          if (f < fg && has)
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const int t = ((const int*)tval)[a[s] + f * fs];
              if ((__float_as_int(CTB_FMUL(ft[f], w[s])) == t) & (t != 0))
                atomicMin((unsigned*)targ + a[s] + f * fs, (unsigned)(s * N + n));
            }
        __syncthreads();
        CTB_STAMP(5);                            // arg pass
      }
    }
  } else if constexpr (LAYOUT == TILE_CLQ) {
    // quad lanes: a lane owns one point and 4-channel quads q0, q0 + lpp, ... of it.  Lanes of a warp then cover
    // 8 (or 16 / 32) points instead of 32, which divides the same-cell collisions of clustered clouds, and the
    // arg pass reads the 4 tile values of a quad with one 16-byte shared load.
    const int qn = fg >> 2;
    const int lsh = qn >= 4 ? 2 : (qn >= 2 ? 1 : 0);
    const int lpp = 1 << lsh;
#pragma unroll 1
    for (int pass = 0; pass < (want_arg ? 2 : 1); ++pass) {
      for (int slot = threadIdx.x; slot < (cnt << lsh); slot += (int)blockDim.x) {
        const int n = slot >> lsh, q0 = slot & (lpp - 1);
        const Pos<D> p = point_pos<D>(ku, n, N, g);
        float w[S];
        int a[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          w[s] = corner_weight<D>(p, s);
          a[s] = (p.base + corner_offset<D>(g, s)) * cs;
        }
        const float pd = pu ? __ldg(pu + n) : 1.0f;
        for (int q = q0; q < qn; q += lpp) {
          const int f = q << 2;
          float ft[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ft[j] = __ldg(fu + (size_t)(f + j) * N + n);
            if (pu) ft[j] = CTB_FMUL(ft[j], pd);
          }
          if constexpr (SUM) {
            if (fixed_point && limb_bits > 0) {
#pragma unroll
              for (int s = 0; s < S; ++s)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  int qlo, qhi;
                  fixed_split(CTB_FMUL(CTB_FMUL(ft[j], w[s]), scale), qlo, qhi);
                  atomicAdd((int*)tval + a[s] + f + j, qlo);
                  atomicAdd(targ + a[s] + f + j, qhi);
                }
            } else {
#pragma unroll
              for (int s = 0; s < S; ++s)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if (fixed_point) {
                    const long long q = __float2ll_rn(CTB_FMUL(CTB_FMUL(ft[j], w[s]), scale));
                    const unsigned lo = (unsigned)q;
                    const unsigned old = atomicAdd((unsigned*)tval + a[s] + f + j, lo);
                    const int hi = (int)(q >> 32) + ((unsigned)(old + lo) < old ? 1 : 0);
                    if (hi != 0) atomicAdd(targ + a[s] + f + j, hi);
                  } else {
                    atomicAdd(tval + a[s] + f + j, CTB_FMUL(ft[j], w[s]));
                  }
                }
            }
          } else if (pass == 0) {
#pragma unroll
            for (int s = 0; s < S; ++s)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                atomicMax((int*)tval + a[s] + f + j, __float_as_int(fmaxf(CTB_FMUL(ft[j], w[s]), 0.0f)));
          } else {
            unsigned hits = 0;
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const int4 t = *reinterpret_cast<const int4*>((const int*)tval + a[s] + f);
              hits |= ((__float_as_int(CTB_FMUL(ft[0], w[s])) == t.x) & (t.x != 0)) ? (1u << (4 * s)) : 0u;
              hits |= ((__float_as_int(CTB_FMUL(ft[1], w[s])) == t.y) & (t.y != 0)) ? (2u << (4 * s)) : 0u;
              hits |= ((__float_as_int(CTB_FMUL(ft[2], w[s])) == t.z) & (t.z != 0)) ? (4u << (4 * s)) : 0u;
              hits |= ((__float_as_int(CTB_FMUL(ft[3], w[s])) == t.w) & (t.w != 0)) ? (8u << (4 * s)) : 0u;
            }
            while (hits) {
              const int b = __ffs(hits) - 1;
              hits &= hits - 1;
              const int sb = b >> 2;   // rare path: recompute the corner address instead of indexing registers
              atomicMin((unsigned*)targ + (p.base + corner_offset<D>(g, sb)) * cs + f + (b & 3), (unsigned)(sb * N + n));
            }
          }
        }
      }
      __syncthreads();
    }
  } else
#pragma unroll 1
  for (int pass = 0; pass < (want_arg ? 2 : 1); ++pass) {
    // one point of software prefetch: keys, pad and the first kPrefetchCh features of the thread's NEXT point are
    // in flight while the current point works on the tile (ncu r01: the dependent load -> atomic chains were the
    // top stall of the class-b scatters)
    constexpr int kPrefetchCh = 4;
    int nn = 0;
    float nk[D], npd = 1.0f, nft[kPrefetchCh];
    auto prefetch = [&](int i) {
      nn = slabs > 1 ? (int)sel[i] : i;
#pragma unroll
      for (int a2 = 0; a2 < D; ++a2) nk[a2] = __ldg(ku + (size_t)a2 * N + nn);
      if (pu) npd = __ldg(pu + nn);
#pragma unroll
      for (int q = 0; q < kPrefetchCh; ++q) nft[q] = q < fg ? __ldg(fu + (size_t)q * N + nn) : 0.0f;
    };
    if ((int)threadIdx.x < cnt) prefetch(threadIdx.x);
    for (int i = threadIdx.x; i < cnt; i += (int)blockDim.x) {
      const int n = nn;
      const float pd = npd;
      float kv[D], ft0[kPrefetchCh];
#pragma unroll
      for (int a2 = 0; a2 < D; ++a2) kv[a2] = nk[a2];
#pragma unroll
      for (int q = 0; q < kPrefetchCh; ++q) ft0[q] = pu ? CTB_FMUL(nft[q], pd) : nft[q];
      if (i + (int)blockDim.x < cnt) prefetch(i + (int)blockDim.x);
      const Pos<D> p = point_pos_from_values<D>(kv, g);
      // run body(f, feature * pad) over the channels of the group: registers first, then the rest from global
      auto for_channels = [&](auto&& body) {
#pragma unroll
        for (int q = 0; q < kPrefetchCh; ++q)
          if (q < fg) body(q, ft0[q]);
#pragma unroll 2
        for (int f = kPrefetchCh; f < fg; ++f) {
          float ft = __ldg(fu + (size_t)f * N + n);
          if (pu) ft = CTB_FMUL(ft, pd);
          body(f, ft);
        }
      };
      const bool in0 = (p.c0 >= x0) && (p.c0 < x1);
      const bool in1 = (p.c0 + 1 >= x0) && (p.c0 + 1 < x1);
      // corners outside the slab are redirected to their in-slab sibling (other row) with weight 0:
      // max(.., +0) / min-e on a non-positive value are no-ops, so the inner loops stay branch-free
      float w[S];
      int a[S];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const bool ins = (s & 1) ? in1 : in0;
        const int lc = p.base + corner_offset<D>(g, ins ? s : (s ^ 1)) - cell0;
        w[s] = ins ? corner_weight<D>(p, s) : 0.0f;
        a[s] = lc * cs;
      }
      if constexpr (SUM) {
        if (fixed_point && limb_bits > 0) {
          // order-independent accumulation without carries (fixed_split)
          for_channels([&](int f, float ft) {
            const float fsc = CTB_FMUL(ft, scale);     // power-of-two scale: same bits as (ft * w) * scale
#pragma unroll
            for (int s = 0; s < S; s += 2) {
              int l0, h0, l1, h1;
              fixed_split2(fsc, w[s], w[s + 1], l0, h0, l1, h1);
              atomicAdd((int*)tval + a[s] + f * fs, l0);
              atomicAdd(targ + a[s] + f * fs, h0);
              atomicAdd((int*)tval + a[s + 1] + f * fs, l1);
              atomicAdd(targ + a[s + 1] + f * fs, h1);
            }
          });
        } else if (fixed_point) {
          // exact, order-independent accumulation: q = v * 2^k as int64, added as (lo, hi) 32-bit limbs with the
          // carry taken from the value the lo atomic returns; sum of carries == number of lo wrap-arounds
          unsigned* lo_t = (unsigned*)tval;
          for_channels([&](int f, float ft) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const long long q = __float2ll_rn(CTB_FMUL(CTB_FMUL(ft, w[s]), scale));
              const unsigned lo = (unsigned)q;
              const unsigned old = atomicAdd(lo_t + a[s] + f * fs, lo);
              const int hi = (int)(q >> 32) + ((unsigned)(old + lo) < old ? 1 : 0);
              if (hi != 0) atomicAdd(targ + a[s] + f * fs, hi);
            }
          });
        } else {
          for_channels([&](int f, float ft) {
#pragma unroll
            for (int s = 0; s < S; ++s)
              if (w[s] != 0.0f) atomicAdd(tval + a[s] + f * fs, CTB_FMUL(ft, w[s]));
          });
        }
      } else if (pass == 0) {
        for_channels([&](int f, float ft) {
#pragma unroll
          for (int s = 0; s < S; ++s)
            atomicMax((int*)tval + a[s] + f * fs, __float_as_int(fmaxf(CTB_FMUL(ft, w[s]), 0.0f)));
        });
      } else {
        for_channels([&](int f, float ft) {
          if constexpr (D == 2) {
            // four corners: tile values first (independent loads), then predicated reds for the (rare) winners
            // (measured: 16^2 F16 0.119 -> 0.109 ms; with eight corners the extra registers spill, 3-D keeps the
            // test-then-branch form below)
            int t[S];
#pragma unroll
            for (int s = 0; s < S; ++s) t[s] = ((const int*)tval)[a[s] + f * fs];
#pragma unroll
            for (int s = 0; s < S; ++s)
              red_shared_min_u32_if((__float_as_int(CTB_FMUL(ft, w[s])) == t[s]) & (t[s] != 0),
                                    targ_base + ((unsigned)(a[s] + f * fs) << 2), (unsigned)(s * N + n));
          } else {
            // winners are rare: test all corners branch-free first, take the atomic path only if one matched
            bool hit[S];
            bool any = false;
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const int t = ((const int*)tval)[a[s] + f * fs];
              hit[s] = (__float_as_int(CTB_FMUL(ft, w[s])) == t) & (t != 0);
              any |= hit[s];
            }
            if (any) {
#pragma unroll
              for (int s = 0; s < S; ++s)
                if (hit[s]) atomicMin((unsigned*)targ + a[s] + f * fs, (unsigned)(s * N + n));
            }
          }
        });
      }
    }
    __syncthreads();
  }

  // store the slab once, coalesced
  GT* zu = z + ((size_t)unit * F + f0) * g.C + cell0;
  int* au = want_arg ? arg + ((size_t)unit * F + f0) * g.C + cell0 : nullptr;
  if constexpr (LAYOUT == TILE_PM4 && !SUM && F32) {
    // the tile already holds the final bits of z (and arg): hand it to the copy engine, plane by plane
#ifdef CTB_FENCE_SPLIT
    CTB_STAMP(8);
    fence_proxy_async();
    CTB_STAMP(9);
    __syncthreads();
#else
    fence_proxy_async();
    __syncthreads();
#endif
    CTB_STAMP(6);                                // (generic-path passes end here too)
    if (threadIdx.x == 0) {
      for (int f = 0; f < fg; ++f) {
        bulk_s2g((float*)zu + (size_t)f * g.C, tval + (size_t)f * tile_cells, (uint32_t)ncell * 4u);
        if (want_arg) bulk_s2g(au + (size_t)f * g.C, targ + (size_t)f * tile_cells, (uint32_t)ncell * 4u);
      }
      bulk_commit_and_wait_read();
    }
    CTB_STAMP(7);                                // TMA store until the tile has been read
    return;
  }
  auto limbs_to_float = [&](float lo_bits, int hi) {
    if (limb_bits > 0) return fixed_join(__float_as_int(lo_bits), hi, inv_scale);
    const long long q = ((long long)hi << 32) | (long long)(unsigned)__float_as_int(lo_bits);
    return __ll2float_rn(q) * inv_scale;
  };
  if constexpr (LAYOUT == TILE_PM4 && F32) {
    for_each_plane_element(fg, ncell >> 2, [&](int f, int r) {
      float4 v4 = reinterpret_cast<const float4*>(tval + (size_t)f * tile_cells)[r];
      if (SUM && fixed_point) {
        const int4 h4 = reinterpret_cast<const int4*>(targ + (size_t)f * tile_cells)[r];
        v4.x = limbs_to_float(v4.x, h4.x);
        v4.y = limbs_to_float(v4.y, h4.y);
        v4.z = limbs_to_float(v4.z, h4.z);
        v4.w = limbs_to_float(v4.w, h4.w);
      }
      __stcs(reinterpret_cast<float4*>((float*)zu + (size_t)f * g.C) + r, v4);
      if (want_arg)
        __stcs(reinterpret_cast<int4*>(au + (size_t)f * g.C) + r,
               reinterpret_cast<const int4*>(targ + (size_t)f * tile_cells)[r]);
    });
  } else if (LAYOUT == TILE_PM4 && !F32) {
    // bf16 grid storage, plane-major: four cells per step, rounded to bf16 and stored as 8 bytes
    if constexpr (!F32) {
      for_each_plane_element(fg, ncell >> 2, [&](int f, int r) {
        float4 v4 = reinterpret_cast<const float4*>(tval + (size_t)f * tile_cells)[r];
        const int4 h4 = (SUM || want_arg) ? reinterpret_cast<const int4*>(targ + (size_t)f * tile_cells)[r]
                                          : make_int4(0, 0, 0, 0);
        if (SUM && fixed_point) {
          v4.x = limbs_to_float(v4.x, h4.x);
          v4.y = limbs_to_float(v4.y, h4.y);
          v4.z = limbs_to_float(v4.z, h4.z);
          v4.w = limbs_to_float(v4.w, h4.w);
        }
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v4.x, v4.y), hi = __floats2bfloat162_rn(v4.z, v4.w);
        uint2 packed;
        packed.x = *reinterpret_cast<const unsigned*>(&lo);
        packed.y = *reinterpret_cast<const unsigned*>(&hi);
        __stcs(reinterpret_cast<uint2*>(zu + (size_t)f * g.C) + r, packed);
        if (want_arg) __stcs(reinterpret_cast<int4*>(au + (size_t)f * g.C) + r, h4);
      });
    }
  } else {
    for_each_plane_element(fg, ncell, [&](int f, int r) {
      float v1 = tval[r * cs + f * fs];
      if (SUM && fixed_point) v1 = limbs_to_float(v1, targ[r * cs + f * fs]);
      grid_store(zu + (size_t)f * g.C + r, v1);
      if (want_arg) __stcs(au + (size_t)f * g.C + r, targ[r * cs + f * fs]);
    });
  }
}

inline bool tile_scatter_config(const ctb_shape* s, bool sum, bool want_arg, TileConfig* out) {
  // quad lanes measured faster only for the 3-D max scatter (c3d 0.43 -> 0.35 ms); 2-D and sum keep point lanes
  static const bool quad_sum = getenv("CTB_QUAD_SUM") != nullptr;
  return tile_config(s, (sum || want_arg) ? 2 : 1, 0, out, (!sum || quad_sum) && s->dim == 3);
}

template <int D, bool SUM, int LAYOUT, typename GT>
cudaError_t launch_tile_scatter(const float* keys, const float* feat, const float* pad, GT* z, int* arg,
                                const ctb_shape* s, const TileConfig& c, cudaStream_t stream) {
  const Grid<D> g = make_grid<D>(s->size);
  const int groups = (s->F + c.FG - 1) / c.FG;
  const long long blocks = (long long)s->B * s->H * groups * c.slabs;
  if (blocks >= (1ll << 31)) return cudaErrorNotSupported;
  cudaError_t e = cudaFuncSetAttribute(tile_scatter_kernel<D, SUM, LAYOUT, GT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
  if (e != cudaSuccess) return e;
  tile_scatter_kernel<D, SUM, LAYOUT, GT><<<(unsigned)blocks, tile_threads(), c.smem, stream>>>(
      keys, feat, pad, z, arg, g, s->H, s->F, s->N, c.FG, c.R, c.slabs, groups, c.words);
  return cudaGetLastError();
}

inline int effective_layout(int layout, const void* p1, const void* p2, bool raw16 = false, int stride0 = 0) {
  if (layout == TILE_PM4 && (((reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(p2)) & 15) != 0))
    return TILE_PM1;
  if (layout == TILE_PM4 && raw16 && (stride0 % 8) != 0) return TILE_PM1;   // bf16 rows must be 16-byte multiples
  return layout;
}

template <int D, typename GT>
cudaError_t tile_scatter(const float* keys, const float* feat, const float* pad, GT* z, int* arg,
                         const ctb_shape* s, bool sum, cudaStream_t stream) {
  TileConfig c;
  if (!tile_scatter_config(s, sum, arg != nullptr, &c)) return cudaErrorNotSupported;
  switch (effective_layout(c.layout, z, arg)) {
    case TILE_PM4:
      return sum ? launch_tile_scatter<D, true, TILE_PM4, GT>(keys, feat, pad, z, arg, s, c, stream)
                 : launch_tile_scatter<D, false, TILE_PM4, GT>(keys, feat, pad, z, arg, s, c, stream);
    case TILE_PM1:
      return sum ? launch_tile_scatter<D, true, TILE_PM1, GT>(keys, feat, pad, z, arg, s, c, stream)
                 : launch_tile_scatter<D, false, TILE_PM1, GT>(keys, feat, pad, z, arg, s, c, stream);
    case TILE_CLQ:
      return sum ? launch_tile_scatter<D, true, TILE_CLQ, GT>(keys, feat, pad, z, arg, s, c, stream)
                 : launch_tile_scatter<D, false, TILE_CLQ, GT>(keys, feat, pad, z, arg, s, c, stream);
    default:
      return sum ? launch_tile_scatter<D, true, TILE_CL, GT>(keys, feat, pad, z, arg, s, c, stream)
                 : launch_tile_scatter<D, false, TILE_CL, GT>(keys, feat, pad, z, arg, s, c, stream);
  }
}

// ------------------------------------------------------------------------------------------------------
enum GatherMode { GATHER_SLICE_FWD = 0, GATHER_SLICE_BWD_KEYS = 1, GATHER_SPLAT_BWD = 2 };

// One CTA = (unit, slab).  Loops over channel groups; every point is resolved in the single slab that holds
// its base row, so grad_keys needs no cross-CTA reduction.
template <int D, int MODE, int LAYOUT, typename GT>
__global__ void __launch_bounds__(kTileThreads, CTB_TILE_MIN_CTAS)
tile_gather_kernel(const float* __restrict__ keys, const GT* __restrict__ t1, const int* __restrict__ t2,
                   const float* __restrict__ in, const float* __restrict__ pad, float* __restrict__ out,
                   float* __restrict__ grad_keys, Grid<D> g, int H, int F, int N, int FG, int R, int slabs, int tw,
                   int gsplit, int psplit) {
  // psplit CTAs share the POINTS of one (unit, slab, group set): used when B * H is too small to fill the GPU
  // (every CTA loads the tile, L2 serves the repeats).
  // gsplit CTAs share the channel groups of one (unit, slab): Slice forward gives every group its own CTA (more,
  // shorter CTAs: better balance over the SMs and tile loads that overlap other CTAs' gathers); the backward modes
  // keep gsplit == 1 because one thread accumulates grad_keys over the groups in a fixed order.
  constexpr int S = 1 << D;
  constexpr bool CL = LAYOUT == TILE_CL || LAYOUT == TILE_CLQ;
  // plane-major tiles arrive as raw copies by the copy engine: fp32 grids as they are, bf16 grids (RAW16) stay bf16
  // in shared memory and are widened when a corner is read (the host only picks PM4 for bf16 if rows are 16-byte
  // multiples, see effective_layout)
  constexpr bool RAW16 = LAYOUT == TILE_PM4 && std::is_same<GT, __nv_bfloat16>::value;
  constexpr bool TMA = LAYOUT == TILE_PM4;
  constexpr unsigned ESH = RAW16 ? 1u : 2u;       // log2 of the tile element size of s1
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int stride0 = g.stride[0];
  const int W0 = g.W[0];
  const int tile_cells = (slabs == 1 ? W0 : R + 1) * stride0;
  const int cs = CL ? cl_pitch(FG, LAYOUT) : 1;
  const int fs = CL ? 1 : tile_cells;
  float* s1 = (float*)smem_raw;
  int* s2 = (int*)((unsigned char*)s1 + ((size_t)tw << ESH));       // (SPLAT_BWD: arg) behind the tw elements of s1
  const unsigned s1_base = smem_u32(s1), s12 = (unsigned)tw << ESH; // shared-window address of s1; byte distance to s2
  unsigned short* sel = (unsigned short*)(s2 + (MODE == GATHER_SPLAT_BWD ? tw : 0));
  int* counter = (int*)(sel + ((N + 7) & ~7));
  uint64_t* bar = (uint64_t*)(counter + 4);     // 16-byte aligned: the list is padded to 8 entries

  const int gi = blockIdx.x % gsplit;
  const int pi = (blockIdx.x / gsplit) % psplit;
  const int slab = (blockIdx.x / (gsplit * psplit)) % slabs;
  const int unit = blockIdx.x / (gsplit * psplit * slabs);
  const int x0 = slab * R;
  const int x1 = min(x0 + R, W0 - 1);           // base rows [x0, x1)
  const int xe = min(x1 + 1, W0);               // tile rows [x0, xe)
  const int cell0 = x0 * stride0;
  const int ncell = (xe - x0) * stride0;
  const float* ku = keys + (size_t)unit * D * N;
  const float* pu = pad ? pad + (size_t)(unit / H) * N : nullptr;

  if constexpr (TMA) {
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
  }
  // my share of the unit's points (by point index, so that the CTAs of an item agree on the partition)
  const int per_n = (N + psplit - 1) / psplit;
  const int n_lo = min(N, pi * per_n), n_hi = min(N, pi * per_n + per_n);
  int cnt = n_hi - n_lo;      // slabs == 1: points n_lo + i; slabs > 1: the compacted list of my range
  if (!TMA && slabs > 1) cnt = compact_slab_points<D, false>(ku, n_hi, g, x0, x1, sel, counter, n_lo);

  uint32_t parity = 0;
  for (int f0 = gi * FG; f0 < F; f0 += FG * gsplit) {
    const int fg = min(FG, F - f0);
    const bool first = f0 == gi * FG;
    if (!first) __syncthreads();                 // previous group's readers are done with the tile
    const GT* g1 = t1 + ((size_t)unit * F + f0) * g.C + cell0;
    const int* g2 = MODE == GATHER_SPLAT_BWD ? t2 + ((size_t)unit * F + f0) * g.C + cell0 : nullptr;
    if constexpr (TMA) {
      // one thread hands the whole tile to the copy engine; the CTA compacts its points meanwhile
      if (threadIdx.x == 0) {
        const uint32_t plane_bytes = (uint32_t)ncell << ESH, arg_bytes = (uint32_t)ncell * 4u;
        mbar_expect_tx(bar, (plane_bytes + (MODE == GATHER_SPLAT_BWD ? arg_bytes : 0u)) * fg);
        for (int f = 0; f < fg; ++f) {
          bulk_g2s((unsigned char*)s1 + (((size_t)f * tile_cells) << ESH), g1 + (size_t)f * g.C, plane_bytes, bar);
          if constexpr (MODE == GATHER_SPLAT_BWD)
            bulk_g2s(s2 + (size_t)f * tile_cells, g2 + (size_t)f * g.C, arg_bytes, bar);
        }
      }
      if (first && slabs > 1) cnt = compact_slab_points<D, false>(ku, n_hi, g, x0, x1, sel, counter, n_lo);
      mbar_wait(bar, parity);
      parity ^= 1u;
    } else {
      for_each_plane_element(fg, ncell, [&](int f, int r) {
        s1[r * cs + f * fs] = grid_load(g1 + (size_t)f * g.C + r);
        if constexpr (MODE == GATHER_SPLAT_BWD) s2[r * cs + f * fs] = __ldcs(g2 + (size_t)f * g.C + r);
      });
      __syncthreads();
    }
    if constexpr (LAYOUT == TILE_CLQ) {
      // quad lanes (see tile_scatter_kernel): lane = (point, 4-channel quads q0, q0 + lpp, ...); 16-byte tile reads
      const int qn = fg >> 2;
      const int lsh = qn >= 4 ? 2 : (qn >= 2 ? 1 : 0);
      const int lpp = 1 << lsh;
      // Slice forward: keys and pad of the thread's next slot are fetched one slot ahead
      constexpr bool kAheadQ = MODE == GATHER_SLICE_FWD;
      float nk[D], npd = 1.0f;
      auto prefetch = [&](int slot) {
        const int n = slot >> lsh;
#pragma unroll
        for (int a2 = 0; a2 < D; ++a2) nk[a2] = __ldg(ku + (size_t)a2 * N + n);
        if (pu) npd = __ldg(pu + n);
      };
      const int s_lo = n_lo << lsh, s_hi = n_hi << lsh;                                          // my slots
      if (kAheadQ && s_lo + (int)threadIdx.x < s_hi) prefetch(s_lo + threadIdx.x);
#pragma unroll 1
      for (int slot = s_lo + threadIdx.x; slot < s_hi; slot += (int)blockDim.x) {
        const unsigned active = __activemask();
        const int n = slot >> lsh, q0 = slot & (lpp - 1);
        if (!kAheadQ) prefetch(slot);
        const float pd = npd;
        float kv[D];
#pragma unroll
        for (int a2 = 0; a2 < D; ++a2) kv[a2] = nk[a2];
        if (kAheadQ && slot + (int)blockDim.x < s_hi) prefetch(slot + (int)blockDim.x);
        const Pos<D> p = point_pos_from_values<D>(kv, g);
        float w[S], gw[S];
        int a[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          w[s] = corner_weight<D>(p, s);
          a[s] = (p.base + corner_offset<D>(g, s) - cell0) * cs;
          gw[s] = 0.0f;
        }
        for (int q = q0; q < qn; q += lpp) {
          const int f = q << 2;
          const size_t po = ((size_t)unit * F + f0 + f) * N + n;
          if constexpr (MODE == GATHER_SLICE_FWD) {
            float4 acc;
            {
              const float4 t = *reinterpret_cast<const float4*>(s1 + a[0] + f);
              acc = make_float4(CTB_FMUL(t.x, w[0]), CTB_FMUL(t.y, w[0]), CTB_FMUL(t.z, w[0]), CTB_FMUL(t.w, w[0]));
            }
#pragma unroll
            for (int s = 1; s < S; ++s) {
              const float4 t = *reinterpret_cast<const float4*>(s1 + a[s] + f);
              acc.x = fmaf(t.x, w[s], acc.x);
              acc.y = fmaf(t.y, w[s], acc.y);
              acc.z = fmaf(t.z, w[s], acc.z);
              acc.w = fmaf(t.w, w[s], acc.w);
            }
            if (pu) {
              acc.x = CTB_FMUL(acc.x, pd);
              acc.y = CTB_FMUL(acc.y, pd);
              acc.z = CTB_FMUL(acc.z, pd);
              acc.w = CTB_FMUL(acc.w, pd);
            }
            out[po] = acc.x;
            out[po + (size_t)N] = acc.y;
            out[po + (size_t)2 * N] = acc.z;
            out[po + (size_t)3 * N] = acc.w;
          } else if constexpr (MODE == GATHER_SLICE_BWD_KEYS) {
            float go[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              go[j] = __ldg(in + po + (size_t)j * N);
              if (pu) go[j] *= pd;
            }
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const float4 t = *reinterpret_cast<const float4*>(s1 + a[s] + f);
              gw[s] = fmaf(t.x, go[0], gw[s]);
              gw[s] = fmaf(t.y, go[1], gw[s]);
              gw[s] = fmaf(t.z, go[2], gw[s]);
              gw[s] = fmaf(t.w, go[3], gw[s]);
            }
          } else {
            float ft[4], gf[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              ft[j] = __ldg(in + po + (size_t)j * N);
              if (pu) ft[j] *= pd;
            }
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const int4 av = *reinterpret_cast<const int4*>(s2 + a[s] + f);
              const float4 t = *reinterpret_cast<const float4*>(s1 + a[s] + f);
              const int e = s * N + n;
              const float g0 = av.x == e ? t.x : 0.0f, g1 = av.y == e ? t.y : 0.0f;
              const float g2 = av.z == e ? t.z : 0.0f, g3 = av.w == e ? t.w : 0.0f;
              gf[0] = fmaf(g0, w[s], gf[0]);
              gf[1] = fmaf(g1, w[s], gf[1]);
              gf[2] = fmaf(g2, w[s], gf[2]);
              gf[3] = fmaf(g3, w[s], gf[3]);
              gw[s] = fmaf(g0, ft[0], gw[s]);
              gw[s] = fmaf(g1, ft[1], gw[s]);
              gw[s] = fmaf(g2, ft[2], gw[s]);
              gw[s] = fmaf(g3, ft[3], gw[s]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) out[po + (size_t)j * N] = pu ? gf[j] * pd : gf[j];
          }
        }
        if constexpr (MODE != GATHER_SLICE_FWD) {
          // fold the partial weight gradients of the point's lanes, one lane writes grad_keys
          for (int o = 1; o < lpp; o <<= 1)
#pragma unroll
            for (int s = 0; s < S; ++s) gw[s] += __shfl_xor_sync(active, gw[s], o);
          if (q0 == 0) {
            float part[D];
            weight_grad_to_key_grad<D>(p, gw, part);
#pragma unroll
            for (int a2 = 0; a2 < D; ++a2) {
              float* gp = grad_keys + ((size_t)unit * D + a2) * N + n;
              *gp = f0 == 0 ? part[a2] : *gp + part[a2];
            }
          }
        }
      }
    } else
    {
    // one point of software prefetch for the keys and pad of the thread's next point
    int nn = 0;
    float nk[D], npd = 1.0f;
    auto prefetch = [&](int i) {
      nn = slabs > 1 ? (int)sel[i] : i;
#pragma unroll
      for (int a2 = 0; a2 < D; ++a2) nk[a2] = __ldg(ku + (size_t)a2 * N + nn);
      if (pu) npd = __ldg(pu + nn);
    };
    // (only where it fits the 64-register budget of two 512-thread CTAs per SM; the backward modes would spill)
    constexpr bool kAhead = MODE == GATHER_SLICE_FWD;
    const int p_lo = slabs > 1 ? 0 : n_lo, p_hi = slabs > 1 ? cnt : n_hi;                         // my points
    if (kAhead && p_lo + (int)threadIdx.x < p_hi) prefetch(p_lo + threadIdx.x);
#pragma unroll 1
    for (int i = p_lo + threadIdx.x; i < p_hi; i += (int)blockDim.x) {
      if (!kAhead) prefetch(i);
      const int n = nn;
      const float pd = npd;
      float kv[D];
#pragma unroll
      for (int a2 = 0; a2 < D; ++a2) kv[a2] = nk[a2];
      if (kAhead && i + (int)blockDim.x < p_hi) prefetch(i + (int)blockDim.x);
      const Pos<D> p = point_pos_from_values<D>(kv, g);
      float w[S];
      int a[S];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        w[s] = corner_weight<D>(p, s);
        a[s] = (p.base + corner_offset<D>(g, s) - cell0) * cs;
      }
      float gw[S];
#pragma unroll
      for (int s = 0; s < S; ++s) gw[s] = 0.0f;
      const size_t po = ((size_t)unit * F + f0) * N + n;
      // byte addresses of the point's corner cells in the shared window; channel f adds f * fstep
      unsigned ab[S];
#pragma unroll
      for (int s = 0; s < S; ++s) ab[s] = s1_base + ((unsigned)a[s] << ESH);
      const unsigned fstep = (unsigned)fs << ESH;
      auto tile_val = [&](unsigned addr) {
        if constexpr (RAW16) return lds_bf16(addr);
        else return lds_f32(addr);
      };
      if constexpr (MODE == GATHER_SLICE_FWD) {
#pragma unroll 4
        for (int f = 0; f < fg; ++f) {
          const unsigned off = (unsigned)f * fstep;
          float acc = CTB_FMUL(tile_val(ab[0] + off), w[0]);
#pragma unroll
          for (int s = 1; s < S; ++s) acc = fmaf(tile_val(ab[s] + off), w[s], acc);
          if (pu) acc = CTB_FMUL(acc, pd);
          out[po + (size_t)f * N] = acc;
        }
      } else if constexpr (MODE == GATHER_SLICE_BWD_KEYS) {
        // the upstream gradients run one batch of four channels ahead of the shared-memory work
        float gn[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) gn[q] = q < fg ? __ldg(in + po + (size_t)q * N) : 0.0f;
#pragma unroll 1
        for (int f = 0; f < fg; f += 4) {
          float gc[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) gc[q] = pu ? gn[q] * pd : gn[q];
#pragma unroll
          for (int q = 0; q < 4; ++q) gn[q] = f + 4 + q < fg ? __ldg(in + po + (size_t)(f + 4 + q) * N) : 0.0f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (f + q < fg) {
              const unsigned off = (unsigned)(f + q) * fstep;
#pragma unroll
              for (int s = 0; s < S; ++s) gw[s] = fmaf(tile_val(ab[s] + off), gc[q], gw[s]);
            }
          }
        }
      } else {
        // features run one batch of two channels ahead of the shared-memory work
        float fn[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) fn[q] = q < fg ? __ldg(in + po + (size_t)q * N) : 0.0f;
#pragma unroll 1
        for (int f = 0; f < fg; f += 2) {
          float fc[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) fc[q] = pu ? fn[q] * pd : fn[q];
#pragma unroll
          for (int q = 0; q < 2; ++q) fn[q] = f + 2 + q < fg ? __ldg(in + po + (size_t)(f + 2 + q) * N) : 0.0f;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (f + q < fg) {
              float gf = 0.0f;
              const unsigned off = (unsigned)(f + q) * fstep;
#pragma unroll
              for (int s = 0; s < S; ++s) {
                const unsigned ad = ab[s] + off;
                // arg is an int32 tile behind s1: same element index, 4-byte elements
                const unsigned arg_ad = RAW16 ? s1_base + s12 + ((ad - s1_base) << 1) : ad + s12;
                const bool win = lds_s32(arg_ad) == s * N + n;
                float gz = 0.0f;
                if (win) gz = tile_val(ad);
                gf = fmaf(gz, w[s], gf);
                gw[s] = fmaf(gz, fc[q], gw[s]);
              }
              if (pu) gf *= pd;
              out[po + (size_t)(f + q) * N] = gf;
            }
          }
        }
      }
      if constexpr (MODE != GATHER_SLICE_FWD) {
        // the same thread owns point n in every channel group => plain read-modify-write, fixed order
        float part[D];
        weight_grad_to_key_grad<D>(p, gw, part);
#pragma unroll
        for (int a2 = 0; a2 < D; ++a2) {
          float* gp = grad_keys + ((size_t)unit * D + a2) * N + n;
          *gp = f0 == 0 ? part[a2] : *gp + part[a2];
        }
      }
    }
    }
  }
  // launched with programmatic stream serialization behind the grad_grid scatter of Slice backward: nothing above
  // reads what that kernel writes, but the grid must not RETIRE before it, or the next kernel on the stream could
  // start while the scatter still runs (PTX: a dependent grid uses griddepcontrol.wait for correct ordering).
  // Without the launch attribute this is a no-op.
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

inline bool gather_config(const ctb_shape* s, int mode, TileConfig* out, bool aligned16 = true) {
  // quad lanes measured faster only for the 3-D Slice forward (c3d 0.16 -> 0.13 ms)
  const int arrays = mode == GATHER_SPLAT_BWD ? 2 : 1;
  const bool quad = mode == GATHER_SLICE_FWD && s->dim == 3;
  if (!tile_config(s, arrays, 1, out, quad)) return false;
  // bf16 storage, plane-major, rows that are 16-byte multiples: the tile stays raw bf16 (see tile_gather_kernel),
  // so the same budget holds twice the rows
  const int stride0 = s->dim == 2 ? s->size[1] : s->size[1] * s->size[2];
  if (s->grid_dtype == CTB_DTYPE_BF16 && out->layout == TILE_PM4 && aligned16 && stride0 % 8 == 0) {
    TileConfig c2;
    if (tile_config(s, arrays, 1, &c2, quad, mode == GATHER_SPLAT_BWD ? 6 : 2) && c2.layout == TILE_PM4) *out = c2;
  }
  return true;
}

template <int D, int MODE, int LAYOUT, typename GT>
cudaError_t launch_gather(const float* keys, const GT* t1, const int* t2, const float* in, const float* pad,
                          float* out, float* grad_keys, const ctb_shape* s, const TileConfig& c, cudaStream_t stream,
                          bool overlap_prev) {
  const Grid<D> g = make_grid<D>(s->size);
  static const bool no_split = getenv("CTB_GATHER_NO_SPLIT") != nullptr;
  const int gsplit = (MODE == GATHER_SLICE_FWD && !no_split) ? (s->F + c.FG - 1) / c.FG : 1;
  long long blocks = (long long)s->B * s->H * c.slabs * gsplit;
  // too few work items for 148 SMs x 2 CTAs: split the points of an item over up to 8 CTAs (>= 1024 points each)
  // as long as a CTA keeps >= 8 corner entries per tile cell (the tile is loaded by every CTA of the item)
  int psplit = 1;
  const long long tile_cells = (long long)(c.slabs == 1 ? s->size[0] : c.R + 1) *
                               (s->dim == 2 ? s->size[1] : s->size[1] * s->size[2]);
  const long long entries = (long long)s->N << s->dim;
  while (psplit < 8 && blocks * psplit * 2 <= 2 * 296 && entries / (psplit * 2) >= 8 * tile_cells &&
         s->N / (psplit * 2) >= 2048)
    psplit *= 2;
  blocks *= psplit;
  if (blocks >= (1ll << 31)) return cudaErrorNotSupported;
  cudaError_t e = cudaFuncSetAttribute(tile_gather_kernel<D, MODE, LAYOUT, GT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
  if (e != cudaSuccess) return e;
  static const bool no_pdl = getenv("CTB_NO_PDL") != nullptr;
  if (overlap_prev && !no_pdl) {
    // Programmatic dependent launch: this grid does not consume anything the previous kernel on the stream writes
    // (the caller vouches for that), so its CTAs may start as soon as the previous grid's last wave is resident
    // (that kernel signals griddepcontrol.launch_dependents at its start) and fill the tail of its last wave.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3((unsigned)tile_threads());
    cfg.dynamicSmemBytes = c.smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, tile_gather_kernel<D, MODE, LAYOUT, GT>, keys, t1, t2, in, pad, out, grad_keys, g, s->H,
                              s->F, s->N, c.FG, c.R, c.slabs, c.words, gsplit, psplit);
  }
  tile_gather_kernel<D, MODE, LAYOUT, GT><<<(unsigned)blocks, tile_threads(), c.smem, stream>>>(
      keys, t1, t2, in, pad, out, grad_keys, g, s->H, s->F, s->N, c.FG, c.R, c.slabs, c.words, gsplit, psplit);
  return cudaGetLastError();
}

template <int D, int MODE, typename GT>
cudaError_t tile_gather(const float* keys, const GT* t1, const int* t2, const float* in, const float* pad,
                        float* out, float* grad_keys, const ctb_shape* s, cudaStream_t stream, bool overlap_prev = false) {
  TileConfig c;
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(t1) | reinterpret_cast<uintptr_t>(t2)) & 15) == 0;
  if (!gather_config(s, MODE, &c, aligned16)) return cudaErrorNotSupported;
  const int stride0 = s->dim == 2 ? s->size[1] : s->size[1] * s->size[2];
  switch (effective_layout(c.layout, t1, t2, std::is_same<GT, __nv_bfloat16>::value, stride0)) {
    case TILE_PM4: return launch_gather<D, MODE, TILE_PM4, GT>(keys, t1, t2, in, pad, out, grad_keys, s, c, stream, overlap_prev);
    case TILE_PM1: return launch_gather<D, MODE, TILE_PM1, GT>(keys, t1, t2, in, pad, out, grad_keys, s, c, stream, overlap_prev);
    case TILE_CLQ: return launch_gather<D, MODE, TILE_CLQ, GT>(keys, t1, t2, in, pad, out, grad_keys, s, c, stream, overlap_prev);
    default: return launch_gather<D, MODE, TILE_CL, GT>(keys, t1, t2, in, pad, out, grad_keys, s, c, stream, overlap_prev);
  }
}

}  // namespace ctb
