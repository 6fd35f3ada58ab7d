// ctb_binned.cuh -- output-stationary scatters over the plan's sorted entry list: no atomics, no position arithmetic.
//
// The point-lane tile scatters (ctb_tile.cuh) issue one shared-memory atomic per (point, corner, channel); on the
// coarse grids of the MHCT blocks (16^3 x F16, 16^2 x F16, 8^3 x F32: 4 .. 32 entries per cell on average, hundreds in
// the cells a surface crosses) ncu shows 72-79 % of their shared wavefronts as bank conflicts and ~40 warp
// instructions per (entry, channel) warp-step, half of them position arithmetic repeated per channel group and pass
// (profiles/r01_ncu_full_summary.csv).  Here the plan (ctb_plan.cuh) has sorted the unit's (point, corner) entries by
// destination cell, ascending e = s N + n inside a cell, weights attached.  The scatter is then a SEGMENTED REDUCTION
// over that list, done in registers:
//
//   CTA = (unit, group of FG = 4 LP channels).  The group's features (times the padding mask) are staged once as
//   xs[n][FG].  The entry list streams through shared memory in chunks that end on cell boundaries (<= cap entries,
//   <= CC cells).  Inside a chunk every lane group (LP lanes, 4 channels each) takes an EQUAL window of consecutive
//   entries -- the work per lane does not depend on how the points cluster -- and walks it:
//       v = xs[n] * w;   max:  if (v > cur) { cur = v; tag = entry; }     sum:  acc = acc + v
//   ("first strictly greater wins" from the reference's zero floor, layers/cloud_transform.py:164-173, is
//   torch-scatter's CPU rule because the list is in ascending e).  When the cell changes, the finished segment goes
//   to the chunk's output tile in shared memory.  A cell that spans several windows is closed by the window that
//   holds its last entry: it folds the partial results of the earlier windows (kept in shared memory) in window
//   order, so ties still go to the smallest e and sums are added in a fixed order -- bit-identical from run to run.
//   The tile leaves with coalesced stores; every cell of z / arg / grad_grid is written exactly once.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "../../include/ctb200.h"
#include "ctb_plan.cuh"
#include "ctb_positions.cuh"
#include "ctb_tile.cuh"

namespace ctb {

constexpr int kBinThreadsMax = 1024;
constexpr int kBinSmemTwoCtas = 112 * 1024;
constexpr int kBinSmemMax = 224 * 1024;

struct BinConfig {
  int LP;        // lanes per window; a lane holds 4 channels, FG = 4 LP
  int groups;
  int threads;
  int CC;        // cells per chunk (the output tile)
  size_t smem;
};

inline size_t bin_smem_bytes(const ctb_shape* s, bool sum, int LP, int threads, int CC) {
  const size_t CS = (size_t)((shape_cells(s) + 2 + 7) & ~7ll);
  const size_t FG = 4 * (size_t)LP, G = (size_t)threads / LP;
  const size_t arrays = sum ? 1 : 2;
  return (size_t)s->N * FG * 4 + arrays * FG * CC * 4 + arrays * G * FG * 4 + G * 4 + CS * 2;
}

// Channel group, CTA size and tile size.  Wide groups (16 channels = 4 lanes per window) keep the number of windows
// per warp at 8, so the cell-boundary code of a window is shared by fewer diverging lanes; the whole grid in one
// tile (coarse grids) gives every window 1 / G of the unit's entries.
inline bool bin_config(const ctb_shape* s, bool sum, BinConfig* out) {
  if (!plan_has_entries(s)) return false;
  static const int env_lp = getenv("CTB_BIN_LP") ? atoi(getenv("CTB_BIN_LP")) : 0;
  static const int env_cc = getenv("CTB_BIN_CC") ? atoi(getenv("CTB_BIN_CC")) : 0;
  static const int env_threads = getenv("CTB_BIN_THREADS") ? atoi(getenv("CTB_BIN_THREADS")) : 0;
  const int Cp = ((int)shape_cells(s) + 3) & ~3;
  int lp_max = 1;
  while (lp_max < 4 && lp_max * 4 < s->F) lp_max <<= 1;
  if (env_lp) lp_max = env_lp;
  for (int LP = lp_max; LP >= 1; LP >>= 1) {
    for (int pass = 0; pass < 2; ++pass) {
      const int threads = env_threads ? env_threads : (pass == 0 ? 512 : 1024);
      const size_t budget = pass == 0 ? kBinSmemTwoCtas : kBinSmemMax;
      int CC = env_cc ? (env_cc < Cp ? env_cc : Cp) : Cp;
      while (true) {
        const size_t b = bin_smem_bytes(s, sum, LP, threads, CC);
        if (b <= budget) {
          out->LP = LP;
          out->groups = (s->F + 4 * LP - 1) / (4 * LP);
          out->threads = threads;
          out->CC = CC;
          out->smem = b;
          return true;
        }
        if (env_cc || CC <= 256) break;
        CC = (CC / 2 + 3) & ~3;
      }
    }
  }
  return false;
}

enum { BIN_NONE = 0, BIN_TAIL = 1, BIN_WHOLE = 2 };

template <bool SUM, int LP, typename GT>
__global__ void __launch_bounds__(kBinThreadsMax, 1)
ent_scatter_kernel(const float* __restrict__ feat, const float* __restrict__ pad, const uint2* __restrict__ ent,
                   const uint16_t* __restrict__ cstart, GT* __restrict__ z, int* __restrict__ arg, int C, int H, int F,
                   int N, int E, int CS, int groups, int nbits, int dim, int CC) {
  constexpr int FG = 4 * LP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = blockDim.x;
  const int G = T / LP;                                           // windows per chunk
  float* xs = (float*)smem_raw;                                   // [N][FG]   features * pad
  float* ov = xs + (size_t)N * FG;                                // [CC][FG]  output tile: values
  unsigned* ot = (unsigned*)(ov + (SUM ? 0 : FG * CC));           // [CC][FG]  output tile: winning entry (max only)
  float* pacc = (float*)(ot + FG * CC);                           // [G][FG]   partial of a window's open segment
  unsigned* ptag = (unsigned*)(pacc + (SUM ? 0 : G * FG));        // [G][FG]
  int* kind = (int*)(ptag + G * FG);                              // [G]
  uint16_t* cs = (uint16_t*)(kind + G);                           // [CS]      first entry of every cell

  asm volatile("griddepcontrol.launch_dependents;");   // see tile_scatter_kernel
  CTB_STAMP_INIT;
  const int f0 = (blockIdx.x % groups) * FG;
  const int unit = blockIdx.x / groups;
  const int fgn = min(FG, F - f0);
  const float* pu = pad ? pad + (size_t)(unit / H) * N : nullptr;
  const float* fu = feat + ((size_t)unit * F + f0) * N;
  const uint2* eu = ent + (size_t)unit * E;

  // ---- staging: cell starts, features (consecutive lanes = consecutive points of one channel quad) ------------
  {
    const uint4* src = reinterpret_cast<const uint4*>(cstart + (size_t)unit * CS);
    uint4* dst = reinterpret_cast<uint4*>(cs);
    for (int i = threadIdx.x; i < (CS >> 3); i += T) dst[i] = __ldg(src + i);
  }
  // (a lane takes (point, channel quad): the LP lanes of a point store its 16 LP bytes contiguously -- no bank
  // conflicts -- and a warp's load of one channel plane covers full 32-byte sectors)
#pragma unroll 4
  for (int i = threadIdx.x; i < N * LP; i += T) {
    const int n = i / LP, qq = i % LP;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (4 * qq + k < fgn) ? __ldg(fu + (size_t)(4 * qq + k) * N + n) : 0.0f;
    if (pu) {
      const float pd = __ldg(pu + n);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = CTB_FMUL(v[k], pd);
    }
    *reinterpret_cast<float4*>(xs + (size_t)n * FG + 4 * qq) = make_float4(v[0], v[1], v[2], v[3]);
  }

  const int g = threadIdx.x / LP, q = threadIdx.x % LP;
  const unsigned xs_q = smem_u32(xs) + (unsigned)q * 16u;
  const unsigned nmask = (1u << nbits) - 1u, smask = (1u << dim) - 1u;
  const int ebits = nbits + dim;
  // the output tile is channel-last [cell][FG]: a finished segment leaves with one 16-byte store per array
  const unsigned ov_q = smem_u32(ov) + (unsigned)q * 16u, ot_q = smem_u32(ot) + (unsigned)q * 16u;
  __syncthreads();
  CTB_STAMP(10);

  int ca = 0;
  while (ca < C) {
    // chunk = cells [ca, cb): as many as the output tile holds; its entries are read straight from L1 / L2 (every
    // window walks a contiguous run of 8-byte records, 16 per line)
    const int base = cs[ca];
    const int cb = min(C, ca + CC);
    const int ncell = cb - ca;
    const int cnt = (int)cs[cb] - base;
    const uint2* rec = eu + base;
    for (int i = threadIdx.x; i < FG * CC; i += T) {
      ov[i] = 0.0f;
      if constexpr (!SUM) ot[i] = 0xffffffffu;
    }
    __syncthreads();
    CTB_STAMP(11);

    // ---- windows ---------------------------------------------------------------------------------------
    // The 8-byte records of a window are fetched ONCE, 16 bytes per lane: the LP lanes of a window load one aligned
    // block of EB = 2 LP consecutive records (64 bytes at LP = 4) and hand them round by shuffle, the next block
    // being in flight meanwhile.  (Per-record loads re-read a 128-byte line 16 times; with ~225 KB of shared memory
    // carved out, the 28 KB that is left of L1 does not hold the 256 lines the windows of a CTA are walking.)
    constexpr int EB = 2 * LP;
    const int L = (cnt + G - 1) / G;
    const int j0 = min(cnt, g * L), j1 = min(cnt, j0 + L);
    const bool active = j0 < j1;
    int cur = 0;
    bool open_left = false;
    if (active) {
      cur = (int)(__ldg(&rec[j0].x) >> ebits);
      open_left = j0 > 0 && (int)(__ldg(&rec[j0 - 1].x) >> ebits) == cur;
    }
    const int hcell = cur;
    int my_kind = BIN_NONE;
    bool head_closing = false;
    {
      unsigned rel = (unsigned)(cur - ca) * (FG * 4u);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      unsigned g0 = 0xffffffffu, g1 = 0xffffffffu, g2 = 0xffffffffu, g3 = 0xffffffffu;
      // one record of the window
      auto take = [&](unsigned tg, unsigned wb) {
        float v0, v1, v2, v3;
        const float w = __uint_as_float(wb);
        const unsigned xa = xs_q + (tg & nmask) * (FG * 4u);
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3) : "r"(xa));
        const int c = (int)(tg >> ebits);
        if (c != cur) {
          // the segment is finished (a first segment that continues an earlier window is finished off after the
          // barrier: its partial waits in the tile like any other result)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ov_q + rel), "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
          if constexpr (!SUM)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ot_q + rel), "r"(g0), "r"(g1), "r"(g2), "r"(g3) : "memory");
          cur = c;
          rel = (unsigned)(c - ca) * (FG * 4u);
          a0 = a1 = a2 = a3 = 0.0f;
          g0 = g1 = g2 = g3 = 0xffffffffu;
        }
        const float t0 = CTB_FMUL(v0, w), t1 = CTB_FMUL(v1, w), t2 = CTB_FMUL(v2, w), t3 = CTB_FMUL(v3, w);
        if constexpr (SUM) {
          a0 = CTB_FADD(a0, t0);
          a1 = CTB_FADD(a1, t1);
          a2 = CTB_FADD(a2, t2);
          a3 = CTB_FADD(a3, t3);
        } else {
          if (t0 > a0) { a0 = t0; g0 = tg; }
          if (t1 > a1) { a1 = t1; g1 = tg; }
          if (t2 > a2) { a2 = t2; g2 = tg; }
          if (t3 > a3) { a3 = t3; g3 = tg; }
        }
      };
      if (L < 2 * EB) {
        // short windows (fine grids cut into many tiles): block fetches would mostly load records of the neighbours
#pragma unroll 2
        for (int j = j0; j < j1; ++j) {
          const uint2 r = __ldg(rec + j);
          take(r.x, r.y);
        }
      } else {
        const int A0 = base + j0, A1 = base + j1;            // the window in record indices of the unit
        const int NB = (L + 2 * EB - 2) / EB;                // blocks a window can touch (same for the whole CTA)
        const uint4* e4 = reinterpret_cast<const uint4*>(eu);
        int blk = A0 / EB;
        auto fetch = [&](int bk) {
          const int p = bk * LP + q;                          // record pair
          return (active && bk * EB < A1 && 2 * p < E) ? __ldg(e4 + p) : make_uint4(0u, 0u, 0u, 0u);
        };
        uint4 rv = fetch(blk);
        for (int t = 0; t < NB; ++t, ++blk) {
          const uint4 rn = fetch(blk + 1);
          const int ab = blk * EB;
#pragma unroll
          for (int k = 0; k < EB; ++k) {
            unsigned tg = (k & 1) ? rv.z : rv.x, wb = (k & 1) ? rv.w : rv.y;
            if constexpr (LP > 1) {
              tg = __shfl_sync(0xffffffffu, tg, k >> 1, LP);
              wb = __shfl_sync(0xffffffffu, wb, k >> 1, LP);
            }
            const int a = ab + k;
            if (a >= A0 && a < A1) take(tg, wb);
          }
          rv = rn;
        }
      }
      if (active) {
        const bool open_right = j1 < cnt && (int)(__ldg(&rec[j1].x) >> ebits) == cur;
        const bool single = cur == hcell;                    // the window never left its first cell
        if (open_right) {
          // the cell goes on in the next window: park the partial (the closing window folds it in)
          my_kind = (single && open_left) ? BIN_WHOLE : BIN_TAIL;
          *reinterpret_cast<float4*>(pacc + (size_t)g * FG + 4 * q) = make_float4(a0, a1, a2, a3);
          if constexpr (!SUM) *reinterpret_cast<uint4*>(ptag + (size_t)g * FG + 4 * q) = make_uint4(g0, g1, g2, g3);
        } else {
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ov_q + rel), "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
          if constexpr (!SUM)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ot_q + rel), "r"(g0), "r"(g1), "r"(g2), "r"(g3) : "memory");
        }
        head_closing = open_left && !(single && open_right);
      }
    }
    if (q == 0) kind[g] = my_kind;
    CTB_STAMP(15);
    __syncthreads();
    CTB_STAMP(12);
    // ---- cells that span windows: the window that closes the cell folds the parked partials in window order,
    // then its own (already in the tile): ties keep the earliest entry, sums have a fixed order ----------------------
    if (head_closing) {
      int g1w = g - 1;
      while (kind[g1w] == BIN_WHOLE) --g1w;
      float4 ra = *reinterpret_cast<const float4*>(pacc + (size_t)g1w * FG + 4 * q);
      uint4 rt = make_uint4(0, 0, 0, 0);
      if constexpr (!SUM) rt = *reinterpret_cast<const uint4*>(ptag + (size_t)g1w * FG + 4 * q);
      auto fold = [&](const float4 b, const uint4 u) {
        if constexpr (SUM) {
          ra.x = CTB_FADD(ra.x, b.x); ra.y = CTB_FADD(ra.y, b.y); ra.z = CTB_FADD(ra.z, b.z); ra.w = CTB_FADD(ra.w, b.w);
        } else {
          if (b.x > ra.x) { ra.x = b.x; rt.x = u.x; }
          if (b.y > ra.y) { ra.y = b.y; rt.y = u.y; }
          if (b.z > ra.z) { ra.z = b.z; rt.z = u.z; }
          if (b.w > ra.w) { ra.w = b.w; rt.w = u.w; }
        }
      };
      for (int h = g1w + 1; h < g; ++h) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if constexpr (!SUM) u = *reinterpret_cast<const uint4*>(ptag + (size_t)h * FG + 4 * q);
        fold(*reinterpret_cast<const float4*>(pacc + (size_t)h * FG + 4 * q), u);
      }
      float4* tv = reinterpret_cast<float4*>(ov + (size_t)(hcell - ca) * FG + 4 * q);
      uint4* tt = reinterpret_cast<uint4*>(ot + (size_t)(hcell - ca) * FG + 4 * q);
      uint4 u = make_uint4(0, 0, 0, 0);
      if constexpr (!SUM) u = *tt;
      fold(*tv, u);
      *tv = ra;
      if constexpr (!SUM) *tt = rt;
    }
    __syncthreads();
    CTB_STAMP(13);
    // ---- the tile leaves: a lane takes (cell, channel quad), a warp writes 32-byte runs of 4 planes -------------
    GT* zu = z + ((size_t)unit * F + f0) * C + ca;
    int* au = (!SUM && arg != nullptr) ? arg + ((size_t)unit * F + f0) * C + ca : nullptr;
    for (int i = threadIdx.x; i < ncell * LP; i += T) {
      const int r = i / LP, qq = i % LP;
      const float4 v = *reinterpret_cast<const float4*>(ov + (size_t)r * FG + 4 * qq);
      const float vv[4] = {v.x, v.y, v.z, v.w};
      unsigned tt[4] = {0, 0, 0, 0};
      if constexpr (!SUM) {
        const uint4 t = *reinterpret_cast<const uint4*>(ot + (size_t)r * FG + 4 * qq);
        tt[0] = t.x; tt[1] = t.y; tt[2] = t.z; tt[3] = t.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int f = 4 * qq + k;
        if (f < fgn) {
          grid_store(zu + (size_t)f * C + r, vv[k]);
          if constexpr (!SUM) {
            if (au) __stcs(au + (size_t)f * C + r,
                           tt[k] == 0xffffffffu ? -1 : (int)((tt[k] >> nbits) & smask) * N + (int)(tt[k] & nmask));
          }
        }
      }
    }
    __syncthreads();
    CTB_STAMP(14);
    ca = cb;
  }
}

template <bool SUM, typename GT>
cudaError_t bin_scatter(const float* feat, const float* pad, const void* plan, GT* z, int* arg, const ctb_shape* s,
                        cudaStream_t stream) {
  BinConfig c;
  if (!bin_config(s, SUM, &c)) return cudaErrorNotSupported;
  const PlanView v = plan_view(plan, s);
  const long long blocks = (long long)s->B * s->H * c.groups;
  if (blocks >= (1ll << 31)) return cudaErrorNotSupported;
  const int C = (int)shape_cells(s);
  auto launch = [&](auto kernel) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
    if (e != cudaSuccess) return e;
    kernel<<<(unsigned)blocks, c.threads, c.smem, stream>>>(feat, pad, v.ent, v.cstart, z, arg, C, s->H, s->F, s->N, v.E,
                                                            v.CS, c.groups, v.nbits, s->dim, c.CC);
    return cudaGetLastError();
  };
  switch (c.LP) {
    case 8: return launch(ent_scatter_kernel<SUM, 8, GT>);
    case 4: return launch(ent_scatter_kernel<SUM, 4, GT>);
    case 2: return launch(ent_scatter_kernel<SUM, 2, GT>);
    default: return launch(ent_scatter_kernel<SUM, 1, GT>);
  }
}

}  // namespace ctb
