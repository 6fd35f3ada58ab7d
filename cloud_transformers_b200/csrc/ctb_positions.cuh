// ctb_positions.cuh -- the per-point grid-position arithmetic shared by every kernel.
//
// Bit-exact replay of the reference's float32 op sequence (one rounding per op, no FMA contraction):
//   layers/cloud_transform.py:91   k = clamp(key, -1 + 1e-7, 1 - 1e-7)     (bounds rounded to f32)
//   layers/cloud_transform.py:94   x = (k + 1.0) * ((W - 1) * 0.5)         (GradientBalancing fwd)
//   layers/utils.py:112/166        fl = floor(x)
//   layers/utils.py:144-151/179-182  w_s = prod_a ((fl_a + 1) - x_a  |  x_a - fl_a)   left-assoc
//   layers/cloud_transform.py:113-119 flat = x*W1*W2 + y*W2 + z
// The same code compiles for the host (g++ -ffp-contract=off) so that tests can check it against the
// oracle without a GPU (tests/test_positions_host.py).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CTB_HD __host__ __device__ __forceinline__
#else
#define CTB_HD inline
#endif

namespace ctb {

#if defined(__CUDA_ARCH__)
#define CTB_FADD(a, b) __fadd_rn((a), (b))
#define CTB_FSUB(a, b) __fsub_rn((a), (b))
#define CTB_FMUL(a, b) __fmul_rn((a), (b))
#else
// host build: compiled with -ffp-contract=off so a*b+c is never fused
static inline float ctb_host_add(float a, float b) { volatile float r = a + b; return r; }
static inline float ctb_host_sub(float a, float b) { volatile float r = a - b; return r; }
static inline float ctb_host_mul(float a, float b) { volatile float r = a * b; return r; }
#define CTB_FADD(a, b) ::ctb::ctb_host_add((a), (b))
#define CTB_FSUB(a, b) ::ctb::ctb_host_sub((a), (b))
#define CTB_FMUL(a, b) ::ctb::ctb_host_mul((a), (b))
#endif

// float32(-1 + 1e-7) = -0x1.fffffcp-1 and float32(1 - 1e-7) = 0x1.fffffcp-1
#define CTB_KEY_LO (-0.99999988079071044921875f)
#define CTB_KEY_HI (0.99999988079071044921875f)

// geometry of a grid, passed by value to kernels
template <int D>
struct Grid {
  int W[D];        // extent per axis
  float scale[D];  // (W - 1) * 0.5 in float32
  int stride[D];   // flat-index stride per axis (axis 0 slowest)
  int C;           // cells
};

template <int D>
inline Grid<D> make_grid(const int32_t* size) {
  Grid<D> g;
  int c = 1;
  for (int a = D - 1; a >= 0; --a) {
    g.W[a] = size[a];
    g.stride[a] = c;
    c *= size[a];
    g.scale[a] = ((float)size[a] - 1.0f) * 0.5f;
  }
  g.C = c;
  return g;
}

// One point's position on the grid: base cell (floor corner) and the two per-axis linear weights.
template <int D>
struct Pos {
  float up[D];  // (fl + 1) - x : weight factor of the floor corner on this axis  (corner bit = 0)
  float dn[D];  // x - fl       : weight factor of the +1 corner on this axis      (corner bit = 1)
  int base;     // flat index of corner s = 0
  int c0;       // cell along axis 0 (grid row) of corner s = 0
  bool in_range[D];  // lo <= key <= hi (clamp passes gradient), cloud_transform.py:91
};

CTB_HD float clamp_key(float key) {
  // NaN keys clamp to CTB_KEY_LO (fmaxf drops the NaN) so indices always stay inside the grid; the
  // reference would trip its bounds assert instead (cloud_transform.py:101-102).
  return fminf(fmaxf(key, CTB_KEY_LO), CTB_KEY_HI);
}

template <int D>
CTB_HD void axis_pos(float key, float scale, int W, float& up, float& dn, int& cell, bool& in_range) {
  in_range = (key >= CTB_KEY_LO) && (key <= CTB_KEY_HI);
  const float k = clamp_key(key);
  const float x = CTB_FMUL(CTB_FADD(k, 1.0f), scale);
  const float fl = floorf(x);
  up = CTB_FSUB(CTB_FADD(fl, 1.0f), x);
  dn = CTB_FSUB(x, fl);
  int c = (int)fl;
  // x < W-1 always holds for clamped keys; the clamp only guards memory safety.
  c = c < 0 ? 0 : (c > W - 2 ? W - 2 : c);
  cell = c;
}

// keys_u points at this unit's [D][N] slab (stride N between axes).
template <int D>
CTB_HD Pos<D> point_pos(const float* __restrict__ keys_u, int n, int N, const Grid<D>& g) {
  Pos<D> p;
  int base = 0;
#pragma unroll
  for (int a = 0; a < D; ++a) {
    int c;
    axis_pos<D>(keys_u[(size_t)a * N + n], g.scale[a], g.W[a], p.up[a], p.dn[a], c, p.in_range[a]);
    base += c * g.stride[a];
    if (a == 0) p.c0 = c;
  }
  p.base = base;
  return p;
}

// same as point_pos, from key values already in registers
template <int D>
CTB_HD Pos<D> point_pos_from_values(const float* kv, const Grid<D>& g) {
  Pos<D> p;
  int base = 0;
#pragma unroll
  for (int a = 0; a < D; ++a) {
    int c;
    axis_pos<D>(kv[a], g.scale[a], g.W[a], p.up[a], p.dn[a], c, p.in_range[a]);
    base += c * g.stride[a];
    if (a == 0) p.c0 = c;
  }
  p.base = base;
  return p;
}

// weight of corner s: left-associated product over axes, exactly as the reference multiplies.
template <int D>
CTB_HD float corner_weight(const Pos<D>& p, int s) {
  float w = (s & 1) ? p.dn[0] : p.up[0];
#pragma unroll
  for (int a = 1; a < D; ++a) w = CTB_FMUL(w, ((s >> a) & 1) ? p.dn[a] : p.up[a]);
  return w;
}

template <int D>
CTB_HD int corner_offset(const Grid<D>& g, int s) {
  int o = 0;
#pragma unroll
  for (int a = 0; a < D; ++a) o += ((s >> a) & 1) ? g.stride[a] : 0;
  return o;
}

// d(sum_s gw[s] * w_s) / d x_a, times the clamp mask: the chain through floor is zero and the chain
// through GradientBalancing is the identity (cloud_transform.py:21-23).
template <int D>
CTB_HD void weight_grad_to_key_grad(const Pos<D>& p, const float* gw, float* gk) {
#pragma unroll
  for (int a = 0; a < D; ++a) {
    float acc = 0.0f;
#pragma unroll
    for (int s = 0; s < (1 << D); ++s) {
      float other = 1.0f;
#pragma unroll
      for (int a2 = 0; a2 < D; ++a2)
        if (a2 != a) other *= ((s >> a2) & 1) ? p.dn[a2] : p.up[a2];
      const float t = gw[s] * other;
      acc += ((s >> a) & 1) ? t : -t;
    }
    gk[a] = p.in_range[a] ? acc : 0.0f;
  }
}

}  // namespace ctb
