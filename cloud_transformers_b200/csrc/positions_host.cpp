// positions_host.cpp -- host build of ctb_positions.cuh (TEST INFRASTRUCTURE: lets the CPU test suite check
// the exact arithmetic the kernels use against the oracle without a GPU; never used by the product path).
// Build: g++ -O2 -ffp-contract=off -shared -fPIC positions_host.cpp -o libctb_pos_host.so
#include <stdint.h>
#include "ctb_positions.cuh"

template <int D>
static void run(const float* keys, float* lc, int64_t* idx, float* gk, const float* glc, int units, int N,
                const int32_t* size) {
  constexpr int S = 1 << D;
  const ctb::Grid<D> g = ctb::make_grid<D>(size);
  for (int u = 0; u < units; ++u)
    for (int n = 0; n < N; ++n) {
      const ctb::Pos<D> p = ctb::point_pos<D>(keys + (size_t)u * D * N, n, N, g);
      float gw[S];
      for (int s = 0; s < S; ++s) {
        lc[((size_t)u * S + s) * N + n] = ctb::corner_weight<D>(p, s);
        idx[((size_t)u * S + s) * N + n] = p.base + ctb::corner_offset<D>(g, s);
        gw[s] = glc ? glc[((size_t)u * S + s) * N + n] : 0.0f;
      }
      if (gk && glc) {
        float out[D];
        ctb::weight_grad_to_key_grad<D>(p, gw, out);
        for (int a = 0; a < D; ++a) gk[((size_t)u * D + a) * N + n] = out[a];
      }
    }
}

extern "C" int ctb_host_positions(const float* keys, float* lc, int64_t* idx, float* grad_keys,
                                  const float* grad_lc, int units, int N, int dim, const int32_t* size) {
  if (dim == 2) run<2>(keys, lc, idx, grad_keys, grad_lc, units, N, size);
  else if (dim == 3) run<3>(keys, lc, idx, grad_keys, grad_lc, units, N, size);
  else return -1;
  return 0;
}
