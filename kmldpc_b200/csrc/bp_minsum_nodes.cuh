// Node updates of the normalised min-sum decoder (bp_minsum.cu; also the min-sum variant of the quasi-cyclic plan
// kernel in bp_decode.cu).  See bp_minsum.cu for the algorithm statement and the word format.
#ifndef KML_BP_MINSUM_NODES_CUH
#define KML_BP_MINSUM_NODES_CUH
#include <cstdint>

#include "kml_internal.h"

namespace kml {
namespace msn {

constexpr float kLlrClip = 27.631021f;  // ln((1-1e-12)/1e-12)

__device__ __forceinline__ float load_channel_llr(const float *in, int idx, int in_is_lr) {
  float v = __ldg(in + idx);
  if (in_is_lr) v = __logf(fminf(fmaxf(v, kLrMin), kLrMax));
  return fminf(fmaxf(v, -kLlrClip), kLlrClip);
}

template <int D>
__device__ __forceinline__ uint32_t ms_vn(uint32_t *msg, const uint32_t *a, float ch) {
  float x[D], total = ch;
#pragma unroll
  for (int k = 0; k < D; k++) {
    x[k] = __uint_as_float(msg[a[k]]);
    total += x[k];
  }
  const uint32_t bit = (total > 0.0f) ? 0u : 1u;  // tie → 1, like alpha0 > alpha1 ? 0 : 1
#pragma unroll
  for (int k = 0; k < D; k++) msg[a[k]] = (__float_as_uint(total - x[k]) & ~1u) | bit;
  return bit;
}

// returns the XOR of the row's words: bit 31 = sign parity, bit 0 = syndrome of the current decisions
template <int D>
__device__ __forceinline__ uint32_t ms_cn(uint32_t *msg, int plane, int slot, float alpha, float beta = 0.0f) {
  uint32_t w[D], x = 0;
  float m1 = 3.0e38f, m2 = 3.0e38f, a[D];
#pragma unroll
  for (int k = 0; k < D; k++) {
    w[k] = msg[k * plane + slot];
    x ^= w[k];
    a[k] = fabsf(__uint_as_float(w[k]));
    m2 = fminf(m2, fmaxf(m1, a[k]));
    m1 = fminf(m1, a[k]);
  }
  // normalised (beta = 0) / offset (alpha = 1) / both: magnitude = max(alpha m - beta, 0), clipped like the reference's messages
  const float s1 = fminf(fmaxf(fmaf(alpha, m1, -beta), 0.0f), kLlrClip), s2 = fminf(fmaxf(fmaf(alpha, m2, -beta), 0.0f), kLlrClip);
#pragma unroll
  for (int k = 0; k < D; k++) {
    const float mag = (a[k] == m1) ? s2 : s1;  // ties at the minimum: m2 == m1, either choice gives the same value
    msg[k * plane + slot] = __float_as_uint(mag) | ((x ^ w[k]) & 0x80000000u);
  }
  return x;
}

}  // namespace msn
}  // namespace kml
#endif
