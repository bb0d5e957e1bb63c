// Node updates of the normalised min-sum decoder (bp_minsum.cu; also the min-sum variant of the quasi-cyclic plan
// kernel in bp_decode.cu).  See bp_minsum.cu for the algorithm statement and the word format.
#ifndef KML_BP_MINSUM_NODES_CUH
#define KML_BP_MINSUM_NODES_CUH
#include <cstdint>

#include <cuda_fp16.h>

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

// ---- fp16 messages, two frames per 32-bit word (algorithm = 2): the halves of a word are the same edge of two frames;
// mantissa bit 0 of each half carries that frame's posterior decision.
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }

// returns the two decisions: bit 0 = first frame, bit 16 = second frame
template <int D>
__device__ __forceinline__ uint32_t ms2_vn(uint32_t *msg, const uint32_t *a, __half2 ch) {
  __half2 x[D], total = ch;
#pragma unroll
  for (int k = 0; k < D; k++) {
    x[k] = u2h(msg[a[k]]);
    total = __hadd2(total, x[k]);
  }
  const uint32_t pb = __hle2_mask(total, __float2half2_rn(0.0f)) & 0x00010001u;  // decision 1 unless total > 0 (tie → 1)
#pragma unroll
  for (int k = 0; k < D; k++) msg[a[k]] = (h2u(__hsub2(total, x[k])) & 0xFFFEFFFEu) | pb;
  return pb;
}

// returns the XOR of the row's words: per half, bit 15 = sign parity, bit 0 = syndrome of that frame's current decisions
template <int D>
__device__ __forceinline__ uint32_t ms2_cn(uint32_t *row, __half2 alpha2, __half2 nbeta2, bool has_offset) {
  uint32_t w[D], x = 0;
  __half2 a[D], m1 = __float2half2_rn(60000.0f), m2 = m1;
#pragma unroll
  for (int k = 0; k < D; k++) {
    w[k] = row[k];
    x ^= w[k];
    a[k] = u2h(w[k] & 0x7FFE7FFEu);
    m2 = __hmin2(m2, __hmax2(m1, a[k]));
    m1 = __hmin2(m1, a[k]);
  }
  __half2 t1 = __hfma2(alpha2, m1, nbeta2), t2 = __hfma2(alpha2, m2, nbeta2);
  if (has_offset) {  // (uniform: the normalised rule never goes negative)
    t1 = __hmax2(t1, __float2half2_rn(0.0f));
    t2 = __hmax2(t2, __float2half2_rn(0.0f));
  }
  const __half2 clip2 = __float2half2_rn(kLlrClip);
  const uint32_t s1 = h2u(__hmin2(t1, clip2)), s2 = h2u(__hmin2(t2, clip2));
#pragma unroll
  for (int k = 0; k < D; k++) {
    const uint32_t eq = __heq2_mask(a[k], m1);
    row[k] = ((s2 & eq) | (s1 & ~eq)) | ((x ^ w[k]) & 0x80008000u);
  }
  return x;
}

}  // namespace msn
}  // namespace kml
#endif
