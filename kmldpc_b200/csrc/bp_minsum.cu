// Normalised min-sum flooding decoder — the THROUGHPUT mode (`algorithm = 1`), not reference-pinned.
//
// BASELINE.json's north_star lists min-sum next to sum-product, but the reference implements only the sum-product
// (SURVEY §0.3, §8(c)): there is no reference output to be bit-compared with.  This decoder therefore keeps everything
// that IS pinned — schedule (variable phase → hard decision → syndrome test → check phase), stopping rule, return value
// (binaryldpccodec.cc:175-277), shared-memory layout and frame queue of bp_decode.cu — and replaces only the node
// updates.  It is checked against an independent numpy statement of the same algorithm (oracle/minsum_ref.py) and gated
// by BER/FER against the sum-product decoder (tests/test_gpu_minsum.py).
//
//   VN   total = L_ch + sum_e c2v_e ;  v2c_e = total - c2v_e ;  decision = (total > 0) ? 0 : 1      (LLR = ln P0/P1)
//   CN   c2v_e = max(alpha * min_{e' != e} |v2c_e'| - beta, 0) * prod_{e' != e} sign(v2c_e'), clipped to +-27.63 (the
//        reference's clip): normalised min-sum (alpha = 0.8, beta = 0, the default), offset min-sum (alpha = 1, beta > 0)
//
// A v2c word is the float itself with mantissa bit 0 replaced by the variable's posterior decision (an LLR does not need
// its last bit), so a check node gets sign parity and the syndrome of the current decisions from one XOR chain.
// ~14 issue slots and no MUFU per edge-iteration, against 25 and 2.6 for the exact sum-product.
#include <cuda_fp16.h>

#include <cstdlib>

#include "kml_internal.h"
#include "kml_kernels.cuh"
#include "bp_minsum_nodes.cuh"

namespace kml {
namespace {

using namespace msn;

// ---- (3,6)-regular codes: same thread/frame organisation as bp_regular_kernel
template <int VPT, int CPT, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) ms_regular_kernel(const DecParams p) {
  extern __shared__ uint32_t msg[];
  __shared__ int s_frame;
  const int tid = threadIdx.x;
  constexpr int plane = CPT * T + 1;
  uint32_t va[VPT][3];
#pragma unroll
  for (int j = 0; j < VPT; j++)
#pragma unroll
    for (int k = 0; k < 3; k++) va[j][k] = p.t.vn_addr[(j * T + tid) * 3 + k];

  while (true) {
    if (tid == 0) s_frame = next_frame(p);
    __syncthreads();
    const int f = s_frame;
    if (f < 0) break;
    const float *in = p.in + (size_t)(p.sel ? f * p.n_cand + __ldg(p.sel + f) : f) * p.t.n_tx;
    float ch[VPT];
#pragma unroll
    for (int j = 0; j < VPT; j++) ch[j] = load_channel_llr(in, j * T + tid, p.in_is_lr);
    for (int i = tid; i < 6 * plane; i += T) msg[i] = 0u;  // c2v = (0.5, 0.5) → LLR 0
    __syncthreads();
    uint32_t bits = 0, latched_bits = 0;
    int ret = p.iters + (p.iters < p.max_iter);
    bool latched = false;
    for (int t = 0; t < p.iters; t++) {
      bits = 0;
#pragma unroll
      for (int j = 0; j < VPT; j++) bits |= ms_vn<3>(msg, va[j], ch[j]) << j;
      __syncthreads();
      uint32_t x = 0;
#pragma unroll
      for (int j = 0; j < CPT; j++) x |= ms_cn<6>(msg, plane, j * T + tid, p.alpha, p.beta);
      const int any_fail = __syncthreads_or((int)(x & 1u));
      if (!any_fail && !latched) {
        latched = true;
        latched_bits = bits;
        ret = t + (t < p.max_iter);
        if (p.early_exit) break;
      }
    }
    if (!latched) latched_bits = bits;
#pragma unroll
    for (int j = 0; j < VPT; j++) {
      const uint32_t word = __ballot_sync(0xffffffffu, (latched_bits >> j) & 1u);
      if ((tid & 31) == 0) p.out_bits[(size_t)f * p.words_n + ((j * T + tid) >> 5)] = word;
    }
    if (tid == 0) p.out_ret[f] = ret;
  }
}

// ---- any graph: same organisation as bp_generic_kernel (exact-degree dispatch)
#define KML_MS_VN_CASE(D)                                           \
  case D:                                                           \
    if (D <= DV) {                                                  \
      uint32_t a[(D <= DV ? D : 1)];                                \
      _Pragma("unroll") for (int k = 0; k < (D <= DV ? D : 1); k++) a[k] = __ldg(ad + k); \
      bit = ms_vn<(D <= DV ? D : 1)>(msg, a, chan[v]);              \
    }                                                               \
    break;
#define KML_MS_CN_CASE(D)                                           \
  case D:                                                           \
    if (D <= DC) x = ms_cn<(D <= DC ? D : 1)>(msg, plane, slot, p.alpha, p.beta); \
    break;

template <int DV, int DC>
__global__ void __launch_bounds__(512) ms_generic_kernel(const DecParams p) {
  static_assert(DV <= 16 && DC <= 16, "add switch cases");
  extern __shared__ uint32_t smem[];
  __shared__ int s_frame;
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31;
  const int mpad = p.t.m_pad, plane = p.t.plane, n = p.t.n, dcm = p.t.dc_max;
  uint32_t *msg = smem;
  float *chan = reinterpret_cast<float *>(smem + dcm * plane);
  uint32_t *dec = reinterpret_cast<uint32_t *>(chan + n);
  const int n_round = (n + 31) & ~31;
  while (true) {
    if (tid == 0) s_frame = next_frame(p);
    __syncthreads();
    const int f = s_frame;
    if (f < 0) break;
    const float *in = p.in + (size_t)(p.sel ? f * p.n_cand + __ldg(p.sel + f) : f) * p.t.n_tx;
    for (int v = tid; v < n; v += T) chan[v] = v < p.t.punct ? 0.0f : load_channel_llr(in, v - p.t.punct, p.in_is_lr);
    for (int i = tid; i < dcm * plane; i += T) msg[i] = 0u;
    __syncthreads();
    int ret = p.iters + (p.iters < p.max_iter), last_t = 0;
    bool latched = false;
    for (int t = 0; t < p.iters; t++) {
      last_t = t;
      uint32_t *dcur = dec + (t & 1) * p.words_n;
      for (int v = tid; v < n_round; v += T) {
        uint32_t bit = 0;
        if (v < n) {
          const uint16_t *ad = p.t.vn_addr + (size_t)v * p.t.dv_max;
          switch (__ldg(p.t.vn_deg + v)) {
            case 0: bit = (chan[v] > 0.0f) ? 0u : 1u; break;
            KML_MS_VN_CASE(1) KML_MS_VN_CASE(2) KML_MS_VN_CASE(3) KML_MS_VN_CASE(4) KML_MS_VN_CASE(5) KML_MS_VN_CASE(6)
            KML_MS_VN_CASE(7) KML_MS_VN_CASE(8) KML_MS_VN_CASE(9) KML_MS_VN_CASE(10) KML_MS_VN_CASE(11) KML_MS_VN_CASE(12)
            KML_MS_VN_CASE(13) KML_MS_VN_CASE(14) KML_MS_VN_CASE(15) KML_MS_VN_CASE(16)
            default: break;
          }
        }
        const uint32_t word = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) dcur[v >> 5] = word;
      }
      __syncthreads();
      uint32_t xall = 0;
      for (int slot = tid; slot < mpad; slot += T) {
        uint32_t x = 0;
        switch (__ldg(p.t.cn_deg + slot)) {
          KML_MS_CN_CASE(1) KML_MS_CN_CASE(2) KML_MS_CN_CASE(3) KML_MS_CN_CASE(4) KML_MS_CN_CASE(5) KML_MS_CN_CASE(6)
          KML_MS_CN_CASE(7) KML_MS_CN_CASE(8) KML_MS_CN_CASE(9) KML_MS_CN_CASE(10) KML_MS_CN_CASE(11) KML_MS_CN_CASE(12)
          KML_MS_CN_CASE(13) KML_MS_CN_CASE(14) KML_MS_CN_CASE(15) KML_MS_CN_CASE(16)
          default: break;
        }
        xall |= x;
      }
      const int any_fail = __syncthreads_or((int)(xall & 1u));
      if (!any_fail && !latched) {
        latched = true;
        ret = t + (t < p.max_iter);
        for (int w = tid; w < p.words_n; w += T) p.out_bits[(size_t)f * p.words_n + w] = dcur[w];
        if (p.early_exit) break;
      }
    }
    if (!latched) {
      const uint32_t *dl = dec + (last_t & 1) * p.words_n;
      for (int w = tid; w < p.words_n; w += T) p.out_bits[(size_t)f * p.words_n + w] = dl[w];
    }
    if (tid == 0) p.out_ret[f] = ret;
  }
}
#undef KML_MS_VN_CASE
#undef KML_MS_CN_CASE

// ---------------------------------------------------------------------------------------------------------------
// algorithm = 2: the same normalised min-sum with fp16 messages, TWO frames per 32-bit shared-memory word (half2):
// HADD2 / HMNMX2 / HSET2 and the bit logic act on both frames at once, so instructions and shared-memory traffic per
// frame halve.  (3,6)-regular codes only; other graphs use algorithm 1.  Mantissa bit 0 of each half carries the
// posterior decision of that frame's variable.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }

template <int VPT, int CPT, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) ms2_regular_kernel(const DecParams p) {
  extern __shared__ uint32_t msg[];
  __shared__ int s_pair[2];
  const int tid = threadIdx.x;
  constexpr int plane = CPT * T + 1;
  uint32_t va[VPT][3];
#pragma unroll
  for (int j = 0; j < VPT; j++)
#pragma unroll
    for (int k = 0; k < 3; k++) va[j][k] = p.t.vn_addr[(j * T + tid) * 3 + k];
  const __half2 alpha2 = __float2half2_rn(p.alpha), clip2 = __float2half2_rn(kLlrClip), zero2 = __float2half2_rn(0.0f);
  const __half2 nbeta2 = __float2half2_rn(-p.beta);
  const bool has_offset = p.beta != 0.0f;

  while (true) {
    if (tid == 0) {  // a work item = two consecutive entries of the frame queue (the last one may be single: fb = fa)
      const int nB = frame_count(p), fp = (int)atomicAdd(p.work_counter, 1u);
      const int ia = 2 * fp, ib = min(2 * fp + 1, nB - 1);
      s_pair[0] = ia < nB ? frame_at(p, ia) : -1;
      s_pair[1] = ia < nB ? frame_at(p, ib) : -1;
    }
    __syncthreads();
    const int fa = s_pair[0], fb = s_pair[1];
    if (fa < 0) break;
    const float *ina = p.in + (size_t)(p.sel ? fa * p.n_cand + __ldg(p.sel + fa) : fa) * p.t.n_tx;
    const float *inb = p.in + (size_t)(p.sel ? fb * p.n_cand + __ldg(p.sel + fb) : fb) * p.t.n_tx;
    __half2 ch[VPT];
#pragma unroll
    for (int j = 0; j < VPT; j++)
      ch[j] = __floats2half2_rn(load_channel_llr(ina, j * T + tid, p.in_is_lr), load_channel_llr(inb, j * T + tid, p.in_is_lr));
    for (int i = tid; i < 6 * plane; i += T) msg[i] = 0u;
    __syncthreads();
    uint32_t bits_a = 0, bits_b = 0, lat_a = 0, lat_b = 0;
    const int ret_full = p.iters + (p.iters < p.max_iter);
    int ret_a = ret_full, ret_b = ret_full;
    bool done_a = false, done_b = false;
    for (int t = 0; t < p.iters; t++) {
      bits_a = bits_b = 0;
#pragma unroll
      for (int j = 0; j < VPT; j++) {
        const __half2 x0 = u2h(msg[va[j][0]]), x1 = u2h(msg[va[j][1]]), x2 = u2h(msg[va[j][2]]);
        const __half2 total = __hadd2(__hadd2(__hadd2(ch[j], x0), x1), x2);
        const uint32_t pb = __hle2_mask(total, zero2) & 0x00010001u;  // decision 1 unless total > 0 (tie → 1)
        bits_a |= (pb & 1u) << j;
        bits_b |= (pb >> 16) << j;
        msg[va[j][0]] = (h2u(__hsub2(total, x0)) & 0xFFFEFFFEu) | pb;
        msg[va[j][1]] = (h2u(__hsub2(total, x1)) & 0xFFFEFFFEu) | pb;
        msg[va[j][2]] = (h2u(__hsub2(total, x2)) & 0xFFFEFFFEu) | pb;
      }
      __syncthreads();
      uint32_t fail = 0;
#pragma unroll
      for (int j = 0; j < CPT; j++) {
        const int slot = j * T + tid;
        uint32_t w[6], x = 0;
        __half2 a[6], m1 = __float2half2_rn(60000.0f), m2 = m1;
#pragma unroll
        for (int k = 0; k < 6; k++) {
          w[k] = msg[k * plane + slot];
          x ^= w[k];
          a[k] = u2h(w[k] & 0x7FFE7FFEu);
          m2 = __hmin2(m2, __hmax2(m1, a[k]));
          m1 = __hmin2(m1, a[k]);
        }
        fail |= x;
        // max(alpha m - beta, 0) clipped: one HFMA2 (FMA pipe) + max + min per minimum
        __half2 t1 = __hfma2(alpha2, m1, nbeta2), t2 = __hfma2(alpha2, m2, nbeta2);
        if (has_offset) { t1 = __hmax2(t1, zero2); t2 = __hmax2(t2, zero2); }  // (uniform: the normalised rule never goes negative)
        const uint32_t s1 = h2u(__hmin2(t1, clip2)), s2 = h2u(__hmin2(t2, clip2));
#pragma unroll
        for (int k = 0; k < 6; k++) {
          const uint32_t eq = __heq2_mask(a[k], m1);
          msg[k * plane + slot] = ((s2 & eq) | (s1 & ~eq)) | ((x ^ w[k]) & 0x80008000u);
        }
      }
      const int fail_a = __syncthreads_or((int)(fail & 1u));
      const int fail_b = __syncthreads_or((int)((fail >> 16) & 1u));
      if (!fail_a && !done_a) { done_a = true; lat_a = bits_a; ret_a = t + (t < p.max_iter); }
      if (!fail_b && !done_b) { done_b = true; lat_b = bits_b; ret_b = t + (t < p.max_iter); }
      if (done_a && done_b && p.early_exit) break;
    }
    if (!done_a) lat_a = bits_a;
    if (!done_b) lat_b = bits_b;
#pragma unroll
    for (int j = 0; j < VPT; j++) {
      const uint32_t wa = __ballot_sync(0xffffffffu, (lat_a >> j) & 1u), wb = __ballot_sync(0xffffffffu, (lat_b >> j) & 1u);
      if ((tid & 31) == 0) {
        p.out_bits[(size_t)fa * p.words_n + ((j * T + tid) >> 5)] = wa;
        if (fb != fa) p.out_bits[(size_t)fb * p.words_n + ((j * T + tid) >> 5)] = wb;
      }
    }
    if (tid == 0) {
      p.out_ret[fa] = ret_a;
      if (fb != fa) p.out_ret[fb] = ret_b;
    }
  }
}

}  // namespace

dec_kernel_t minsum_kernel_of(DecKernelKind k, int alg) {
  if (alg == 2) {  // fp16 x 2 frames per word: regular codes only, otherwise the fp32 min-sum kernels below
#ifdef KML_TUNING
    const char *e = tuning_knob("KML_DEC_MINB");
    const int b = e ? atoi(e) : 3;
    if (k == DEC_REG_6_3 && b == 4) return ms2_regular_kernel<6, 3, 384, 4>;
    if (k == DEC_REG_6_3 && b == 2) return ms2_regular_kernel<6, 3, 384, 2>;
#endif
    if (k == DEC_REG_6_3) return ms2_regular_kernel<6, 3, 384, 3>;
    if (k == DEC_REG_12_6) return ms2_regular_kernel<12, 6, 672, 1>;
  }
  switch (k) {
    case DEC_REG_6_3: return ms_regular_kernel<6, 3, 384, 3>;
    case DEC_REG_12_6: return ms_regular_kernel<12, 6, 672, 1>;
    case DEC_GEN_4_8: return ms_generic_kernel<4, 8>;
    case DEC_GEN_9_10: return ms_generic_kernel<9, 10>;
    case DEC_GEN_16_32: return ms_generic_kernel<16, 16>;
  }
  return nullptr;
}

}  // namespace kml
