// Flooding sum-product LDPC decoder for sm_100a — the hot kernel of the link path.
//
// Replaces BinaryLDPCCodec::Decoder (lib/lab/src/binaryldpccodec.cc:165-278) and Binary5GLDPCCodec::Decoder
// (lib/lab/src/binary5gldpccodec.cc:112-232): same schedule (variable phase → hard decision → syndrome test →
// check phase), same clipping of check messages to [1e-12, 1-1e-12], same return value.
//
// Number representation (fp32, chosen so that no quantity the reference keeps in fp64 is lost):
//   * check→variable message: likelihood ratio  x = P0/P1  in [1e-12, 1e12]  (the clipped pair (c2v0, 1-c2v0));
//     variable nodes multiply ratios: extrinsic_e = L_ch * prod_{e' != e} x_e'  (reference: alpha/beta products with
//     per-step renormalisation, which cancels in the ratio).
//   * variable→check message: (hard bit h, small probability s <= 0.5): P(bit = h) = 1 - s.  s is computed directly
//     as min(x,1)/(1+x) (in the regular kernel in the division-free form min(post, x_k)/(x_k + post) with post the
//     product of all ratios), so probabilities as small as 1e-36 keep full relative precision — the fp32 analogue of the
//     reference carrying both members of the pair in fp64.  A word packs s, the extrinsic hard bit in the SIGN bit
//     (so the check node reads s as |word| — a free operand modifier on FFMA2/FFMA/FMNMX) and the POSTERIOR hard
//     decision of the variable in the mantissa LSB (a 1-ulp perturbation of s), so the check phase can evaluate the
//     syndrome of the current decisions without a second gather.
//   * check node: the reference's 2-state trellis is, in this representation, s_ab = s_a + s_b - 2 s_a s_b on the small
//     probabilities (all terms positive → no cancellation) and XOR on the hard bits; forward/backward partial
//     combinations give every extrinsic output in 3(d-2) combines.
//
// One CTA decodes one frame at a time (persistent CTAs, dynamic frame queue).  Per iteration and edge the regular kernel
// makes 4 shared-memory word accesses (16 B), 2 MUFU.RCP and ~19.5 issue slots; nothing but the channel ratios
// (4 B/variable, coalesced) and the packed decisions (1 bit/variable) touches HBM.
//
// Three kernels: bp_regular_kernel ((3,6)-regular PEG codes, everything unrolled), bp_qc_kernel (quasi-cyclic codes with a
// compile-time plan: 5G BG2), bp_generic_kernel (any other Tanner graph, run-time work lists).
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <utility>

#include <cuda_fp16.h>

#include "kml_internal.h"
#include "kml_kernels.cuh"
#include "bp_minsum_nodes.cuh"

namespace kml {
namespace {

constexpr float kLlrClip = 27.631021f;  // ln((1-1e-12)/1e-12)

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <int DIAG>
__device__ __forceinline__ float rcp_d(float x) { return (DIAG & 2) ? 2.0f - x : rcp_approx(x); }

// s_ab = s_a + s_b - 2 s_a s_b
// (measured: the 3-instruction form with two independent operands beats fmaf(a, fmaf(-2, b, 1), b) — the kernel is
// sensitive to dependent-FFMA latency, not only to instruction count)
__device__ __forceinline__ float sp_combine(float a, float b) { return fmaf(-2.0f * a, b, a + b); }
// same with t_b = 1 - 2 s_b precomputed
__device__ __forceinline__ float sp_chain(float acc, float s, float t) { return fmaf(acc, t, s); }

// v2c word: s = min(P0, P1) = min(ext, 1) / (1 + ext); bit 31 = extrinsic hard decision (P1 > P0, i.e. ext < 1), taken
// from the sign of (ext - 1) — one FADD on the FMA pipe instead of FSETP + SEL on the (binding) ALU pipe; `pb` = the
// variable's posterior decision (0 / 1), which replaces the mantissa LSB of s.  Two LOP3.
// (inline PTX pins the association: left to itself ptxas emits three LOP3)
__device__ __forceinline__ uint32_t v2c_pack(float sgn, float s, uint32_t pb) {
  uint32_t w;
  asm("{.reg .b32 t; lop3.b32 t, %1, 0x80000000, %3, 0xEA; lop3.b32 %0, %2, 0xfffffffe, t, 0xEA;}"
      : "=r"(w) : "r"(__float_as_uint(sgn)), "r"(__float_as_uint(s)), "r"(pb));
  return w;
}
__device__ __forceinline__ uint32_t v2c_word(float ext, uint32_t pb) {
  const float s = fminf(ext, 1.0f) * rcp_approx(1.0f + ext);
  return v2c_pack(ext - 1.0f, s, pb);
}

// sign bit of (x ^ w) as a predicate in ONE instruction (LOP3.LUT with a predicate destination) — the C form
// `(int)(x ^ w) < 0` compiles to LOP3 + ISETP.  LUT of (a ^ b) & c = (0xF0 ^ 0xCC) & 0xAA = 0x28.
#pragma nv_diag_suppress 550  // (`d` is the instruction's mandatory 32-bit destination)
__device__ __forceinline__ bool sign_xor(uint32_t x, uint32_t w) {
  uint32_t d, r;
  asm("{.reg .pred pp; lop3.or.b32 %0|pp, %2, %3, 0x80000000, 0x28, 0; selp.u32 %1, 1, 0, pp;}"
      : "=r"(d), "=r"(r) : "r"(x), "r"(w));
  (void)d;
  return r != 0;
}
#pragma nv_diag_default 550

// c2v ratio P0/P1 from the small probability s and the output's hard decision (sign bit of xw).  The clip of c2v0 to
// [1e-12, 1-1e-12] (binaryldpccodec.cc:259-263) acts on the small side only.  q = (1-s)/s is the ratio for hard = 0;
// hard = 1 takes a second (predicated) reciprocal — cheaper on this ALU-bound kernel than selecting numerator and
// denominator (2 FSEL): the XU pipe has head-room (profiles/r1b).
__device__ __forceinline__ float c2v_ratio(float s, uint32_t x, uint32_t w) {
  s = fmaxf(s, kSmallProbF);
  float q = (1.0f - s) * rcp_approx(s);
  if (sign_xor(x, w)) q = rcp_approx(q);
  return q;
}

// ln(syndrom_soft[r]) (binaryldpccodec.cc:274): syndrom_soft = P(check satisfied) = the row's small probability when its
// hard bits have odd parity, 1 - s otherwise — taken as log1p(-s) so that rows the decoder is sure of still contribute
// their ~ -s (the reference sums them in fp64)
__device__ __forceinline__ float soft_log(uint32_t odd, float s) {
  // (graphs with variable degree above 3 can still underflow s to 0 in fp32 where the reference's fp64 holds 1e-40…:
  //  such a row counts as ln(1.2e-38) = -87.3 instead of -inf)
  if (odd) return __logf(fmaxf(s, 1.18e-38f));
  return s < 0.01f ? -s * fmaf(s, fmaf(s, 0.33333334f, 0.5f), 1.0f) : __logf(1.0f - s);
}

__device__ __forceinline__ float load_channel_ratio(const float *in, int idx, int in_is_lr) {
  float v = __ldg(in + idx);
  if (!in_is_lr) v = __expf(fminf(fmaxf(v, -kLlrClip), kLlrClip));
  return fminf(fmaxf(v, kLrMin), kLrMax);
}

// Blackwell packed fp32 (PTX mul/add/fma.rn.f32x2 → SASS FMUL2 / FADD2 / FFMA2, sm_100+).  ptxas folds {x, x} and
// {b.y, b.x} operand constructions into the instruction's .F32 (broadcast) and .LO_HI (swap) selectors — no moves.
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 splat(float c) { return make_float2(c, c); }

// ---------------------------------------------------------------------------------------------------------------
// (3,6)-regular codes (PEG2304, PEG8064): NV <= VPT * T variables, NV / 2 <= CPT * T checks (the last round of a
// thread may be empty when the tiling is not exact), everything unrolled,
// edge addresses and channel ratios resident in registers.
// ---------------------------------------------------------------------------------------------------------------
//
// SOFT = the syndrom_soft sum of the soft-syndrome metric ([ldpc] metric_type = true) is produced; a compile-time switch
// because even the skipped branch costs 7 issue slots per check in this issue-bound loop.
//
// Two shared-memory layouts (layout_opt.cpp makes the variable-node gathers conflict free for either):
//   ROWM = false ("planar")   : edge (row slot, k) at word k * (M + 1) + slot — check nodes issue 6 LDS.32 + 6 STS.32
//   ROWM = true  ("row-major"): at word 6 * slot + phys(k) — the six words of a check are contiguous and move as
//                               3 LDS.64 + 3 STS.64 (a half-warp covers 16 distinct even banks: conflict free), which
//                               takes one issue slot per edge-iteration out of this issue-bound kernel.
// DIAG (KML_DEC_DIAG, timing experiments only — the RESULTS ARE WRONG): bit 0 drops the two barriers of an iteration,
// bit 1 replaces every MUFU.RCP of the loop by an FADD; what the launch time loses says what the barrier / the XU pipe
// cost (profiles/r1_decoder_ablation.txt).
template <int VPT, int CPT, int T, int MINB, bool PACK = true, int RATIO = 2, bool ROWM = false, bool SOFT = false,
          int NV = VPT * T, int DIAG = 0>
__global__ void __launch_bounds__(T, MINB) bp_regular_kernel(const DecParams p) {
  static_assert(T % 32 == 0 && NV % 64 == 0 && NV <= VPT * T && NV / 2 <= CPT * T, "warps must line up with the 32-node groups");
  constexpr int NC = NV / 2;
  constexpr bool kExact = NV == VPT * T && NC == CPT * T;
  extern __shared__ __align__(16) uint32_t msg[];
  __shared__ int s_frame;
  const int tid = threadIdx.x;
  constexpr int mpad = NC + 1;  // planar: words between the k planes, bank = (slot + k) mod 32
  constexpr int n_words = ROWM ? 6 * NC : 6 * mpad;
  static_assert(!ROWM || PACK, "row-major layout is implemented for the packed path");

  uint32_t va[VPT][3];
#pragma unroll
  for (int j = 0; j < VPT; j++) {
    const int v = j * T + tid;
#pragma unroll
    for (int k = 0; k < 3; k++) va[j][k] = (kExact || v < NV) ? p.t.vn_addr[v * 3 + k] : 0;
  }

  while (true) {
    if (tid == 0) s_frame = next_frame(p);
    __syncthreads();
    const int f = s_frame;
    if (f < 0) break;
    const float *in = p.in + (size_t)(p.sel ? f * p.n_cand + __ldg(p.sel + f) : f) * p.t.n_tx;
    float ch[VPT];
#pragma unroll
    for (int j = 0; j < VPT; j++) ch[j] = (kExact || j * T + tid < NV) ? load_channel_ratio(in, j * T + tid, p.in_is_lr) : 1.0f;
    constexpr bool kPair = PACK && kExact && VPT % 2 == 0;  // two-variable variable nodes (below)
    if (!kPair) {
      for (int i = tid; i < n_words; i += T) msg[i] = 0x3f800000u;  // InitMsg: c2v = (0.5, 0.5) → ratio 1
      __syncthreads();
    }  // (kPair: the first variable phase takes x = 1 instead of reading, and writes every word — no initialisation pass,
       //  which is a fifth of the work of a frame that converges in four iterations)

    uint32_t bits = 0, latched_bits = 0;
    int ret = p.iters + (p.iters < p.max_iter);
    bool latched = false;
    float soft = 0.0f, soft_out = 0.0f;
    for (int t = 0; t < p.iters; t++) {
      // ---- variable nodes (binaryldpccodec.cc:177-212)
      bits = 0;
      uint32_t nbits = 0;  // complemented decisions (two-variable path)
      if (kPair) {
       auto vn2 = [&](auto first_tag) {
        constexpr bool kFirst = decltype(first_tag)::value;  // iteration 0: every c2v is InitMsg's ratio 1
        // Two variables per step, so that EVERY operation is packed.  With post = ch x0 x1 x2 the extrinsic ratio is
        // e_k = post / x_k, hence  s_k = min(e_k, 1) / (1 + e_k) = min(post, x_k) / (x_k + post)  and the hard bit is
        // the sign of post - x_k: the three e_k are never formed.  fp32 range: post overflows / underflows only when
        // every e_k is beyond 1e±26, where s_k = 0 is the right limit.
#pragma unroll
        for (int j = 0; j < VPT; j += 2) {
          float2 x[3];
#pragma unroll
          for (int k = 0; k < 3; k++)
            x[k] = kFirst ? splat(1.0f)
                          : make_float2(__uint_as_float(msg[va[j][k]]), __uint_as_float(msg[va[j + 1][k]]));
          const float2 post = mul2(mul2(make_float2(ch[j], ch[j + 1]), x[0]), mul2(x[1], x[2]));
          // The messages carry the COMPLEMENT of the posterior decision: it is the sign bit of 1 - post (set iff
          // post > 1, a tie gives +0 → decision 1 like the reference), one shift instead of FSETP + SEL on the
          // half-rate ALU pipe; a row has six edges, so the XOR of the complements is the XOR of the decisions.
          const float2 om = fma2(post, splat(-1.0f), splat(1.0f));
          const uint32_t ba = __float_as_uint(om.x) >> 31, bb = __float_as_uint(om.y) >> 31;
          nbits += (ba << j) + (bb << (j + 1));
#pragma unroll
          for (int k = 0; k < 3; k++) {
            const float2 den = add2(x[k], post), sgn = fma2(x[k], splat(-1.0f), post);
            const float2 s2 = mul2(make_float2(fminf(post.x, x[k].x), fminf(post.y, x[k].y)),
                                   make_float2(rcp_d<DIAG>(den.x), rcp_d<DIAG>(den.y)));
            msg[va[j][k]] = v2c_pack(sgn.x, s2.x, ba);
            msg[va[j + 1][k]] = v2c_pack(sgn.y, s2.y, bb);
          }
        }
       };
       if (t == 0) vn2(std::true_type{});
       else vn2(std::false_type{});
      } else
#pragma unroll
      for (int j = 0; j < VPT; j++) {
        if (!kExact && j * T + tid >= NV) continue;  // empty last round (warp-uniform: NV is a multiple of 32)
        const float x0 = __uint_as_float(msg[va[j][0]]);
        const float x1 = __uint_as_float(msg[va[j][1]]);
        const float x2 = __uint_as_float(msg[va[j][2]]);
        uint32_t w0, w1, w2, bit;
        if (PACK) {
          // With post = ch x0 x1 x2 the extrinsic ratio is e_k = post / x_k, so  s_k = min(e_k, 1) / (1 + e_k)
          // = min(post, x_k) / (x_k + post)  and the hard bit is the sign of post - x_k: the three e_k are never
          // formed (3 issue slots for the products instead of 4, and the edge-2 add / subtract are one FFMA2).  fp32 range: post overflows / underflows only when
          // every e_k is beyond 1e±26, where s_k = 0 is the right limit.
          const float post = (ch[j] * x0) * (x1 * x2);  // scalar: pairing ch with a loaded value would cost a move
          bit = (post > 1.0f) ? 0u : 1u;
          const uint32_t pb = bit;
          const float2 x01 = make_float2(x0, x1);
          const float2 den01 = add2(x01, splat(post)), sgn01 = fma2(x01, splat(-1.0f), splat(post));
          const float2 ds2 = fma2(splat(x2), make_float2(1.0f, -1.0f), splat(post));  // (x2 + post, post - x2)
          const float2 s01 = mul2(make_float2(fminf(post, x0), fminf(post, x1)),
                                  make_float2(rcp_approx(den01.x), rcp_approx(den01.y)));
          const float s2 = fminf(post, x2) * rcp_approx(ds2.x);
          w0 = v2c_pack(sgn01.x, s01.x, pb);
          w1 = v2c_pack(sgn01.y, s01.y, pb);
          w2 = v2c_pack(ds2.y, s2, pb);
        } else {
          const float a = ch[j] * x0, b = ch[j] * x1;
          const float e2 = a * x1, e1 = a * x2, e0 = b * x2;
          const float post = e2 * x2;
          bit = (post > 1.0f) ? 0u : 1u;  // alpha0 > alpha1 ? 0 : 1 (tie → 1)
          const uint32_t pb = bit;
          w0 = v2c_word(e0, pb);
          w1 = v2c_word(e1, pb);
          w2 = v2c_word(e2, pb);
        }
        bits |= bit << j;
        msg[va[j][0]] = w0;
        msg[va[j][1]] = w1;
        msg[va[j][2]] = w2;
      }
      if (kPair) bits = ~nbits;
      if (!(DIAG & 1)) __syncthreads();
      // ---- check nodes + syndrome of the decisions just made (binaryldpccodec.cc:217-275)
      int fail = 0;
      const float soft_before = soft;  // syndrom_soft_ as the reference holds it when it leaves at this iteration
      soft = 0.0f;
#pragma unroll
      for (int j = 0; j < CPT; j++) {
        const int slot = j * T + tid;
        if (!kExact && slot >= NC) continue;
        uint32_t w[6];
        float s[6];
        uint32_t x = 0;
        // row-major: the check node is indifferent to the order of its edges, so the physical pairs (0,1) (2,3) (4,5)
        // are taken as the logical pairs (1,4) (2,3) (5,0) the packed arithmetic produces its outputs in
        uint2 *row = reinterpret_cast<uint2 *>(msg + 6 * slot);
        if (ROWM) {
          const uint2 p0 = row[0], p1 = row[1], p2 = row[2];
          w[1] = p0.x; w[4] = p0.y; w[2] = p1.x; w[3] = p1.y; w[5] = p2.x; w[0] = p2.y;
        } else {
#pragma unroll
          for (int k = 0; k < 6; k++) w[k] = msg[k * mpad + slot];
        }
#pragma unroll
        for (int k = 0; k < 6; k++) {
          x ^= w[k];
          s[k] = fabsf(__uint_as_float(w[k]));
        }
        fail |= (int)(x & 1u);
        float so[6], sall = 0.0f;
        if (PACK) {
          // prefix and suffix chains advance in lock step in the two halves of one FFMA2:
          //   C_i = (pre_i, suf_{6-i}),  C_{i+1} = C_i (t_i, t_{5-i}) + (s_i, s_{5-i}),  C_1 = (s_0, s_5)
          // and the outputs pair up as (so_1, so_4) = C_1 ⊕ swap(C_4), (so_2, so_3) = C_2 ⊕ swap(C_3), (so_5, so_0) = C_5.
          const float2 s05 = make_float2(s[0], s[5]), s14 = make_float2(s[1], s[4]), s23 = make_float2(s[2], s[3]);
          const float2 t14 = fma2(s14, splat(-2.0f), splat(1.0f)), t23 = fma2(s23, splat(-2.0f), splat(1.0f));
          const float2 c1 = s05;
          const float2 c2 = fma2(c1, t14, s14);
          const float2 c3 = fma2(c2, t23, s23);
          const float2 c4 = fma2(c3, make_float2(t23.y, t23.x), make_float2(s23.y, s23.x));
          const float2 c5 = fma2(c4, make_float2(t14.y, t14.x), make_float2(s14.y, s14.x));
          const float2 c4s = make_float2(c4.y, c4.x), c3s = make_float2(c3.y, c3.x);
          const float2 o14 = fma2(mul2(c1, splat(-2.0f)), c4s, add2(c1, c4s));
          const float2 o23 = fma2(mul2(c2, splat(-2.0f)), c3s, add2(c2, c3s));
          so[1] = o14.x; so[4] = o14.y; so[2] = o23.x; so[3] = o23.y; so[5] = c5.x; so[0] = c5.y;
          if (SOFT) sall = sp_chain(c5.x, s[5], fmaf(-2.0f, s[5], 1.0f));
          // ratios two at a time: clip, 1 - s, reciprocal, product
          const int ka[3] = {1, 2, 5}, kb[3] = {4, 3, 0};
#pragma unroll
          for (int i = 0; i < 3; i++) {
            const float2 sc = make_float2(fmaxf(so[ka[i]], kSmallProbF), fmaxf(so[kb[i]], kSmallProbF));
            const float2 big = fma2(sc, splat(-1.0f), splat(1.0f));
            const bool ha = sign_xor(x, w[ka[i]]), hb = sign_xor(x, w[kb[i]]);
            float2 q;
            if (RATIO == 0) {         // second, predicated reciprocal for hard = 1 (XU pipe)
              q = mul2(big, make_float2(rcp_approx(sc.x), rcp_approx(sc.y)));
              if (ha) q.x = rcp_approx(q.x);
              if (hb) q.y = rcp_approx(q.y);
            } else if (RATIO == 1) {  // select numerator / denominator (ALU pipe)
              const float2 num = make_float2(ha ? sc.x : big.x, hb ? sc.y : big.y);
              const float2 den = make_float2(ha ? big.x : sc.x, hb ? big.y : sc.y);
              q = mul2(num, make_float2(rcp_d<DIAG>(den.x), rcp_d<DIAG>(den.y)));
            } else if (RATIO == 3) {  // selects + ONE reciprocal for the pair: r = 1 / (den_a den_b), q_a = num_a den_b r
              const float2 num = make_float2(ha ? sc.x : big.x, hb ? sc.y : big.y);
              const float2 den = make_float2(ha ? big.x : sc.x, hb ? big.y : sc.y);  // >= 1e-12 each: product is normal
              const float r = rcp_approx(den.x * den.y);
              q = mul2(mul2(num, make_float2(den.y, den.x)), splat(r));
            } else {                  // one of each: balances the XU and ALU pipes
              const float numb = hb ? sc.y : big.y, denb = hb ? big.y : sc.y;
              q = mul2(make_float2(big.x, numb), make_float2(rcp_approx(sc.x), rcp_approx(denb)));
              if (ha) q.x = rcp_approx(q.x);
            }
            if (ROWM) {
              row[i] = make_uint2(__float_as_uint(q.x), __float_as_uint(q.y));
            } else {
              msg[ka[i] * mpad + slot] = __float_as_uint(q.x);
              msg[kb[i] * mpad + slot] = __float_as_uint(q.y);
            }
          }
        } else {
          float tt[6];
#pragma unroll
          for (int k = 0; k < 6; k++) tt[k] = fmaf(-2.0f, s[k], 1.0f);
          float pre[6], suf[6];  // pre[k] = s_0 ⊕ … ⊕ s_{k-1}, suf[k] = s_k ⊕ … ⊕ s_5
          pre[1] = s[0];
#pragma unroll
          for (int k = 2; k < 6; k++) pre[k] = sp_chain(pre[k - 1], s[k - 1], tt[k - 1]);
          suf[5] = s[5];
#pragma unroll
          for (int k = 4; k >= 1; k--) suf[k] = sp_chain(suf[k + 1], s[k], tt[k]);
          so[0] = suf[1];
          so[5] = pre[5];
#pragma unroll
          for (int k = 1; k < 5; k++) so[k] = sp_combine(pre[k], suf[k + 1]);
#pragma unroll
          for (int k = 0; k < 6; k++) msg[k * mpad + slot] = __float_as_uint(c2v_ratio(so[k], x, w[k]));
          if (SOFT) sall = sp_chain(pre[5], s[5], tt[5]);
        }
        if (SOFT)  // syndrom_soft[r] = P(check satisfied) = row_head.alpha[0] (binaryldpccodec.cc:274)
          soft += soft_log(x >> 31, sall);
      }
      const int any_fail = (DIAG & 1) ? 1 : __syncthreads_or(fail);
      if (!any_fail && !latched) {
        latched = true;
        latched_bits = bits;
        soft_out = soft_before;
        ret = t + (t < p.max_iter);
        if (p.early_exit) break;  // the reference leaves BEFORE the check phase; its c2v are never read again
      }
    }
    if (!latched) {
      latched_bits = bits;
      soft_out = soft;
    }
    const int lane = tid & 31;
#pragma unroll
    for (int j = 0; j < VPT; j++) {
      const uint32_t word = __ballot_sync(0xffffffffu, (latched_bits >> j) & 1u);
      if (lane == 0 && (kExact || j * T + tid < NV)) p.out_bits[(size_t)f * p.words_n + ((j * T + tid) >> 5)] = word;
    }
    if (tid == 0) p.out_ret[f] = ret;
    if (SOFT) {
      // NOTE: when the reference leaves at t = 0 its syndrom_soft_ is stale (left over from the previous call);
      // here that case reports 0.
      double sd = (double)soft_out;  // per-thread partial sums are short; everything across threads is fp64
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
      if (lane == 0) atomicAdd(p.out_soft + f, sd);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Any Tanner graph with column degree <= DV and row degree <= DC (5G BG2, irregular codes, odd sizes).
// * Work is handed out in WARP ITEMS: 32 consecutive variables (or row slots).  The host balances the items over the
//   CTA's warps by node degree (longest-processing-time first; kml_api.cu) — a tid-strided loop leaves the warps that
//   drew the degree-9 columns of BG2 with 2x the work of the others.
// * The node update is dispatched on the node's EXACT degree (warp-uniform for quasi-cyclic codes: the 32 nodes of an
//   item share a degree), so no work is spent on padding.
// * Node arithmetic is packed like the regular kernel's: prefix and suffix products / combinations advance in the two
//   halves of one FMUL2 / FFMA2, outputs pair up as (k, D-1-k).
// * Variable nodes above degree 3 rely on fp32 saturation (inf / 0) instead of clamping every partial product: a
//   saturated extrinsic ratio gives s = 0 exactly like a clamped one gives s < 1e-36, and the one pathological
//   combination (inf x 0 = NaN: >= 4 saturated messages each way on one variable) is absorbed by the check node's
//   clip (fmaxf(NaN, 1e-12) = 1e-12) — the clamped product was equally arbitrary there.
// * Messages are row-major with an odd compile-time row stride RS (word 'slot * RS + k': conflict free for the check
//   nodes, every access an immediate offset from one base register); variable-node address lists are stored per item as
//   [edge k][lane] so a warp's 32 loads are one 64-byte line at an immediate offset.  Channel ratios are staged in
//   shared memory.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 swap2(float2 a) { return make_float2(a.y, a.x); }

// a[k] = shared-memory word address of the variable's k-th edge
template <int D>
__device__ __forceinline__ uint32_t vn_core(uint32_t *msg, const uint32_t *a, float ch) {
  float x[D];
#pragma unroll
  for (int k = 0; k < D; k++) x[k] = __uint_as_float(msg[a[k]]);
  // P[i] = (pre_i, suf_{D-i}):  pre_i = ch x_0 … x_{i-1},  suf_j = x_j … x_{D-1}
  float2 P[D + 1];
  P[0] = make_float2(ch, 1.0f);
#pragma unroll
  for (int i = 0; i < D; i++) P[i + 1] = mul2(P[i], make_float2(x[i], x[D - 1 - i]));
  const uint32_t bit = (P[D].x > 1.0f) ? 0u : 1u;  // alpha0 > alpha1 ? 0 : 1 (tie → 1)
  // extrinsic ratios  ext_k = pre_k suf_{k+1}:  (ext_k, ext_{D-1-k}) = P[k] * swap(P[D-1-k])
#pragma unroll
  for (int k = 0; 2 * k < D; k++) {
    const int kk = D - 1 - k;
    if (k < kk) {
      const float2 e = mul2(P[k], swap2(P[kk]));
      const float2 den = add2(e, splat(1.0f)), sgn = add2(e, splat(-1.0f));
      const float2 s2 = mul2(make_float2(fminf(e.x, 1.0f), fminf(e.y, 1.0f)),
                             make_float2(rcp_approx(den.x), rcp_approx(den.y)));
      msg[a[k]] = v2c_pack(sgn.x, s2.x, bit);
      msg[a[kk]] = v2c_pack(sgn.y, s2.y, bit);
    } else {
      msg[a[k]] = v2c_word(P[k].x * P[k].y, bit);
    }
  }
  return bit;
}

template <int D>
__device__ __forceinline__ uint32_t vn_node(uint32_t *msg, const uint16_t *ad, float ch) {
  uint32_t a[D];
#pragma unroll
  for (int k = 0; k < D; k++) a[k] = __ldg(ad + k * 32);
  return vn_core<D>(msg, a, ch);
}

// c2v ratios of a pair of outputs (select-based inversion, as the regular kernel's RATIO = 1)
__device__ __forceinline__ float2 c2v_ratio2(float2 so, uint32_t x, uint32_t wa, uint32_t wb) {
  const float2 sc = make_float2(fmaxf(so.x, kSmallProbF), fmaxf(so.y, kSmallProbF));
  const float2 big = fma2(sc, splat(-1.0f), splat(1.0f));
  const bool ha = sign_xor(x, wa), hb = sign_xor(x, wb);
  const float2 num = make_float2(ha ? sc.x : big.x, hb ? sc.y : big.y);
  const float2 den = make_float2(ha ? big.x : sc.x, hb ? big.y : sc.y);
  return mul2(num, make_float2(rcp_approx(den.x), rcp_approx(den.y)));
}

// returns the XOR of the row's words (bit 0 = syndrome of the current decisions, bit 31 = parity of the extrinsic hard
// bits); *s_all = small probability of the whole row (for syndrom_soft)
template <int D, bool SOFT>
__device__ __forceinline__ uint32_t cn_node(uint32_t *row, float *s_all) {
  uint32_t w[D], x = 0;
  float s[D];
#pragma unroll
  for (int k = 0; k < D; k++) {
    w[k] = row[k];
    x ^= w[k];
    s[k] = fabsf(__uint_as_float(w[k]));
  }
  if (D < 3) {  // degree 1: the only output is the neutral element (clipped); degree 2: the two inputs swap
#pragma unroll
    for (int k = 0; k < D; k++) row[k] = __float_as_uint(c2v_ratio(D == 1 ? 0.0f : s[1 - k], x, w[k]));
    if (SOFT) *s_all = D == 1 ? s[0] : sp_combine(s[0], s[1]);
    return x;
  }
  // C[j] = (pre_j, suf_{D-j}), j = 1 … D-1:  pre_j = s_0 ⊕ … ⊕ s_{j-1},  suf_j = s_j ⊕ … ⊕ s_{D-1};
  // C[j+1] = C[j] (t_j, t_{D-1-j}) + (s_j, s_{D-1-j})
  float2 C[D];
  C[1] = make_float2(s[0], s[D - 1]);
#pragma unroll
  for (int j = 1; j < D - 1; j++) {
    const float2 sj = make_float2(s[j], s[D - 1 - j]);
    C[j + 1] = fma2(C[j], fma2(sj, splat(-2.0f), splat(1.0f)), sj);
  }
  // outputs: (so_{D-1}, so_0) = C[D-1];  (so_k, so_{D-1-k}) = C[k] ⊕ swap(C[D-1-k])
  {
    const float2 q = c2v_ratio2(C[D - 1], x, w[D - 1], w[0]);
    row[D - 1] = __float_as_uint(q.x);
    row[0] = __float_as_uint(q.y);
  }
#pragma unroll
  for (int k = 1; 2 * k < D; k++) {
    const int kk = D - 1 - k;
    if (k < kk) {
      const float2 cs = swap2(C[kk]);
      const float2 so = fma2(mul2(C[k], splat(-2.0f)), cs, add2(C[k], cs));
      const float2 q = c2v_ratio2(so, x, w[k], w[kk]);
      row[k] = __float_as_uint(q.x);
      row[kk] = __float_as_uint(q.y);
    } else {
      row[k] = __float_as_uint(c2v_ratio(sp_combine(C[k].x, C[k].y), x, w[k]));
    }
  }
  if (SOFT) *s_all = sp_chain(C[D - 1].x, s[D - 1], fmaf(-2.0f, s[D - 1], 1.0f));
  return x;
}

// one case of the per-item dispatch: the whole run of `cnt` groups is processed inside the case, so the compare chain
// of the switch is paid once per item
#define KML_VN_CASE(D)                                                                      \
  case D:                                                                                   \
    if (D <= DV) {                                                                          \
      for (int r = 0; r < cnt; r++, v += 32, ad += 32 * dvm) {                              \
        const uint32_t bit = vn_node<(D <= DV ? D : 1)>(msg, ad, chan[v]);                  \
        const uint32_t word = __ballot_sync(0xffffffffu, bit);                              \
        if (lane == 0) dcur[g0 + r] = word;                                                 \
      }                                                                                     \
    }                                                                                       \
    break;
#define KML_CN_CASE(D)                                                                      \
  case D:                                                                                   \
    if (D <= DC) {                                                                          \
      for (int r = 0; r < cnt; r++, slot += 32) {                                           \
        float s_all = 0.0f;                                                                 \
        const uint32_t x = cn_node<(D <= DC ? D : 1), SOFT>(msg + slot * RS, &s_all);       \
        fail += (int)(x & 1u);                                                              \
        if (SOFT) soft += soft_log(x >> 31, s_all);                                         \
      }                                                                                     \
    }                                                                                       \
    break;
#define KML_VN_LANE_CASE(D)                                                                 \
  case D:                                                                                   \
    if (D <= DV) bit = vn_node<(D <= DV ? D : 1)>(msg, ad, ch);                             \
    break;
#define KML_CN_LANE_CASE(D)                                                                 \
  case D:                                                                                   \
    if (D <= DC) x = cn_node<(D <= DC ? D : 1), SOFT>(row, s_all);                          \
    break;

// groups whose 32 nodes do not share a degree (irregular non-quasi-cyclic graphs, the ragged last group): per-lane
// dispatch, kept out of line so the common path stays small
template <int DV>
__device__ __noinline__ uint32_t vn_lane_dispatch(uint32_t *msg, const uint16_t *ad, float ch, int deg) {
  uint32_t bit = 0;
  switch (deg) {
    case 0: bit = (ch > 1.0f) ? 0u : 1u; break;
    KML_VN_LANE_CASE(1) KML_VN_LANE_CASE(2) KML_VN_LANE_CASE(3) KML_VN_LANE_CASE(4) KML_VN_LANE_CASE(5) KML_VN_LANE_CASE(6)
    KML_VN_LANE_CASE(7) KML_VN_LANE_CASE(8) KML_VN_LANE_CASE(9) KML_VN_LANE_CASE(10) KML_VN_LANE_CASE(11)
    KML_VN_LANE_CASE(12) KML_VN_LANE_CASE(13) KML_VN_LANE_CASE(14) KML_VN_LANE_CASE(15) KML_VN_LANE_CASE(16)
    default: break;
  }
  return bit;
}
template <int DC, bool SOFT>
__device__ __noinline__ uint32_t cn_lane_dispatch(uint32_t *row, float *s_all, int deg) {
  uint32_t x = 0;
  switch (deg) {
    KML_CN_LANE_CASE(1) KML_CN_LANE_CASE(2) KML_CN_LANE_CASE(3) KML_CN_LANE_CASE(4) KML_CN_LANE_CASE(5) KML_CN_LANE_CASE(6)
    KML_CN_LANE_CASE(7) KML_CN_LANE_CASE(8) KML_CN_LANE_CASE(9) KML_CN_LANE_CASE(10) KML_CN_LANE_CASE(11)
    KML_CN_LANE_CASE(12) KML_CN_LANE_CASE(13) KML_CN_LANE_CASE(14) KML_CN_LANE_CASE(15) KML_CN_LANE_CASE(16)
    default: break;  // padding slot
  }
  return x;
}

constexpr int kGenericThreads = 384;  // 3 CTAs per SM at <= 56 registers

template <int DV, int DC, bool SOFT>
__global__ void __launch_bounds__(kGenericThreads, 3) bp_generic_kernel(const DecParams p) {
  static_assert(DV <= 16 && DC <= 16, "add switch cases");
  extern __shared__ uint32_t smem[];
  __shared__ int s_frame;
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
  constexpr int RS = DC | 1;  // odd row stride: lanes 0..31 of a check-node access hit 32 distinct banks
  const int n = p.t.n, n_words = p.t.m_pad * RS;
  uint32_t *msg = smem;                                  // [m_pad][RS]
  float *chan = reinterpret_cast<float *>(smem + n_words);  // [n]
  uint32_t *dec = reinterpret_cast<uint32_t *>(chan + n);      // [2][words_n] decisions, double buffered

  while (true) {
    if (tid == 0) s_frame = next_frame(p);
    __syncthreads();
    const int f = s_frame;
    if (f < 0) break;
    const float *in = p.in + (size_t)(p.sel ? f * p.n_cand + __ldg(p.sel + f) : f) * p.t.n_tx;
    for (int v = tid; v < n; v += T)  // punctured variables: prior (0.5, 0.5) (binary5gldpccodec.cc:126-130)
      chan[v] = v < p.t.punct ? 1.0f : load_channel_ratio(in, v - p.t.punct, p.in_is_lr);
    for (int i = tid; i < n_words; i += T) msg[i] = 0x3f800000u;
    __syncthreads();

    int ret = p.iters + (p.iters < p.max_iter), nfail = 0;
    bool latched = false;
    int last_t = 0;
    float soft = 0.0f, soft_out = 0.0f;
    for (int t = 0; t < p.iters; t++) {
      last_t = t;
      uint32_t *dcur = dec + (t & 1) * p.words_n;
      for (int i = warp; i < p.t.vn_items_n; i += W) {
        // item = first group | degree << 16 (0xFF = mixed degrees or ragged: per-lane dispatch) | number of groups << 24
        const uint32_t item = __ldg(p.t.vn_items + i);
        if (item == 0xFFFFFFFFu) continue;  // padding of a shorter list (warp-uniform)
        const int g0 = item & 0xFFFF, cnt = (int)(item >> 24), dvm = p.t.dv_max;
        // edges in the order the layout optimiser coloured them (= gather instruction index inside the warp)
        const uint16_t *ad = p.t.vn_addr_g + (size_t)g0 * (32 * dvm) + lane;
        int v = g0 * 32 + lane;
        switch ((item >> 16) & 0xFF) {
          KML_VN_CASE(1) KML_VN_CASE(2) KML_VN_CASE(3) KML_VN_CASE(4) KML_VN_CASE(5) KML_VN_CASE(6) KML_VN_CASE(7) KML_VN_CASE(8)
          KML_VN_CASE(9) KML_VN_CASE(10) KML_VN_CASE(11) KML_VN_CASE(12) KML_VN_CASE(13) KML_VN_CASE(14) KML_VN_CASE(15)
          KML_VN_CASE(16)
          default:
            for (int r = 0; r < cnt; r++, v += 32, ad += 32 * dvm) {
              uint32_t bit = 0;
              if (v < n) bit = vn_lane_dispatch<DV>(msg, ad, chan[v], (int)__ldg(p.t.vn_deg + v));
              const uint32_t word = __ballot_sync(0xffffffffu, bit);
              if (lane == 0) dcur[g0 + r] = word;
            }
            break;
        }
      }
      __syncthreads();
      int fail = 0;
      const float soft_before = soft;
      soft = 0.0f;
      for (int i = warp; i < p.t.cn_items_n; i += W) {
        const uint32_t item = __ldg(p.t.cn_items + i);
        if (item == 0xFFFFFFFFu) continue;
        const int cnt = (int)(item >> 24);
        int slot = (item & 0xFFFF) * 32 + lane;
        switch ((item >> 16) & 0xFF) {
          KML_CN_CASE(1) KML_CN_CASE(2) KML_CN_CASE(3) KML_CN_CASE(4) KML_CN_CASE(5) KML_CN_CASE(6) KML_CN_CASE(7) KML_CN_CASE(8)
          KML_CN_CASE(9) KML_CN_CASE(10) KML_CN_CASE(11) KML_CN_CASE(12) KML_CN_CASE(13) KML_CN_CASE(14) KML_CN_CASE(15)
          KML_CN_CASE(16)
          default:
            for (int r = 0; r < cnt; r++, slot += 32) {
              float s_all = 0.0f;
              const int deg = (int)__ldg(p.t.cn_deg + slot);
              if (deg == 0) continue;  // padding slot
              const uint32_t x = cn_lane_dispatch<DC, SOFT>(msg + slot * RS, &s_all, deg);
              fail += (int)(x & 1u);
              if (SOFT) soft += soft_log(x >> 31, s_all);
            }
            break;
        }
      }
      nfail = fail;  // this thread's unsatisfied checks of the decisions just made
      const int any_fail = __syncthreads_or(fail);
      if (!any_fail && !latched) {
        latched = true;
        soft_out = soft_before;
        ret = t + (t < p.max_iter);
        for (int w = tid; w < p.words_n; w += T) p.out_bits[(size_t)f * p.words_n + w] = dcur[w];
        if (p.early_exit) break;
      }
    }
    if (!latched) {
      soft_out = soft;
      const uint32_t *dl = dec + (last_t & 1) * p.words_n;
      for (int w = tid; w < p.words_n; w += T) p.out_bits[(size_t)f * p.words_n + w] = dl[w];
    }
    if (tid == 0) p.out_ret[f] = ret;
    if (p.out_synd) {  // ParityCheck(cc_hat) of the final decisions (kmcodec.cc:157-160), summed over the CTA bit by bit
      int total = 0;
#pragma unroll
      for (int b = 0; b < 8; b++) total += __syncthreads_count((nfail >> b) & 1) << b;
      if (tid == 0) p.out_synd[f] = latched ? 0.0f : (float)total;
    }
    if (SOFT) {
      double sd = (double)soft_out;  // per-thread partial sums are short; everything across threads is fp64
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
      if (lane == 0) atomicAdd(p.out_soft + f, sd);
    }
  }
}
#undef KML_VN_CASE
#undef KML_CN_CASE
#undef KML_VN_LANE_CASE
#undef KML_CN_LANE_CASE

// ---------------------------------------------------------------------------------------------------------------
// Quasi-cyclic codes with a compile-time plan: the block structure (which block columns / block rows a thread serves,
// and their degrees) is a template argument, so the node updates are fully unrolled, the edge addresses and channel
// ratios live in registers and nothing is dispatched at run time — the regular kernel's organisation for an
// irregular graph.  The CTA's 4 * Z threads form 4 "quarters" of Z threads; thread z of quarter q serves node z of
// every block in the quarter's list.  The lists balance the edges per thread (19/19/19/20 for BG2).
// The host (kml_api.cu) checks that the loaded graph has exactly this block structure before choosing the kernel;
// any other graph runs bp_generic_kernel.
// ---------------------------------------------------------------------------------------------------------------
struct QcPlanBg2R12 {  // 5G NR BG2 truncated to 12 block rows x 22 block columns (rate 1/2 after puncturing 2 Z), Z = 96:
                       // config/5GLDPCBG2a3_R12_K960.txt in the column order the reference's elimination leaves
  static constexpr int Z = 96, MAXV = 6, MAXC = 3, DVM = 9, RS = 11;
  static constexpr int vblk[4][6] = {{0, 5, 3, 14, 15, -1}, {1, 9, 4, 16, 17, -1}, {7, 6, 8, 2, 18, 19}, {11, 10, 13, 12, 20, 21}};
  static constexpr int vdeg[4][6] = {{9, 5, 3, 1, 1, 0}, {9, 5, 3, 1, 1, 0}, {7, 4, 4, 2, 1, 1}, {7, 4, 4, 3, 1, 1}};
  static constexpr int cblk[4][3] = {{1, 9, 4}, {3, 10, 8}, {0, 5, 11}, {2, 6, 7}};
  static constexpr int cdeg[4][3] = {{10, 5, 4}, {10, 5, 4}, {8, 6, 5}, {8, 6, 6}};
};

template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
  static_for_impl(static_cast<F &&>(f), std::make_integer_sequence<int, N>{});
}

template <class P, int Q>
__device__ __forceinline__ constexpr int qc_voff(int i) {  // edges of the quarter's variables before variable i
  int o = 0;
  for (int j = 0; j < i; j++) o += P::vdeg[Q][j];
  return o;
}

// One quarter's whole decoding loop.  All quarters execute the same sequence of barriers (frame queue, iteration count
// and exit decisions are CTA-uniform), so the BAR instructions of the four instances pair up.
// ALG = 0: the reference's sum-product; ALG = 1: normalised min-sum (bp_minsum.cu; messages are LLRs, not ratios).
template <class P, int Q, int ALG>
__device__ __forceinline__ void qc_quarter(const DecParams &p, uint32_t *msg, volatile int *s_frame, int z) {
  constexpr int NE = qc_voff<P, Q>(P::MAXV);
  const int tid = threadIdx.x, lane = tid & 31;
  uint32_t va[NE];
  static_for<P::MAXV>([&](auto ic) {
    constexpr int i = decltype(ic)::value, D = P::vdeg[Q][i], O = qc_voff<P, Q>(i);
    if constexpr (D > 0) {
      const int v = P::vblk[Q][i] * P::Z + z;
#pragma unroll
      for (int k = 0; k < D; k++) va[O + k] = __ldg(p.t.vn_addr + (size_t)v * P::DVM + k);
    }
  });
  const int n_words = p.t.m_pad * P::RS;
  while (true) {
    if (tid == 0) *s_frame = next_frame(p);
    __syncthreads();
    const int f = *s_frame;
    if (f < 0) break;
    const float *in = p.in + (size_t)(p.sel ? f * p.n_cand + __ldg(p.sel + f) : f) * p.t.n_tx;
    float ch[P::MAXV];
    static_for<P::MAXV>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      ch[i] = ALG == 0 ? 1.0f : 0.0f;
      if constexpr (P::vdeg[Q][i] > 0) {
        const int v = P::vblk[Q][i] * P::Z + z;  // punctured variables: prior (0.5, 0.5) (binary5gldpccodec.cc:126-130)
        if (v >= p.t.punct)
          ch[i] = ALG == 0 ? load_channel_ratio(in, v - p.t.punct, p.in_is_lr) : msn::load_channel_llr(in, v - p.t.punct, p.in_is_lr);
      }
    });
    for (int i = tid; i < n_words; i += blockDim.x) msg[i] = ALG == 0 ? 0x3f800000u : 0u;  // c2v = (0.5, 0.5)
    __syncthreads();

    uint32_t bits = 0, latched_bits = 0;
    int ret = p.iters + (p.iters < p.max_iter), nfail = 0;
    bool latched = false;
    for (int t = 0; t < p.iters; t++) {
      bits = 0;
      static_for<P::MAXV>([&](auto ic) {
        constexpr int i = decltype(ic)::value, D = P::vdeg[Q][i], O = qc_voff<P, Q>(i);
        if constexpr (D > 0) bits |= (ALG == 0 ? vn_core<D>(msg, va + O, ch[i]) : msn::ms_vn<D>(msg, va + O, ch[i])) << i;
      });
      __syncthreads();
      int fail = 0;  // this thread's unsatisfied checks (of the decisions just made)
      static_for<P::MAXC>([&](auto jc) {
        constexpr int j = decltype(jc)::value, D = P::cdeg[Q][j];
        float unused;
        uint32_t *row = msg + (P::cblk[Q][j] * P::Z + z) * P::RS;
        fail += (int)((ALG == 0 ? cn_node<D, false>(row, &unused) : msn::ms_cn<D>(row, 1, 0, p.alpha, p.beta)) & 1u);
      });
      nfail = fail;
      const int any_fail = __syncthreads_or(fail);
      if (!any_fail && !latched) {
        latched = true;
        latched_bits = bits;
        ret = t + (t < p.max_iter);
        if (p.early_exit) break;  // the reference leaves BEFORE the check phase; its c2v are never read again
      }
    }
    if (!latched) latched_bits = bits;
    static_for<P::MAXV>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      if constexpr (P::vdeg[Q][i] > 0) {
        const uint32_t word = __ballot_sync(0xffffffffu, (latched_bits >> i) & 1u);
        if (lane == 0) p.out_bits[(size_t)f * p.words_n + ((P::vblk[Q][i] * P::Z + z) >> 5)] = word;
      }
    });
    if (tid == 0) p.out_ret[f] = ret;
    if (p.out_synd) {  // ParityCheck(cc_hat) of the final decisions (kmcodec.cc:157-160): 0 when latched, else the count
      static_assert(P::MAXC <= 3, "two count bits");  // of the last check phase, summed over the CTA bit by bit
      const int total = __syncthreads_count(nfail & 1) + 2 * __syncthreads_count(nfail & 2);
      if (tid == 0) p.out_synd[f] = latched ? 0.0f : (float)total;
    }
  }
}

// ALG = 2: the min-sum of ALG = 1 with fp16 messages and TWO frames per 32-bit word (as ms2_regular_kernel in bp_minsum.cu):
// a work item is a pair of queue entries, each frame latches its own decisions and return value.
template <class P, int Q>
__device__ __forceinline__ void qc_quarter_ms2(const DecParams &p, uint32_t *msg, volatile int *s_pair, int z) {
  constexpr int NE = qc_voff<P, Q>(P::MAXV);
  const int tid = threadIdx.x, lane = tid & 31;
  uint32_t va[NE];
  static_for<P::MAXV>([&](auto ic) {
    constexpr int i = decltype(ic)::value, D = P::vdeg[Q][i], O = qc_voff<P, Q>(i);
    if constexpr (D > 0) {
      const int v = P::vblk[Q][i] * P::Z + z;
#pragma unroll
      for (int k = 0; k < D; k++) va[O + k] = __ldg(p.t.vn_addr + (size_t)v * P::DVM + k);
    }
  });
  const int n_words = p.t.m_pad * P::RS;
  const __half2 alpha2 = __float2half2_rn(p.alpha), nbeta2 = __float2half2_rn(-p.beta);
  const bool has_offset = p.beta != 0.0f;
  while (true) {
    if (tid == 0) {  // two consecutive entries of the frame queue (the last item may be single: fb = fa)
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
    __half2 ch[P::MAXV];
    static_for<P::MAXV>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      ch[i] = __float2half2_rn(0.0f);
      if constexpr (P::vdeg[Q][i] > 0) {
        const int v = P::vblk[Q][i] * P::Z + z;  // punctured variables: prior (0.5, 0.5) (binary5gldpccodec.cc:126-130)
        if (v >= p.t.punct)
          ch[i] = __floats2half2_rn(msn::load_channel_llr(ina, v - p.t.punct, p.in_is_lr),
                                    msn::load_channel_llr(inb, v - p.t.punct, p.in_is_lr));
      }
    });
    for (int i = tid; i < n_words; i += blockDim.x) msg[i] = 0u;
    __syncthreads();

    uint32_t bits_a = 0, bits_b = 0, lat_a = 0, lat_b = 0;
    const int ret_full = p.iters + (p.iters < p.max_iter);
    int ret_a = ret_full, ret_b = ret_full, nfail_a = 0, nfail_b = 0;
    bool done_a = false, done_b = false;
    for (int t = 0; t < p.iters; t++) {
      bits_a = bits_b = 0;
      static_for<P::MAXV>([&](auto ic) {
        constexpr int i = decltype(ic)::value, D = P::vdeg[Q][i], O = qc_voff<P, Q>(i);
        if constexpr (D > 0) {
          const uint32_t pb = msn::ms2_vn<D>(msg, va + O, ch[i]);
          bits_a |= (pb & 1u) << i;
          bits_b |= (pb >> 16) << i;
        }
      });
      __syncthreads();
      int fa_cnt = 0, fb_cnt = 0;  // this thread's unsatisfied checks of the decisions just made, per frame
      static_for<P::MAXC>([&](auto jc) {
        constexpr int j = decltype(jc)::value, D = P::cdeg[Q][j];
        const uint32_t x = msn::ms2_cn<D>(msg + (P::cblk[Q][j] * P::Z + z) * P::RS, alpha2, nbeta2, has_offset);
        fa_cnt += (int)(x & 1u);
        fb_cnt += (int)((x >> 16) & 1u);
      });
      const int any_a = __syncthreads_or(fa_cnt), any_b = __syncthreads_or(fb_cnt);
      if (!done_a) nfail_a = fa_cnt;  // (a latched frame reports syndrome weight 0 below)
      if (!done_b) nfail_b = fb_cnt;
      if (!any_a && !done_a) { done_a = true; lat_a = bits_a; ret_a = t + (t < p.max_iter); }
      if (!any_b && !done_b) { done_b = true; lat_b = bits_b; ret_b = t + (t < p.max_iter); }
      if (done_a && done_b && p.early_exit) break;
    }
    if (!done_a) lat_a = bits_a;
    if (!done_b) lat_b = bits_b;
    static_for<P::MAXV>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      if constexpr (P::vdeg[Q][i] > 0) {
        const uint32_t wa = __ballot_sync(0xffffffffu, (lat_a >> i) & 1u), wb = __ballot_sync(0xffffffffu, (lat_b >> i) & 1u);
        if (lane == 0) {
          p.out_bits[(size_t)fa * p.words_n + ((P::vblk[Q][i] * P::Z + z) >> 5)] = wa;
          if (fb != fa) p.out_bits[(size_t)fb * p.words_n + ((P::vblk[Q][i] * P::Z + z) >> 5)] = wb;
        }
      }
    });
    if (tid == 0) {
      p.out_ret[fa] = ret_a;
      if (fb != fa) p.out_ret[fb] = ret_b;
    }
    if (p.out_synd) {  // ParityCheck(cc_hat) of the final decisions, per frame (see qc_quarter)
      static_assert(P::MAXC <= 3, "two count bits");
      const int ta = __syncthreads_count(nfail_a & 1) + 2 * __syncthreads_count(nfail_a & 2);
      const int tb = __syncthreads_count(nfail_b & 1) + 2 * __syncthreads_count(nfail_b & 2);
      if (tid == 0) {
        p.out_synd[fa] = done_a ? 0.0f : (float)ta;
        if (fb != fa) p.out_synd[fb] = done_b ? 0.0f : (float)tb;
      }
    }
  }
}

template <class P, int MINB, int ALG>
__global__ void __launch_bounds__(4 * P::Z, MINB) bp_qc_kernel(const DecParams p) {
  static_assert(P::Z % 32 == 0, "a warp must not straddle two quarters");
  extern __shared__ uint32_t msg[];  // [m_pad][RS]
  __shared__ int s_frame[2];         // the frame (ALG 0 / 1) or the pair of frames (ALG 2) being decoded
  const int q = threadIdx.x / P::Z, z = threadIdx.x % P::Z;
  if constexpr (ALG == 2) {
    switch (q) {  // warp-uniform
      case 0: qc_quarter_ms2<P, 0>(p, msg, s_frame, z); break;
      case 1: qc_quarter_ms2<P, 1>(p, msg, s_frame, z); break;
      case 2: qc_quarter_ms2<P, 2>(p, msg, s_frame, z); break;
      default: qc_quarter_ms2<P, 3>(p, msg, s_frame, z); break;
    }
  } else {
    switch (q) {  // warp-uniform
      case 0: qc_quarter<P, 0, ALG>(p, msg, s_frame, z); break;
      case 1: qc_quarter<P, 1, ALG>(p, msg, s_frame, z); break;
      case 2: qc_quarter<P, 2, ALG>(p, msg, s_frame, z); break;
      default: qc_quarter<P, 3, ALG>(p, msg, s_frame, z); break;
    }
  }
}

// Resolved once per context (dec_prepare), never on the launch path.  `threads` = the CTA size the tables were built
// for.  The shipped library picks between kernels that all reproduce the reference (run-time knobs KML_DEC_PLANAR,
// KML_DEC_NO_QC, KML_DEC_T8064 choose the fallback layouts / tilings and announce themselves on stderr, kml_internal.h);
// the A/B and timing-ablation variants exist only in a -DKML_TUNING build.
dec_kernel_t kernel_of(DecKernelKind k, int alg, int rowmajor, bool soft, int qc_plan, int threads) {
  if (alg == 3) return layered_kernel();
  if (alg == 2 && qc_plan == 1) return bp_qc_kernel<QcPlanBg2R12, 3, 2>;  // fp16 x 2 frames per word
  if (alg != 0 && qc_plan == 1) return bp_qc_kernel<QcPlanBg2R12, 3, 1>;
  if (alg != 0) return minsum_kernel_of(k, alg);
  if (qc_plan == 1 && !soft) {
#ifdef KML_TUNING
    const char *e = tuning_knob("KML_DEC_QC_MINB");
    if (e && atoi(e) == 2) return bp_qc_kernel<QcPlanBg2R12, 2, 0>;
#endif
    return bp_qc_kernel<QcPlanBg2R12, 3, 0>;
  }
  // Soft-syndrome output of the (3,6)-regular codes: the scalar (PACK = false) planar kernel.  Its variable nodes form
  // the three extrinsic ratios directly (each a product of the channel ratio and two clipped messages, within 1e+-36),
  // so the small probabilities never leave the fp32 range — the packed kernels go through post = ch x0 x1 x2 (up to
  // 1e+-48), whose overflow is harmless for the messages (everything below 1e-12 is clipped alike) but would turn
  // ln(syndrom_soft) of a confidently violated check into -inf where the reference sums a finite ~ -83.
  if (soft && k == DEC_REG_6_3) return bp_regular_kernel<6, 3, 384, 3, false, 2, false, true>;
  if (soft && k == DEC_REG_12_6) return bp_regular_kernel<12, 6, 672, 1, false, 2, false, true>;
  switch (k) {
    case DEC_REG_6_3: {
#ifdef KML_TUNING
      const char *e = tuning_knob("KML_DEC_MINB");   // CTAs per SM the register allocation targets
      const int b = e ? atoi(e) : 3;
      const char *re = tuning_knob("KML_DEC_RATIO");  // how check outputs with hard bit 1 are inverted
      const int r = re ? atoi(re) : 1;
      if (!rowmajor) {
        const char *pe = tuning_knob("KML_DEC_NOPACK");  // scalar fp32 instead of FMUL2/FADD2/FFMA2
        if (pe && atoi(pe)) return bp_regular_kernel<6, 3, 384, 3, false>;
        if (r == 0) return bp_regular_kernel<6, 3, 384, 3, true, 0>;
        if (b == 2) return bp_regular_kernel<6, 3, 384, 2>;
        if (b == 4) return bp_regular_kernel<6, 3, 384, 4>;
        if (r == 2) return bp_regular_kernel<6, 3, 384, 3>;
      } else {
        if (const char *de = tuning_knob("KML_DEC_DIAG")) {  // timing ablations, WRONG results (see the kernel's header)
          const int dg = atoi(de);
          if (dg == 1) return bp_regular_kernel<6, 3, 384, 3, true, 1, true, false, 2304, 1>;
          if (dg == 2) return bp_regular_kernel<6, 3, 384, 3, true, 1, true, false, 2304, 2>;
          if (dg == 3) return bp_regular_kernel<6, 3, 384, 3, true, 1, true, false, 2304, 3>;
        }
        if (r == 0) return bp_regular_kernel<6, 3, 384, 3, true, 0, true>;
        if (r == 3) return bp_regular_kernel<6, 3, 384, 3, true, 3, true>;
        if (b == 2) return bp_regular_kernel<6, 3, 384, 2, true, 2, true>;
        if (b == 4) return bp_regular_kernel<6, 3, 384, 4, true, 2, true>;
        if (r == 2) return bp_regular_kernel<6, 3, 384, 3, true, 2, true>;
      }
#endif
      return rowmajor ? bp_regular_kernel<6, 3, 384, 3, true, 1, true> : bp_regular_kernel<6, 3, 384, 3, true, 1>;
    }
    case DEC_REG_12_6: {
#ifdef KML_TUNING
      const char *re = tuning_knob("KML_DEC_RATIO");
      if (re && atoi(re) == 2)
        return rowmajor && threads == 1024 ? bp_regular_kernel<8, 4, 1024, 1, true, 2, true, false, 8064>
               : rowmajor                  ? bp_regular_kernel<12, 6, 672, 1, true, 2, true>
                                           : bp_regular_kernel<12, 6, 672, 1>;
#endif
      if (rowmajor && threads == 1024) return bp_regular_kernel<8, 4, 1024, 1, true, 1, true, false, 8064>;
      return rowmajor ? bp_regular_kernel<12, 6, 672, 1, true, 1, true> : bp_regular_kernel<12, 6, 672, 1, true, 1>;
    }
    case DEC_GEN_4_8: return soft ? bp_generic_kernel<4, 8, true> : bp_generic_kernel<4, 8, false>;
    case DEC_GEN_9_10: return soft ? bp_generic_kernel<9, 10, true> : bp_generic_kernel<9, 10, false>;
    case DEC_GEN_16_32: return soft ? bp_generic_kernel<16, 16, true> : bp_generic_kernel<16, 16, false>;
  }
  return nullptr;
}

}  // namespace

int dec_generic_max_threads() { return kGenericThreads; }

// Does the graph (degrees per variable / per row slot of the row-major layout) have the block structure of a compiled
// quasi-cyclic plan?  Returns the plan id (1 = QcPlanBg2R12) or 0.
int dec_match_qc_plan(int n, int m_pad, const uint8_t *vdeg, const uint8_t *cndeg, int dv_max, int row_stride) {
  using P = QcPlanBg2R12;
  const char *e = knob("KML_DEC_NO_QC");  // run the run-time-graph kernel instead (shipped fallback; announced on stderr)
  if (e && atoi(e)) return 0;
  if (n != 22 * P::Z || m_pad != 12 * P::Z || dv_max != P::DVM || row_stride != P::RS) return 0;
  int seen_v[22] = {0}, seen_c[12] = {0};
  for (int q = 0; q < 4; q++) {
    for (int i = 0; i < P::MAXV; i++) {
      if (P::vdeg[q][i] == 0) continue;
      const int b = P::vblk[q][i];
      seen_v[b]++;
      for (int zz = 0; zz < P::Z; zz++)
        if (vdeg[b * P::Z + zz] != P::vdeg[q][i]) return 0;
    }
    for (int j = 0; j < P::MAXC; j++) {
      const int b = P::cblk[q][j];
      seen_c[b]++;
      for (int zz = 0; zz < P::Z; zz++)
        if (cndeg[b * P::Z + zz] != P::cdeg[q][j]) return 0;
    }
  }
  for (int b = 0; b < 22; b++) if (seen_v[b] != 1) return 0;
  for (int b = 0; b < 12; b++) if (seen_c[b] != 1) return 0;
  return 1;
}
int dec_generic_row_stride(DecKernelKind k) { return k == DEC_GEN_4_8 ? 9 : k == DEC_GEN_9_10 ? 11 : 17; }

int dec_regular_threads(DecKernelKind k) {
  if (k == DEC_REG_12_6) {
    // 1024 threads x (8 variables, 4 checks; the last round is empty for the top 128 threads): 32 warps on the SM's
    // single CTA instead of 21; KML_DEC_T8064=672 is the A/B knob for the exact 672 x (12, 6) tiling, which the planar
    // layout also uses
    const char *e = knob("KML_DEC_T8064");
    return (!(e && atoi(e) == 672) && dec_wants_rowmajor(k, 0)) ? 1024 : 672;
  }
  return 384;  // (576 threads x (4 variables, 2 checks), 2 CTAs per SM measured 7 % slower)
}

bool dec_wants_rowmajor(DecKernelKind k, int alg) {
  if (alg != 0) return false;
  if (k != DEC_REG_6_3 && k != DEC_REG_12_6) return true;  // generic sum-product kernel: always its own row-major tables
  const char *e = knob("KML_DEC_PLANAR");  // the planar layout for the regular sum-product kernels too (shipped fallback)
  return !(e && atoi(e));
}

bool dec_has_synd_output(const DecLaunch &l, bool soft) {
  if (l.qc_plan && !soft) return true;                                  // bp_qc_kernel, both algorithms
  return l.alg == 0 && l.kind != DEC_REG_6_3 && l.kind != DEC_REG_12_6;  // bp_generic_kernel
}

cudaError_t dec_prepare(DecLaunch &l) {
  // l.soft: the launch record of the kernel that also produces DecParams::out_soft (sum-product only; a plan kernel hands
  // over to the run-time-graph kernel)
  dec_kernel_t k = kernel_of(l.kind, l.alg, l.rowmajor, l.soft != 0, l.soft ? 0 : l.qc_plan, l.threads);
  if (!k) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, l.smem_bytes);
  if (e != cudaSuccess) return e;
  l.fn = k;
  int n = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, l.threads, l.smem_bytes);
  if (e != cudaSuccess) return e;
  if (n < 1) return cudaErrorLaunchOutOfResources;
  l.ctas_per_sm = n;
  return cudaSuccess;
}

cudaError_t dec_launch(const DecLaunch &l, const DecParams &p, int num_sms, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned int), s);
  if (e != cudaSuccess) return e;
  int grid = num_sms * l.ctas_per_sm;
  // work items in the frame queue per CTA: a pair of frames (fp16 x 2), or the layered kernel's concurrent frame groups
  const int per = l.alg == 3 ? layered_frames_per_cta() : ((dec_two_frames_per_cta(l.kind, l.alg) || (l.alg == 2 && l.qc_plan)) ? 2 : 1);
  const int units = (p.B + per - 1) / per;
  if (grid > units) grid = units;
  if (grid < 1) return cudaSuccess;
  l.fn<<<grid, l.threads, l.smem_bytes, s>>>(p);
  return cudaGetLastError();
}

}  // namespace kml
