// Everything of the link path that is not the BP decoder: Philox source, GF(2) encoder, mapper + block-fading AWGN
// channel, k-means blind channel estimate, soft demapper + candidate resolver, error counting (sm_100a).
// Reference functions are cited at each kernel (paths relative to the reference's kmldpc/ directory).
#include <algorithm>
#include <cstdio>

#include "kml_internal.h"
#include "kml_kernels.cuh"

namespace kml {
namespace {

// ------------------------------------------------------------------------------------------------ Philox4x32-10
// Counter-based RNG replacing the shared, unlocked LCG of the reference (lib/lab/src/randnum.cc:36-45): any sample is
// a pure function of (seed, stream, frame index, sample index), so results are independent of batching and GPU count.
struct Philox4 {
  uint32_t x, y, z, w;
};
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                  uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return {c0, c1, c2, c3};
}
__device__ __forceinline__ Philox4 philox_at(uint64_t seed, uint32_t stream, uint64_t frame, uint32_t idx) {
  return philox4x32_10(idx, (uint32_t)frame, (uint32_t)(frame >> 32), stream, (uint32_t)seed, (uint32_t)(seed >> 32));
}
// Box–Muller: two independent N(0,1) from two 32-bit words
__device__ __forceinline__ float2 gauss_pair(uint32_t a, uint32_t b) {
  const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;  // (0, 1]
  const float u2 = (float)b * 2.3283064365386963e-10f;           // [0, 1)
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

// 32 bits starting at bit position q of the concatenation [A (len_a bits) | B]
__device__ __forceinline__ uint32_t word_at(const uint32_t *w, int q) {
  const int i = q >> 5, sh = q & 31;
  const uint32_t lo = w[i];
  if (sh == 0) return lo;
  return __funnelshift_r(lo, w[i + 1], sh);
}

// ------------------------------------------------------------------------------------------------ source bits
// SourceSink::GetBitStr (lib/lab/src/sourcesink.cc:5-9): K fair bits per frame.
__global__ void gen_bits_kernel(GenParams g, uint32_t *u_packed) {
  const int groups = (g.k_words + 3) / 4;
  const long long total = (long long)g.B * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / groups), gr = (int)(i % groups);
    uint32_t v[4] = {0, 0, 0, 0};
    if (g.encoder_active) {  // [ldpc] active = false → all-zero word (binaryldpccodec.cc:156-161)
      const Philox4 r = philox_at(g.seed, STREAM_BITS, g.frame0 + f, gr);
      v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int w = gr * 4 + j;
      if (w < g.k_words) {
        uint32_t x = v[j];
        if (w == g.k_words - 1 && (g.k & 31)) x &= (1u << (g.k & 31)) - 1u;
        u_packed[(size_t)f * g.k_words + w] = x;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ encoder
// BinaryLDPCCodec::Encoder (binaryldpccodec.cc:144-162): cc = [p | u], p_t = XOR_j u_j & enc_h[t][chk + j]
// Binary5GLDPCCodec::Encoder (binary5gldpccodec.cc:86-109): cc_np = [u | p], transmitted = cc_np[2Z:]
// One CTA encodes ENC_FT frames: thread t owns parity row t (rows strided by the block), reads its row of the
// transposed matrix once per word and applies it to all ENC_FT frames; the information words are also kept word-major
// ([word][frame]) so that one warp-broadcast LDS.128 delivers the same word of four frames (per matrix word:
// 1 LDG + 4 LDS.128 + 16 LOP3 for 16 frames).
constexpr int ENC_FT = 16;
constexpr int ENC_THREADS = 256;
// GEN: the information bits are drawn here (SourceSink::GetBitStr, the Philox stream of gen_bits_kernel) and written to
// u_packed as well, instead of being read from it — one launch and one HBM round trip less in the fused path.
template <bool GEN>
__global__ void __launch_bounds__(ENC_THREADS) encode_kernel(GenParams g, uint32_t *u_packed, uint32_t *c_packed) {
  extern __shared__ __align__(16) uint32_t sm[];
  const int chk_words = (g.n_chk + 31) / 32;
  uint32_t *su = sm;                                   // [ENC_FT][k_words + 1]
  uint32_t *sp = su + ENC_FT * (g.k_words + 1);        // [ENC_FT][chk_words + 1]
  uint4 *sut = reinterpret_cast<uint4 *>(sm + (((ENC_FT * (g.k_words + 1) + ENC_FT * (chk_words + 1)) + 3) & ~3));  // [k_words][ENC_FT]
  const int f0 = blockIdx.x * ENC_FT;
  const int nf = min(ENC_FT, g.B - f0);
  if (GEN) {
    const int groups = (g.k_words + 3) / 4;
    for (int i = threadIdx.x; i < ENC_FT * groups; i += blockDim.x) {
      const int f = i / groups, gr = i % groups;
      uint32_t v[4] = {0, 0, 0, 0};
      if (f < nf && g.encoder_active) {  // [ldpc] active = false → all-zero word (binaryldpccodec.cc:156-161)
        const Philox4 r = philox_at(g.seed, STREAM_BITS, g.frame0 + f0 + f, gr);
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int w = gr * 4 + j;
        if (w < g.k_words) {
          uint32_t x = v[j];
          if (w == g.k_words - 1 && (g.k & 31)) x &= (1u << (g.k & 31)) - 1u;
          su[f * (g.k_words + 1) + w] = x;
          reinterpret_cast<uint32_t *>(sut)[w * ENC_FT + f] = x;
          if (f < nf) u_packed[(size_t)(f0 + f) * g.k_words + w] = x;
        }
      }
      if (gr == 0) su[f * (g.k_words + 1) + g.k_words] = 0u;
    }
  } else {
    for (int i = threadIdx.x; i < ENC_FT * (g.k_words + 1); i += blockDim.x) {
      const int f = i / (g.k_words + 1), w = i % (g.k_words + 1);
      const uint32_t v = (f < nf && w < g.k_words) ? u_packed[(size_t)(f0 + f) * g.k_words + w] : 0u;
      su[i] = v;
      if (w < g.k_words) reinterpret_cast<uint32_t *>(sut)[w * ENC_FT + f] = v;
    }
  }
  for (int i = threadIdx.x; i < ENC_FT * (chk_words + 1); i += blockDim.x) sp[i] = 0u;
  __syncthreads();
  const int rows_round = (g.n_chk + 31) & ~31;
  for (int t = threadIdx.x; t < rows_round; t += blockDim.x) {
    uint32_t acc[ENC_FT];
#pragma unroll
    for (int f = 0; f < ENC_FT; f++) acc[f] = 0;
    if (t < g.n_chk && g.encoder_active) {
      for (int w = 0; w < g.k_words; w++) {
        const uint32_t e = __ldg(g.enc_t + (size_t)w * g.n_chk + t);
#pragma unroll
        for (int f4 = 0; f4 < ENC_FT / 4; f4++) {
          const uint4 u4 = sut[w * (ENC_FT / 4) + f4];
          acc[4 * f4 + 0] ^= e & u4.x;
          acc[4 * f4 + 1] ^= e & u4.y;
          acc[4 * f4 + 2] ^= e & u4.z;
          acc[4 * f4 + 3] ^= e & u4.w;
        }
      }
    }
#pragma unroll
    for (int f = 0; f < ENC_FT; f++) {
      const uint32_t word = __ballot_sync(0xffffffffu, __popc(acc[f]) & 1);
      if ((threadIdx.x & 31) == 0) sp[f * (chk_words + 1) + (t >> 5)] = word;
    }
  }
  __syncthreads();
  // assemble the transmitted word: S = [p | u] (PEG) or [u | p] (5G), c = S[punct : punct + n_tx]
  for (int i = threadIdx.x; i < nf * g.tx_words; i += blockDim.x) {
    const int f = i / g.tx_words, w = i % g.tx_words;
    const uint32_t *a = g.is_5g ? su + f * (g.k_words + 1) : sp + f * (chk_words + 1);
    const uint32_t *b = g.is_5g ? sp + f * (chk_words + 1) : su + f * (g.k_words + 1);
    const int len_a = g.is_5g ? g.k : g.n_chk;
    const int q = g.punct + 32 * w;  // first S position of this output word
    uint32_t out;
    if (q + 32 <= len_a) out = word_at(a, q);
    else if (q >= len_a) out = word_at(b, q - len_a);
    else {
      const int na = len_a - q;  // bits still coming from A
      out = (word_at(a, q) & ((1u << na) - 1u)) | (b[0] << na);
    }
    const int valid = g.n_tx - 32 * w;
    if (valid < 32) out &= (1u << valid) - 1u;
    if (!g.encoder_active) out = 0u;  // [ldpc] active = false: all-zero codeword (binaryldpccodec.cc:156-161)
    c_packed[(size_t)(f0 + f) * g.tx_words + w] = out;
  }
}

// ------------------------------------------------------------------------------------------------ mapper + channel
// Modem::Mapping (lib/lab/src/modem.cc:12-20), the fading draw (src/simulator.cc:120-128) and
// ModemLinearSystem::PartitionHAWGNSystem (lib/lab/src/modemlinearsystem.cc:37-48): y = h x + (sigma/sqrt2) (a + jb)
__device__ __forceinline__ float2 map_symbol(const GenParams &g, const uint32_t *c, int sidx) {
  int idx = 0;
  const int b0 = sidx * g.bits_per_symbol;
  for (int j = 0; j < g.bits_per_symbol; j++) {  // MSB first
    const int t = b0 + j;
    idx = (idx << 1) | (int)((c[t >> 5] >> (t & 31)) & 1u);
  }
  return __ldg(g.points + idx);
}

// replay of given fades / noise (parity tests against the reference's own channel outputs): one thread per symbol
__global__ void channel_replay_kernel(GenParams g, const uint32_t *c_packed, const float2 *h_in, const float2 *noise, float2 *y) {
  const long long total = (long long)g.B * g.n_sym;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / g.n_sym), sidx = (int)(i % g.n_sym);
    const float2 x = map_symbol(g, c_packed + (size_t)f * g.tx_words, sidx), h = h_in[f], nz = noise[i];
    y[i] = make_float2(h.x * x.x - h.y * x.y + g.sigma_over_sqrt2 * nz.x, h.x * x.y + h.y * x.x + g.sigma_over_sqrt2 * nz.y);
  }
}

// Philox channel.  One CTA per frame (grid-stride): the fade is drawn once per frame, one Philox4x32-10 call feeds the
// noise of TWO symbols (counter = (symbol pair, frame, stream): independent of batch size, launch shape and GPU count).
// Box-Muller with the fast intrinsics (lg2 / sin / cos / sqrt on the SFU: ~1e-6 relative, far below the noise itself).
__device__ __forceinline__ float2 gauss_pair_fast(uint32_t a, uint32_t b) {
  const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;  // (0, 1]
  const float ang = (float)b * (6.283185307179586f * 2.3283064365386963e-10f);  // [0, 2 pi)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(u1)));
  float s, c;
  __sincosf(ang, &s, &c);
  return make_float2(r * c, r * s);
}

constexpr int CH_THREADS = 192;
__global__ void __launch_bounds__(CH_THREADS) channel_kernel(GenParams g, const uint32_t *c_packed, float2 *h_out, float2 *y) {
  const int pairs = (g.n_sym + 1) >> 1;
  for (int f = blockIdx.x; f < g.B; f += gridDim.x) {
    const uint32_t *c = c_packed + (size_t)f * g.tx_words;
    // every thread derives the frame's fade itself (one Philox call; no barrier)
    const Philox4 rh = philox_at(g.seed, STREAM_FADE, g.frame0 + f, 0);
    const float2 gh = gauss_pair(rh.x, rh.y);
    const float2 hf = make_float2(gh.x * 0.70710678118654752f, gh.y * 0.70710678118654752f);  // CN(0,1)
    if (threadIdx.x == 0) h_out[f] = hf;
    float2 *yf = y + (size_t)f * g.n_sym;
    for (int pp = threadIdx.x; pp < pairs; pp += CH_THREADS) {
      const Philox4 rn = philox_at(g.seed, STREAM_NOISE, g.frame0 + f, (uint32_t)pp);
      const float2 n0 = gauss_pair_fast(rn.x, rn.y), n1 = gauss_pair_fast(rn.z, rn.w);
      float2 *yo = yf + 2 * pp;
      const float2 x0 = map_symbol(g, c, 2 * pp);
      const float2 o0 = make_float2(hf.x * x0.x - hf.y * x0.y + g.sigma_over_sqrt2 * n0.x,
                                    hf.x * x0.y + hf.y * x0.x + g.sigma_over_sqrt2 * n0.y);
      if (2 * pp + 1 < g.n_sym) {
        const float2 x1 = map_symbol(g, c, 2 * pp + 1);
        const float2 o1 = make_float2(hf.x * x1.x - hf.y * x1.y + g.sigma_over_sqrt2 * n1.x,
                                      hf.x * x1.y + hf.y * x1.x + g.sigma_over_sqrt2 * n1.y);
        if ((reinterpret_cast<uintptr_t>(yo) & 15) == 0) {  // one 128-bit store for the pair
          *reinterpret_cast<float4 *>(yo) = make_float4(o0.x, o0.y, o1.x, o1.y);
        } else {
          yo[0] = o0;
          yo[1] = o1;
        }
      } else {
        yo[0] = o0;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ k-means
// kmldpc::KMeans::Run as COMPILED (src/kmeans.cc:15-84): anchor = first arg max |y|; every pass assigns
// each sample to its nearest centroid (first minimum) and ADDS to never-reset per-cluster sums; only cluster 0 feeds
// back: hhat = (sum_0 / cnt_0) / s_0, centroids = s_k * hhat; stop when the centroids repeat bit for bit.
// One CTA per frame, samples in registers, warp-shuffle + shared-memory block reductions, fp64 cumulative sums.
constexpr int KM_THREADS = 128;
template <int SPT>
__global__ void __launch_bounds__(KM_THREADS) kmeans_kernel(int B, const float2 *y, int n, const float2 *points, int q,
                                                           int iters, float2 *hhat_out, double2 *hhat64_out, int32_t *passes_out) {
  extern __shared__ float2 s_c[];  // [q] centroids
  __shared__ float s_red[3][KM_THREADS / 32];
  __shared__ unsigned long long s_best[KM_THREADS / 32];
  __shared__ double s_h[2];
  __shared__ int s_stop;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int f = blockIdx.x; f < B; f += gridDim.x) {
    const float2 *yf = y + (size_t)f * n;
    float2 ys[SPT];
    unsigned long long best = 0ull;
#pragma unroll
    for (int j = 0; j < SPT; j++) {
      const int i = j * KM_THREADS + tid;
      if (i < n) {
        ys[j] = yf[i];
        // |y|^2 is monotone in |y|; ties → smallest index (max_element returns the first maximum)
        const float a2 = ys[j].x * ys[j].x + ys[j].y * ys[j].y;
        const unsigned long long key = ((unsigned long long)__float_as_uint(a2) << 32) | (uint32_t)(0x7fffffff - i);
        best = key > best ? key : best;
      } else {
        ys[j] = make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    double cum_cnt = 0.0, cum_re = 0.0, cum_im = 0.0, hr = 0.0, hi = 0.0, prev_r = 0.0, prev_i = 0.0;
    bool have_prev = false;
    const double s0r = (double)__ldg(&points[0].x), s0i = (double)__ldg(&points[0].y);
    const double s0n = s0r * s0r + s0i * s0i;
    if (tid == 0) {
      unsigned long long b = s_best[0];
      for (int w = 1; w < KM_THREADS / 32; w++) b = s_best[w] > b ? s_best[w] : b;
      const int a = 0x7fffffff - (int)(uint32_t)(b & 0xffffffffu);
      const float2 ya = yf[a];
      hr = ((double)ya.x * s0r + (double)ya.y * s0i) / s0n;  // y_a / s_0
      hi = ((double)ya.y * s0r - (double)ya.x * s0i) / s0n;
      s_h[0] = hr;
      s_h[1] = hi;
      s_stop = 0;
    }
    __syncthreads();
    int passes = 0;
    for (int it = 0; it < iters; it++) {
      passes++;
      const float fhr = (float)s_h[0], fhi = (float)s_h[1];
      for (int k = tid; k < q; k += KM_THREADS) {
        const float2 s = __ldg(points + k);
        s_c[k] = make_float2(s.x * fhr - s.y * fhi, s.x * fhi + s.y * fhr);
      }
      __syncthreads();
      const float2 c0 = s_c[0];
      float cnt = 0.f, sr = 0.f, si = 0.f;
#pragma unroll
      for (int j = 0; j < SPT; j++) {
        const int i = j * KM_THREADS + tid;
        if (i < n) {
          const float dx0 = c0.x - ys[j].x, dy0 = c0.y - ys[j].y;
          const float d0 = dx0 * dx0 + dy0 * dy0;
          bool in0 = true;
          for (int k = 1; k < q; k++) {
            const float2 c = s_c[k];
            const float dx = c.x - ys[j].x, dy = c.y - ys[j].y;
            in0 = in0 && (d0 <= dx * dx + dy * dy);  // first minimum wins ties → cluster 0 keeps them
          }
          if (in0) {
            cnt += 1.f;
            sr += ys[j].x;
            si += ys[j].y;
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        si += __shfl_xor_sync(0xffffffffu, si, o);
      }
      if (lane == 0) {
        s_red[0][warp] = cnt;
        s_red[1][warp] = sr;
        s_red[2][warp] = si;
      }
      __syncthreads();
      if (tid == 0) {
        for (int w = 0; w < KM_THREADS / 32; w++) {
          cum_cnt += (double)s_red[0][w];
          cum_re += (double)s_red[1][w];
          cum_im += (double)s_red[2][w];
        }
        if (have_prev && prev_r == hr && prev_i == hi) {
          s_stop = 1;  // clusters_ == tempClusters (kmeans.cc:47-56): leave WITHOUT updating
        } else {
          prev_r = hr;
          prev_i = hi;
          have_prev = true;
          const double mr = cum_re / cum_cnt, mi = cum_im / cum_cnt;  // cluster 0 is never empty (anchor sample)
          hr = (mr * s0r + mi * s0i) / s0n;
          hi = (mi * s0r - mr * s0i) / s0n;
          s_h[0] = hr;
          s_h[1] = hi;
        }
      }
      __syncthreads();
      if (s_stop) break;
    }
    if (tid == 0) {
      hhat_out[f] = make_float2((float)s_h[0], (float)s_h[1]);
      if (hhat64_out) hhat64_out[f] = make_double2(s_h[0], s_h[1]);
      if (passes_out) passes_out[f] = passes;
    }
    __syncthreads();
  }
}

// Warp-per-frame variant (the fast path): no block barriers, samples in registers (lane l holds samples l, l+32, …),
// and the nearest-is-cluster-0 predicate evaluated only against the Voronoi neighbours of s_0 as half-plane tests
//   |y - c_0|^2 <= |y - c_k|^2   <=>   Re(conj(c_k - c_0) y) <= (|c_k|^2 - |c_0|^2) / 2 .
//
// EXACT ASSIGNMENTS.  north_star asks for centroids within 1e-4 of the reference on every frame, and one sample that
// falls on the other side of a cell boundary moves the (cumulative) mean by more than that.  So the estimate is carried
// in fp64 like the reference's, and the fp32 arithmetic is only a FILTER: d = Re(conj(a) y) - th is evaluated in packed
// fp32 with a bound tau on its own rounding error; |d| > tau decides the sample, and a pass in which any sample of the
// warp lands inside the band re-evaluates its membership in fp64 (a few per thousand passes).  The sums the reference
// accumulates sample by sample (kmeans.cc:44-45) are kept exact the cheap way: cluster 0's membership barely changes
// from one pass to the next, so the kernel keeps the fp64 sum of the CURRENT member set and corrects it by the samples
// that entered or left (fp64, from the exact input values) — the common pass adds nothing per sample.
// Every lane carries the same fp64 state (xor-shuffle reductions give all lanes the same value): no broadcasts.
constexpr int KMW_WARPS = 4;
// Blackwell packed fp32 (SASS FMUL2 / FFMA2): two neighbour half-plane tests per instruction
__device__ __forceinline__ float2 km_fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
// 1 / c for an integer-valued c in [1, 2^24): MUFU seed + two Newton steps in fp64 (relative error ~1e-16)
__device__ __forceinline__ double km_rcp_count(double c) {
  float rf;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)c));
  double r = (double)rf;
  r = fma(r, fma(-c, r, 1.0), r);
  r = fma(r, fma(-c, r, 1.0), r);
  return r;
}
// The exact input value of a sample for the rare fp64 blocks.  fp32 input: widened from the register through a volatile
// asm, so that the compiler cannot hoist 2 x SPL conversions out of those blocks into the hot loop's registers.
// fp64 input: re-read from memory (an L1 / L2 hit), again not hoistable.
__device__ __forceinline__ double2 km_widen(float2 v) {
  double x, y;
  asm volatile("cvt.f64.f32 %0, %1;" : "=d"(x) : "f"(v.x));
  asm volatile("cvt.f64.f32 %0, %1;" : "=d"(y) : "f"(v.y));
  return make_double2(x, y);
}
__device__ __forceinline__ double2 km_exact_load(const double2 *p) {
  double x, y;
  asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "l"(p));
  return make_double2(x, y);
}
__device__ __forceinline__ double km_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// (Measured and dropped: a first tier in plain fp32 with a margin test — frames that never came within the band are
// provably assigned like the exact kernel's — followed by this kernel on the rest.  The fp32 pass alone took 131 us per
// 16384 PEG2304 frames against 165 us for this kernel on all of them, but the 2.5 % of frames it handed over cost a 46 us
// tail (one warp per frame, twenty dependent passes: the tail is one frame's latency however few frames there are), and
// at -5 dB, where 40 % of the frames touch the band, the pair was slower than this kernel alone: 316 us against 277 us.)
template <int SPL>
struct KmMask {  // one membership bit per sample of the lane
  uint32_t lo = 0, hi = 0;
  __device__ __forceinline__ void set(int j, bool v) {
    if (v) {
      if (j < 32) lo |= 1u << j;
      else hi |= 1u << (j - 32);
    }
  }
  __device__ __forceinline__ bool get(int j) const { return j < 32 ? (lo >> j) & 1u : (hi >> (j - 32)) & 1u; }
  __device__ __forceinline__ bool any() const { return SPL > 32 ? (lo | hi) != 0 : lo != 0; }
};

template <int SPL, int MAXNB, bool F64IN>
// (CTAs per SM for 36 samples per lane: 4 measured best — 167 us per 16384 frames at 15 dB against 181 with 3 and 183 with 5)
__global__ void __launch_bounds__(KMW_WARPS * 32, SPL <= 16 ? 6 : (SPL <= 24 ? 5 : (SPL <= 36 ? 4 : (SPL <= 48 ? 3 : 2))))
kmeans_warp_kernel(int B, const void *y_in, int n, const KmConst kc, int iters, float2 *hhat_out, double2 *hhat64_out,
                   int32_t *passes_out, float2 *y32_out) {
  static_assert(MAXNB % 2 == 0 && SPL <= 64, "neighbours are tested in pairs; one mask bit per sample");
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ float2 km_copy[];  // [KMW_WARPS][SPL][32], fp32 input only
  const int lane = threadIdx.x & 31;
  const int wglobal = blockIdx.x * KMW_WARPS + (threadIdx.x >> 5), wstride = gridDim.x * KMW_WARPS;
  const float2 s0f = make_float2((float)kc.s0r, (float)kc.s0i);
  float2 snb[MAXNB];
#pragma unroll
  for (int t = 0; t < MAXNB; t++)
    snb[t] = make_float2((float)(kc.s0r + kc.dsr[t < kc.n_nb ? t : 0]), (float)(kc.s0i + kc.dsi[t < kc.n_nb ? t : 0]));
  for (int f = wglobal; f < B; f += wstride) {
    const float2 *yf = reinterpret_cast<const float2 *>(y_in) + (size_t)f * n;
    const double2 *yd = reinterpret_cast<const double2 *>(y_in) + (size_t)f * n;
    auto exact = [&](int j, float2 v) -> double2 {  // the input value itself, as the reference sees it
      return F64IN ? km_exact_load(yd + j * 32 + lane) : km_widen(v);
    };
    float2 ys[SPL];
    float2 *ysm = km_copy + (size_t)(threadIdx.x >> 5) * SPL * 32 + lane;  // fp32 input: a copy that CAN be indexed at run time
    unsigned long long best = 0ull;
    KmMask<SPL> valid;
#pragma unroll
    for (int j = 0; j < SPL; j++) {
      const int i = j * 32 + lane;
      if (i < n) {
        if (F64IN) {
          const double2 v = yd[i];
          ys[j] = make_float2((float)v.x, (float)v.y);
          if (y32_out) y32_out[(size_t)f * n + i] = ys[j];  // the fp32 copy the demapper reads
        } else {
          ys[j] = yf[i];
          ysm[j * 32] = ys[j];
        }
        valid.set(j, true);
        // |y|^2 is monotone in |y|; ties → smallest index (max_element returns the first maximum)
        const float a2 = fmaf(ys[j].y, ys[j].y, ys[j].x * ys[j].x);
        const unsigned long long key = ((unsigned long long)__float_as_uint(a2) << 32) | (uint32_t)(0x7fffffff - i);
        best = key > best ? key : best;
      } else {
        ys[j] = make_float2(3.0e15f, 1.0e15f);  // masked out of the membership below; far from every boundary (never "unsure")
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(FULL, best, o);
      best = other > best ? other : best;
    }
    int anchor = 0x7fffffff - (int)(uint32_t)(best & 0xffffffffu);
    const float amax2 = __uint_as_float((uint32_t)(best >> 32));
    {  // fp32 cannot order magnitudes closer than a few ulp: when a second sample is that close, order them in fp64
      const float thr = amax2 * (1.0f - 2.0e-6f);
      int near = 0;
#pragma unroll
      for (int j = 0; j < SPL; j++) near += (valid.get(j) && fmaf(ys[j].y, ys[j].y, ys[j].x * ys[j].x) >= thr) ? 1 : 0;
      const unsigned who = __ballot_sync(FULL, near > 0);
      if (__popc(who) > 1 || __any_sync(FULL, near > 1)) {
        double bd = -1.0;
        int bi = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < SPL; j++)
          if (valid.get(j) && fmaf(ys[j].y, ys[j].y, ys[j].x * ys[j].x) >= thr) {
            const double2 v = exact(j, ys[j]);
            const double d2 = fma(v.y, v.y, v.x * v.x);
            if (d2 > bd) { bd = d2; bi = j * 32 + lane; }  // ascending index within the lane: first maximum kept
          }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double od = __shfl_xor_sync(FULL, bd, o);
          const int oi = __shfl_xor_sync(FULL, bi, o);
          if (od > bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        anchor = bi;
      }
    }
    const float ymax_l1 = 1.4142137f * sqrtf(amax2);  // |y_x| + |y_y| <= sqrt(2) |y|
    double hr, hi;
    {
      const double2 ya = F64IN ? yd[anchor] : make_double2((double)yf[anchor].x, (double)yf[anchor].y);
      hr = ya.x * kc.is0r - ya.y * kc.is0i;  // y_a / s_0
      hi = ya.x * kc.is0i + ya.y * kc.is0r;
    }
    double cum_cnt = 0.0, cum_re = 0.0, cum_im = 0.0;  // cumulative over passes (never reset: kmeans.cc:33-34 as compiled)
    double set_re = 0.0, set_im = 0.0;                 // exact sums over the current member set of cluster 0
    int set_cnt = 0;
    KmMask<SPL> member;
    double prev_r = 0.0, prev_i = 0.0;
    bool have_prev = false;
    int passes = 0;
    for (int it = 0; it < iters; it++) {
      passes++;
      // ---- fp32 filter: d_t = Re(conj(c_t - c_0) y) - (|c_t|^2 - |c_0|^2) / 2, in cluster 0 iff every d_t <= 0
      const float fhr = (float)hr, fhi = (float)hi;
      const float2 c0 = make_float2(s0f.x * fhr - s0f.y * fhi, s0f.x * fhi + s0f.y * fhr);
      const float n0 = c0.x * c0.x + c0.y * c0.y;
      float2 ax[MAXNB / 2], ay[MAXNB / 2], nth[MAXNB / 2];
      float cmax2 = n0;
#pragma unroll
      for (int t = 0; t < MAXNB; t++) {
        const float2 ck = make_float2(snb[t].x * fhr - snb[t].y * fhi, snb[t].x * fhi + snb[t].y * fhr);
        const float nk = ck.x * ck.x + ck.y * ck.y;
        cmax2 = fmaxf(cmax2, nk);
        const bool used = t < kc.n_nb;
        const float axx = used ? ck.x - c0.x : 0.f, ayy = used ? ck.y - c0.y : 0.f;
        const float nt = used ? -0.5f * (nk - n0) : -3.0e38f;  // unused slot: never the maximum
        if (t & 1) { ax[t / 2].y = axx; ay[t / 2].y = ayy; nth[t / 2].y = nt; }
        else { ax[t / 2].x = axx; ay[t / 2].x = ayy; nth[t / 2].x = nt; }
      }
      // rounding of the filter: hhat, c_t, a_t, th_t and the two FMAs each within a few 6e-8 of |c| (|y| + |c|): the
      // band is >= 5x that bound
      const float cmax = sqrtf(cmax2);
      const float tau = 4.0e-6f * cmax * (ymax_l1 + cmax);
      // Membership = "every d_t <= 0" (first minimum wins ties → cluster 0 keeps them).  Per sample the filter keeps just the
      // SIGN BIT of max_t d_t (shifted into a word) and the smallest |max_t d_t| of the lane: a zero or a value inside the
      // band makes the pass "unsure" and the membership is then decided in fp64 below, so the sign bit is all fp32 has to say.
      KmMask<SPL> now;
      uint32_t sgn_a = 0, sgn_b = 0;
      float closest = 3.0e38f;
#pragma unroll
      for (int j = 0; j < SPL; j++) {
        float dmax;
#pragma unroll
        for (int t = 0; t < MAXNB / 2; t++) {
          const float2 v = km_fma2(ax[t], make_float2(ys[j].x, ys[j].x), km_fma2(ay[t], make_float2(ys[j].y, ys[j].y), nth[t]));
          dmax = t == 0 ? fmaxf(v.x, v.y) : fmaxf(dmax, fmaxf(v.x, v.y));
        }
        if (j < 32) sgn_a = __funnelshift_l(__float_as_uint(dmax), sgn_a, 1);
        else sgn_b = __funnelshift_l(__float_as_uint(dmax), sgn_b, 1);
        closest = fminf(closest, fabsf(dmax));
      }
      now.lo = __brev(sgn_a) >> (SPL >= 32 ? 0 : 32 - SPL);  // sample j of the lane at bit j
      (void)sgn_b;
      if constexpr (SPL > 32) now.hi = __brev(sgn_b) >> (SPL >= 64 ? 0 : 64 - SPL);
      const bool unsure = closest <= tau;
      if (__any_sync(FULL, unsure)) {  // rare: this pass's membership in fp64 from the fp64 estimate
        double ar[MAXNB], ai[MAXNB], th[MAXNB];
        const double h2 = hr * hr + hi * hi;
#pragma unroll
        for (int t = 0; t < MAXNB; t++) {
          ar[t] = kc.dsr[t] * hr - kc.dsi[t] * hi;
          ai[t] = kc.dsr[t] * hi + kc.dsi[t] * hr;
          th[t] = kc.dn[t] * h2;
        }
        now = KmMask<SPL>();
#pragma unroll
        for (int j = 0; j < SPL; j++) {
          if (!valid.get(j)) continue;
          const double2 v = exact(j, ys[j]);
          bool in0 = true;
#pragma unroll
          for (int t = 0; t < MAXNB; t++)
            if (t < kc.n_nb) in0 = in0 && (fma(ar[t], v.x, ai[t] * v.y) <= th[t]);
          now.set(j, in0);
        }
      }
      now.lo &= valid.lo;
      now.hi &= valid.hi;
      KmMask<SPL> chg;
      chg.lo = now.lo ^ member.lo;
      chg.hi = now.hi ^ member.hi;
      if (__any_sync(FULL, chg.any())) {  // samples that entered / left cluster 0 since the last pass
        double dr = 0.0, di = 0.0;
        int dc = 0;
        // A lane has a handful of changed samples at most: it walks the set bits of its mask (ascending, the order the sums
        // have always been taken in) and fetches each sample by its run-time index — from the shared-memory copy for fp32
        // input (the registers cannot be indexed), from global memory for fp64 input.
        auto walk = [&](uint32_t bits, uint32_t in_bits, int base) {
          while (bits) {
            const int j = __ffs(bits) - 1;
            bits &= bits - 1;
            const double2 v = F64IN ? km_exact_load(yd + (base + j) * 32 + lane) : km_widen(ysm[(base + j) * 32]);
            const bool in = (in_bits >> j) & 1u;
            dr += in ? v.x : -v.x;
            di += in ? v.y : -v.y;
            dc += in ? 1 : -1;
          }
        };
        walk(chg.lo, now.lo, 0);
        if (SPL > 32) walk(chg.hi, now.hi, 32);
        set_re += km_warp_sum(dr);
        set_im += km_warp_sum(di);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dc += __shfl_xor_sync(FULL, dc, o);
        set_cnt += dc;
        member = now;
      }
      cum_cnt += (double)set_cnt;
      cum_re += set_re;
      cum_im += set_im;
      if (have_prev && prev_r == hr && prev_i == hi) break;  // clusters_ == tempClusters (kmeans.cc:47-56)
      prev_r = hr;
      prev_i = hi;
      have_prev = true;
      const double inv = km_rcp_count(cum_cnt);  // cluster 0 is never empty: the anchor sample sits on c_0
      const double mr = cum_re * inv, mi = cum_im * inv;
      hr = mr * kc.is0r - mi * kc.is0i;
      hi = mr * kc.is0i + mi * kc.is0r;
    }
    if (lane == 0) {
      hhat_out[f] = make_float2((float)hr, (float)hi);
      if (hhat64_out) hhat64_out[f] = make_double2(hr, hi);
      if (passes_out) passes_out[f] = passes;
    }
  }
}

// ------------------------------------------------------------------------------------------------ demapper + resolver
// ModemLinearSystem::SoftAWGNDemodulation (modemlinearsystem.cc:51-79): p_k = softmax(-|s_k h - y|^2 / var), each
// clipped to [1e-12, 1-1e-12]; Modem::DeMapping (modem.cc:23-79) with priors 0.5: renormalise by the post-clip sum,
// bit marginals, clip.  Output = likelihood ratio P0/P1 (the clipped pair), MSB first within a symbol.
// KmCodec::GetMetrics/Metric/GetParityCheck (kmcodec.cc:105-163), hard non-5G metric: rr = (P0 > 0.5) ? 1 : 0
// (inverted on purpose), metric = number of unsatisfied rows; first argmin (kmcodec.cc:61-65).
constexpr int DM_THREADS = 256;
// One CTA per frame.  Each symbol is read once and demapped against all candidates; the inverted hard decisions of the
// candidates share one byte per variable (bit c = candidate c), so ONE pass over the Tanner graph yields the four
// syndrome weights.  With `winner_only` the four ratio vectors stay in shared memory and only the chosen candidate's
// goes to HBM (4x less write traffic, and the decoder needs no per-frame indirection).
__device__ __forceinline__ float dm_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float dm_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// (dx, dy) = s - y as one packed FADD2 (sm_100 add.rn.f32x2); s arrives from shared memory as a register pair
__device__ __forceinline__ float dm_dist2(const float2 s, const float2 ny) {
  float dx, dy;
  asm("{.reg .b64 a, b, d; mov.b64 a, {%2,%3}; mov.b64 b, {%4,%5}; add.rn.f32x2 d, a, b; mov.b64 {%0,%1}, d;}"
      : "=f"(dx), "=f"(dy) : "f"(s.x), "f"(s.y), "f"(ny.x), "f"(ny.y));
  return fmaf(dy, dy, dx * dx);
}

// One symbol against NC candidates: per-point softmax with the reference's clip, bit marginals, ratio P0/P1; returns in
// rr[j] the candidates' inverted hard decisions of bit j (bit c = candidate c).  lr_base already points at element
// i * BITS of candidate 0; consecutive candidates are lr_stride floats apart.
// SYM: every shipped constellation is mapped onto itself by a 90-degree turn (QPSK, 4PSK, square QAM), so the points of
// candidate c — s_k h e^{j c pi/2} — are the points of candidate 0 in another order: s_k e^{j c pi/2} = s_{perm_c(k)}.  The Q
// clipped probabilities are then the SAME numbers for all four candidates, only attached to other labels: they are
// computed once (Q exponentials instead of 4 Q), parked in this thread's column of a shared-memory stage and re-read
// through perm_c for candidates 1..3.  (The reference's rotation uses a truncated pi: its candidates differ from exact
// quarter turns by 1.6e-15 rad — nine orders below fp32.)  The host finds perm (kml_api.cu) and keeps the plain path for
// a constellation without the symmetry.
template <int BITS, int NC, bool SYM>
__device__ __forceinline__ void demap_symbol(const float2 yy, const float2 *s_pts, float rscale, float *lr_base,
                                             size_t lr_stride, unsigned int (&rr)[BITS], float *stage, const DemapParams &d) {
  constexpr int Q = 1 << BITS;
  const float2 ny = make_float2(-yy.x * rscale, -yy.y * rscale);
#pragma unroll
  for (int j = 0; j < BITS; j++) rr[j] = 0;
  float p[Q];
#pragma unroll
  for (int c = 0; c < NC; c++) {
    if (!SYM || c == 0) {
      float mx = -3.0e38f;
#pragma unroll
      for (int k = 0; k < Q; k++) {
        p[k] = -dm_dist2(s_pts[c * Q + k], ny);  // points and symbol are pre-scaled by sqrt(log2(e) / var)
        mx = fmaxf(mx, p[k]);
      }
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < Q; k++) {
        p[k] = dm_ex2(p[k] - mx);
        sum += p[k];
      }
      const float inv = dm_rcp(sum);
#pragma unroll
      for (int k = 0; k < Q; k++) p[k] = fmaxf(p[k] * inv, kSmallProbF);  // the upper clip 1-1e-12 is 1.0f in fp32
      // (the second normalisation, modem.cc:47-57, cancels in the ratio z0 / z1)
      if (SYM && NC > 1) {
#pragma unroll
        for (int k = 0; k < Q; k++) stage[k * DM_THREADS] = p[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < Q; k++) p[k] = stage[d.perm[c - 1][k] * DM_THREADS];
    }
    float ratio[BITS];
#pragma unroll
    for (int j = 0; j < BITS; j++) {
      float z0 = 0.f, z1 = 0.f;
#pragma unroll
      for (int k = 0; k < Q; k++) {
        if (((k >> (BITS - 1 - j)) & 1) == 0) z0 += p[k];
        else z1 += p[k];
      }
      ratio[j] = fminf(fmaxf(z0 * dm_rcp(z1), kLrMin), kLrMax);
      // rr = (P0 > 0.5) ? 1 : 0 — inverted on purpose (kmcodec.cc:110-115); taken from the ratio as stored, so that it
      // is exactly the complement of the decision the decoder makes from that ratio at iteration 0 (post > 1 ? 0 : 1)
      rr[j] |= (ratio[j] > 1.0f ? 1u : 0u) << c;
    }
    // a symbol's BITS ratios are contiguous and BITS * 4 bytes aligned (rows are n_sym * BITS floats): one wide store
    float *dst = lr_base + c * lr_stride;
    if constexpr (BITS == 2) *reinterpret_cast<float2 *>(dst) = make_float2(ratio[0], ratio[1]);
    else if constexpr (BITS == 4) *reinterpret_cast<float4 *>(dst) = make_float4(ratio[0], ratio[1], ratio[2], ratio[3]);
    else {
#pragma unroll
      for (int j = 0; j < BITS; j++) dst[j] = ratio[j];
    }
  }
}

// 4 points with quarter-turn symmetry (QPSK, 4PSK): candidate c's probabilities are candidate 0's in another order, and a
// bit splits the four points 2 | 2 — one of the three partitions {01|23}, {02|13}, {03|12}, in one of two orientations.
// Which one each (candidate, bit) reads is a property of the constellation file; CODE holds it (3 bits per entry 2c + j:
// 0 = 01|23, 1 = 23|01, 2 = 02|13, 3 = 13|02, 4 = 03|12, 5 = 12|03), found on the host from the label permutations
// (kml_api.cu) and compiled in for the two shipped files.  Four exponentials, at most six ratios, everything in registers —
// no staging through shared memory, no permutation look-ups.  Same arithmetic as demap_symbol otherwise.
constexpr uint32_t kQ4CodeQpsk = 0u | (2u << 3) | (3u << 6) | (0u << 9) | (1u << 12) | (3u << 15) | (2u << 18) | (1u << 21);
constexpr uint32_t kQ4Code4psk = 0u | (2u << 3) | (5u << 6) | (3u << 9) | (1u << 12) | (2u << 15) | (4u << 18) | (3u << 21);
template <uint32_t CODE, int NC>
__host__ __device__ constexpr bool q4_uses(int idx) {
  for (int e = 0; e < 2 * NC; e++)
    if (((CODE >> (3 * e)) & 7u) == (uint32_t)idx) return true;
  return false;
}
template <int NC, uint32_t CODE>
__device__ __forceinline__ void demap_symbol_q4(const float2 yy, const float2 *s_pts, float rscale, float *lr_base, size_t lr_stride,
                                                unsigned int (&rr)[2]) {
  const float2 ny = make_float2(-yy.x * rscale, -yy.y * rscale);
  float p[4], mx = -3.0e38f;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    p[k] = -dm_dist2(s_pts[k], ny);
    mx = fmaxf(mx, p[k]);
  }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    p[k] = dm_ex2(p[k] - mx);
    sum += p[k];
  }
  const float inv = dm_rcp(sum);
#pragma unroll
  for (int k = 0; k < 4; k++) p[k] = fmaxf(p[k] * inv, kSmallProbF);
  float r[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto ratio = [](float z0, float z1) { return fminf(fmaxf(z0 * dm_rcp(z1), kLrMin), kLrMax); };
  if constexpr (q4_uses<CODE, NC>(0) || q4_uses<CODE, NC>(1)) {
    const float a = p[0] + p[1], b = p[2] + p[3];
    if constexpr (q4_uses<CODE, NC>(0)) r[0] = ratio(a, b);
    if constexpr (q4_uses<CODE, NC>(1)) r[1] = ratio(b, a);
  }
  if constexpr (q4_uses<CODE, NC>(2) || q4_uses<CODE, NC>(3)) {
    const float a = p[0] + p[2], b = p[1] + p[3];
    if constexpr (q4_uses<CODE, NC>(2)) r[2] = ratio(a, b);
    if constexpr (q4_uses<CODE, NC>(3)) r[3] = ratio(b, a);
  }
  if constexpr (q4_uses<CODE, NC>(4) || q4_uses<CODE, NC>(5)) {
    const float a = p[0] + p[3], b = p[1] + p[2];
    if constexpr (q4_uses<CODE, NC>(4)) r[4] = ratio(a, b);
    if constexpr (q4_uses<CODE, NC>(5)) r[5] = ratio(b, a);
  }
  rr[0] = rr[1] = 0;
#pragma unroll
  for (int c = 0; c < NC; c++) {
    const float r0 = r[(CODE >> (3 * (2 * c))) & 7u], r1 = r[(CODE >> (3 * (2 * c + 1))) & 7u];
    *reinterpret_cast<float2 *>(lr_base + c * lr_stride) = make_float2(r0, r1);
    rr[0] |= (r0 > 1.0f ? 1u : 0u) << c;  // inverted on purpose, see demap_symbol
    rr[1] |= (r1 > 1.0f ? 1u : 0u) << c;
  }
}

// 64-point constellations WITHOUT the separable structure demap_symbol_grid64 needs (or under KML_DEMAP_NO_GRID=1): FOUR
// lanes share a symbol, 16 points each (lane t of the quad owns the points whose two
// top label bits are t).  One thread per symbol needs all 64 probabilities live (126 registers → 16 warps per SM, the
// kernel then runs at half its own instruction bound); a quad keeps 16, exchanges minimum / sum / the bit marginals
// with 24 shuffles per candidate, and lets 4x more warps be resident.  Same arithmetic as demap_symbol (the sums of a
// marginal are taken pairwise instead of sequentially).  Lane t finalises the bits j with j mod 4 == t.
// (Sharing the exponentials between the candidates, as demap_symbol<…, SYM> does, was measured here too: the permuted
// re-reads through a shared-memory stage cost more than the 48 exponentials they save — 3.65 ms against 3.34 ms per 16384
// PEG8064 frames — so the 64-point path keeps one softmax per candidate.)
template <int NC>
__device__ __forceinline__ void demap_symbol_quad64(const float2 yy, const float2 *s_pts, float rscale, float *lr_base,
                                                    size_t lr_stride, int t, bool valid, unsigned int &rr_a,
                                                    unsigned int &rr_b) {
  constexpr unsigned FULL = 0xffffffffu;
  const float2 ny = make_float2(-yy.x * rscale, -yy.y * rscale);
  rr_a = rr_b = 0;
#pragma unroll
  for (int c = 0; c < NC; c++) {
    float p[16];
    float mn = 3.0e38f;
#pragma unroll
    for (int m = 0; m < 16; m++) {
      p[m] = dm_dist2(s_pts[c * 68 + 17 * t + m], ny);  // chunks padded to 17: the quad's four addresses hit distinct banks
      mn = fminf(mn, p[m]);
    }
    mn = fminf(mn, __shfl_xor_sync(FULL, mn, 1));
    mn = fminf(mn, __shfl_xor_sync(FULL, mn, 2));
    float sum = 0.f;
#pragma unroll
    for (int m = 0; m < 16; m++) {
      p[m] = dm_ex2(mn - p[m]);
      sum += p[m];
    }
    sum += __shfl_xor_sync(FULL, sum, 1);
    sum += __shfl_xor_sync(FULL, sum, 2);
    const float inv = dm_rcp(sum);
#pragma unroll
    for (int m = 0; m < 16; m++) p[m] = fmaxf(p[m] * inv, kSmallProbF);
    // label bit j of point k = (k >> (5 - j)) & 1 with k = 16 t + m: bits 0, 1 come from t, bits 2..5 from m (MSB first)
    float z0[6], z1[6];
    float q8[8], q4[4], q2[2];
    {  // bit 5 = m & 1
      float e = p[0], o = p[1];
#pragma unroll
      for (int i = 1; i < 8; i++) { e += p[2 * i]; o += p[2 * i + 1]; }
      z0[5] = e; z1[5] = o;
#pragma unroll
      for (int i = 0; i < 8; i++) q8[i] = p[2 * i] + p[2 * i + 1];
    }
    {  // bit 4 = (m >> 1) & 1
      float e = q8[0], o = q8[1];
#pragma unroll
      for (int i = 1; i < 4; i++) { e += q8[2 * i]; o += q8[2 * i + 1]; }
      z0[4] = e; z1[4] = o;
#pragma unroll
      for (int i = 0; i < 4; i++) q4[i] = q8[2 * i] + q8[2 * i + 1];
    }
    z0[3] = q4[0] + q4[2]; z1[3] = q4[1] + q4[3];  // bit 3 = (m >> 2) & 1
    q2[0] = q4[0] + q4[1]; q2[1] = q4[2] + q4[3];
    z0[2] = q2[0]; z1[2] = q2[1];                  // bit 2 = (m >> 3) & 1
    const float mine = q2[0] + q2[1];              // this lane's 16 points
#pragma unroll
    for (int j = 2; j < 6; j++) {
      z0[j] += __shfl_xor_sync(FULL, z0[j], 1); z0[j] += __shfl_xor_sync(FULL, z0[j], 2);
      z1[j] += __shfl_xor_sync(FULL, z1[j], 1); z1[j] += __shfl_xor_sync(FULL, z1[j], 2);
    }
    {  // bit 0 = t >> 1: lanes {0,1} against {2,3};  bit 1 = t & 1: lanes {0,2} against {1,3}
      const float pair01 = mine + __shfl_xor_sync(FULL, mine, 1), other01 = __shfl_xor_sync(FULL, pair01, 2);
      z0[0] = (t & 2) ? other01 : pair01; z1[0] = (t & 2) ? pair01 : other01;
      const float pair02 = mine + __shfl_xor_sync(FULL, mine, 2), other02 = __shfl_xor_sync(FULL, pair02, 1);
      z0[1] = (t & 1) ? other02 : pair02; z1[1] = (t & 1) ? pair02 : other02;
    }
    // lane t writes bits t and t + 4 (the latter for t < 2)
    float a0 = z0[0], a1 = z1[0], b0 = z0[4], b1 = z1[4];
    if (t == 1) { a0 = z0[1]; a1 = z1[1]; b0 = z0[5]; b1 = z1[5]; }
    if (t == 2) { a0 = z0[2]; a1 = z1[2]; }
    if (t == 3) { a0 = z0[3]; a1 = z1[3]; }
    const float ra = fminf(fmaxf(a0 * dm_rcp(a1), kLrMin), kLrMax), rb = fminf(fmaxf(b0 * dm_rcp(b1), kLrMin), kLrMax);
    if (valid) {
      lr_base[c * lr_stride + t] = ra;
      if (t < 2) lr_base[c * lr_stride + t + 4] = rb;
    }
    rr_a |= (ra > 1.0f ? 1u : 0u) << c;  // rr = (P0 > 0.5) ? 1 : 0 — inverted on purpose (kmcodec.cc:110-115), see demap_symbol
    rr_b |= (rb > 1.0f ? 1u : 0u) << c;
  }
}

// 64QAM on a square grid with a Gray mapping that splits the label into an in-phase and a quadrature half (the shipped
// 6bits_64QAM_Gray.txt; the host verifies the structure point by point and otherwise keeps demap_symbol_quad64): ONE
// thread per symbol, ALL candidates from 16 exponentials.
//   * With z = y / h the exponent -|s h - y|^2 / var = -(|h|^2 / var) ((Re s - Re z)^2 + (Im s - Im z)^2) separates, so the
//     normalised point probabilities are products  p_(i,j) = a_i b_j  of two 8-point softmaxes (levels against Re z and
//     against Im z): 16 exponentials instead of 64.
//   * The reference clips every p_k to >= 1e-12 BEFORE the bit marginals (modemlinearsystem.cc:78, modem.cc:26-27), which
//     does not separate — so the 8 x 8 grid of max(a_i b_j, 1e-12) is formed (one FMUL + FMNMX per point) and reduced to its
//     row sums R_i and column sums C_j; a bit of the in-phase half is a ratio of sums of R, of the quadrature half of C.
//   * Candidate c divides by h e^{j c pi/2}: z turns by -c quarter turns, which on a symmetric grid only swaps / reverses the
//     roles of the two axes: (I, Q) marginals = (R, C), (C, rev R), (rev R, rev C), (rev C, R) for c = 0..3.  Reversal
//     leaves the label sets {0,1,6,7}|{2..5} and {0,3,4,7}|{1,2,5,6} of the two lower Gray bits in place and swaps the
//     halves of the top bit: eight ratios serve all 24 (candidate, bit) outputs.
// Label layout compiled in (levels ascending, MSB first): bit 0 = level >= 4, bit 1 = level in 2..5, bit 2 = level in
// {1,2,5,6} of the in-phase axis; bit 3 = level < 4, bits 4, 5 like 1, 2 of the quadrature axis.
struct GridSums {
  float t0, t1, m0, m1, l0, l1;  // levels 0-3 | 4-7, {0,1,6,7} | {2..5}, {0,3,4,7} | {1,2,5,6}
};
__device__ __forceinline__ GridSums grid_sums(const float (&x)[8]) {
  const float p01 = x[0] + x[1], p23 = x[2] + x[3], p45 = x[4] + x[5], p67 = x[6] + x[7];
  GridSums g;
  g.t0 = p01 + p23; g.t1 = p45 + p67;
  g.m0 = p01 + p67; g.m1 = p23 + p45;
  g.l0 = (x[0] + x[3]) + (x[4] + x[7]); g.l1 = (x[1] + x[2]) + (x[5] + x[6]);
  return g;
}
__device__ __forceinline__ float grid_ratio(float z0, float z1) { return fminf(fmaxf(z0 * dm_rcp(z1), kLrMin), kLrMax); }

template <int NC>
__device__ __forceinline__ void demap_symbol_grid64(const float2 yy, const float2 hb, float inv_var, const float (&lv)[8],
                                                    float *lr_base, size_t lr_stride, unsigned int (&rr)[6]) {
  const float h2 = fmaf(hb.x, hb.x, hb.y * hb.y);
  const bool live = h2 > 1.0e-30f;  // (h = 0: every point coincides, the softmax is uniform)
  const float ih2 = live ? dm_rcp(h2) : 0.0f;
  const float zr = (yy.x * hb.x + yy.y * hb.y) * ih2, zi = (yy.y * hb.x - yy.x * hb.y) * ih2;
  const float g = live ? h2 * inv_var * 1.4426950408889634f : 0.0f;  // exponents in bits: ex2 below
  float a[8], b[8], ma = -3.0e38f, mb = -3.0e38f;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const float dr = lv[i] - zr, di = lv[i] - zi;
    a[i] = -g * dr * dr;
    b[i] = -g * di * di;
    ma = fmaxf(ma, a[i]);
    mb = fmaxf(mb, b[i]);
  }
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = dm_ex2(a[i] - ma);
    b[i] = dm_ex2(b[i] - mb);
    sa += a[i];
    sb += b[i];
  }
  const float ia = dm_rcp(sa), ib = dm_rcp(sb);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] *= ia;
    b[i] *= ib;
  }
  float R[8], C[8];
#pragma unroll
  for (int i = 0; i < 8; i++) R[i] = C[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float t = fmaxf(a[i] * b[j], kSmallProbF);  // the reference's per-point clip (the upper one is 1.0f in fp32)
      R[i] += t;
      C[j] += t;
    }
  const GridSums sr = grid_sums(R), sc = grid_sums(C);
  const float rt_r = grid_ratio(sr.t0, sr.t1), rti_r = grid_ratio(sr.t1, sr.t0), rm_r = grid_ratio(sr.m0, sr.m1),
              rl_r = grid_ratio(sr.l0, sr.l1);
  float ratio[NC][6];
  ratio[0][0] = rt_r; ratio[0][1] = rm_r; ratio[0][2] = rl_r;
  {
    const float rti_c = grid_ratio(sc.t1, sc.t0), rm_c = grid_ratio(sc.m0, sc.m1), rl_c = grid_ratio(sc.l0, sc.l1);
    ratio[0][3] = rti_c; ratio[0][4] = rm_c; ratio[0][5] = rl_c;
    if constexpr (NC == 4) {
      const float rt_c = grid_ratio(sc.t0, sc.t1);
      ratio[1][0] = rt_c;  ratio[1][1] = rm_c; ratio[1][2] = rl_c; ratio[1][3] = rt_r;  ratio[1][4] = rm_r; ratio[1][5] = rl_r;
      ratio[2][0] = rti_r; ratio[2][1] = rm_r; ratio[2][2] = rl_r; ratio[2][3] = rt_c;  ratio[2][4] = rm_c; ratio[2][5] = rl_c;
      ratio[3][0] = rti_c; ratio[3][1] = rm_c; ratio[3][2] = rl_c; ratio[3][3] = rti_r; ratio[3][4] = rm_r; ratio[3][5] = rl_r;
    }
  }
#pragma unroll
  for (int j = 0; j < 6; j++) rr[j] = 0;
#pragma unroll
  for (int c = 0; c < NC; c++) {
    float2 *dst = reinterpret_cast<float2 *>(lr_base + c * lr_stride);  // 24 bytes per symbol: 8-byte aligned
    dst[0] = make_float2(ratio[c][0], ratio[c][1]);
    dst[1] = make_float2(ratio[c][2], ratio[c][3]);
    dst[2] = make_float2(ratio[c][4], ratio[c][5]);
#pragma unroll
    for (int j = 0; j < 6; j++) rr[j] |= (ratio[c][j] > 1.0f ? 1u : 0u) << c;  // inverted on purpose, see demap_symbol
  }
}

// The same separation for 16 points (4 x 4 grid, the shipped 4bit_16QAM_Gray.txt): 8 exponentials for all candidates.  Label layout
// compiled in (levels ascending, MSB first): bit 0 = quadrature level < 2, bit 1 = quadrature level in {1, 2}, bits 2, 3 the
// same of the in-phase level.  (phi1 / phi2 do not split into an in-phase and a quadrature half: general path.)
template <int NC>
__device__ __forceinline__ void demap_symbol_grid16(const float2 yy, const float2 hb, float inv_var, const float (&lv)[8],
                                                    float *lr_base, size_t lr_stride, unsigned int (&rr)[4]) {
  const float h2 = fmaf(hb.x, hb.x, hb.y * hb.y);
  const bool live = h2 > 1.0e-30f;
  const float ih2 = live ? dm_rcp(h2) : 0.0f;
  const float zr = (yy.x * hb.x + yy.y * hb.y) * ih2, zi = (yy.y * hb.x - yy.x * hb.y) * ih2;
  const float g = live ? h2 * inv_var * 1.4426950408889634f : 0.0f;
  float a[4], b[4], ma = -3.0e38f, mb = -3.0e38f;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float dr = lv[i] - zr, di = lv[i] - zi;
    a[i] = -g * dr * dr;
    b[i] = -g * di * di;
    ma = fmaxf(ma, a[i]);
    mb = fmaxf(mb, b[i]);
  }
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    a[i] = dm_ex2(a[i] - ma);
    b[i] = dm_ex2(b[i] - mb);
    sa += a[i];
    sb += b[i];
  }
  const float ia = dm_rcp(sa), ib = dm_rcp(sb);
  float R[4] = {0.f, 0.f, 0.f, 0.f}, C[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    a[i] *= ia;
    b[i] *= ib;
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float t = fmaxf(a[i] * b[j], kSmallProbF);  // the reference's per-point clip
      R[i] += t;
      C[j] += t;
    }
  const float rlo = R[0] + R[1], rhi = R[2] + R[3], rout = R[0] + R[3], rin = R[1] + R[2];
  const float clo = C[0] + C[1], chi = C[2] + C[3], cout = C[0] + C[3], cin = C[1] + C[2];
  const float rti_r = grid_ratio(rhi, rlo), rm_r = grid_ratio(rout, rin), rti_c = grid_ratio(chi, clo), rm_c = grid_ratio(cout, cin);
  float ratio[NC][4];
  ratio[0][0] = rti_c; ratio[0][1] = rm_c; ratio[0][2] = rti_r; ratio[0][3] = rm_r;
  if constexpr (NC == 4) {
    const float rt_r = grid_ratio(rlo, rhi), rt_c = grid_ratio(clo, chi);
    ratio[1][0] = rt_r;  ratio[1][1] = rm_r; ratio[1][2] = rti_c; ratio[1][3] = rm_c;
    ratio[2][0] = rt_c;  ratio[2][1] = rm_c; ratio[2][2] = rt_r;  ratio[2][3] = rm_r;
    ratio[3][0] = rti_r; ratio[3][1] = rm_r; ratio[3][2] = rt_c;  ratio[3][3] = rm_c;
  }
#pragma unroll
  for (int j = 0; j < 4; j++) rr[j] = 0;
#pragma unroll
  for (int c = 0; c < NC; c++) {
    *reinterpret_cast<float4 *>(lr_base + c * lr_stride) = make_float4(ratio[c][0], ratio[c][1], ratio[c][2], ratio[c][3]);
#pragma unroll
    for (int j = 0; j < 4; j++) rr[j] |= (ratio[c][j] > 1.0f ? 1u : 0u) << c;  // inverted on purpose, see demap_symbol
  }
}

template <int BITS, int NC, bool SYM, uint32_t Q4 = 0xFFFFFFFFu>  // Q4 != all-ones: BITS == 2, the partition code of demap_symbol_q4
__global__ void __launch_bounds__(DM_THREADS) demap_kernel(const DemapParams d) {
  constexpr int Q = 1 << BITS;
  constexpr int MAXS = 6;  // symbols a thread keeps in flight (n_sym <= MAXS * DM_THREADS on the fast path)
  extern __shared__ unsigned char dsm[];
  constexpr int PTS = BITS == 6 ? 68 : Q;  // 64 points: four chunks of 16 padded to 17 (see demap_symbol_quad64)
  float2 *s_pts = reinterpret_cast<float2 *>(dsm);                          // [NC][PTS] s_k * h_cand
  float *s_lr = reinterpret_cast<float *>(dsm + sizeof(float2) * PTS * 4);   // [NC][n_tx]   (winner_only)
  unsigned char *s_rr = dsm + sizeof(float2) * PTS * 4 + (d.winner_only ? sizeof(float) * 4 * (size_t)d.n_tx : 0);
  // SYM: [Q][DM_THREADS] probabilities of candidate 0, one column per thread (16-byte aligned behind the decision bytes)
  float *stage = reinterpret_cast<float *>(dsm + ((sizeof(float2) * PTS * 4 + (d.winner_only ? sizeof(float) * 4 * (size_t)d.n_tx : 0) +
                                                   d.n_tx + d.punct + 32 + 15) & ~(size_t)15)) + threadIdx.x;
  __shared__ int s_cnt[4], s_best;
  const int tid = threadIdx.x;
  // exp(-|s - y|^2 / var) = 2^(-|r s - r y|^2) with r = sqrt(log2(e) / var): the candidate tables and the symbols are
  // pre-scaled, which leaves FADD2 + FMUL + FFMA per point
  const float rscale = sqrtf(d.inv_var * 1.4426950408889634f);
  const int zero_slot = d.punct + d.n_tx;               // a byte that is always 0 (padding of the ELL column table)
  const size_t lr_stride = (size_t)d.n_tx;
  for (int f = blockIdx.x; f < d.B; f += gridDim.x) {
    const float2 hb = d.h[f];
    const float2 *yf = d.y + (size_t)f * d.n_sym;
    float2 yreg[MAXS];
#pragma unroll
    for (int u = 0; u < MAXS; u++) {  // all of this thread's symbols are in flight before the first one is used
      const int i = u * DM_THREADS + tid;
      yreg[u] = (BITS != 6 && i < d.n_sym) ? yf[i] : make_float2(0.f, 0.f);
    }
    if (tid < 4) s_cnt[tid] = 0;
    for (int i = tid; i < NC * Q; i += DM_THREADS) {
      const int c = i / Q;
      const float2 r = c == 0 ? d.rot[0] : (c == 1 ? d.rot[1] : (c == 2 ? d.rot[2] : d.rot[3]));
      const float2 s = __ldg(d.points + (i % Q));
      const float2 hc = make_float2(hb.x * r.x - hb.y * r.y, hb.x * r.y + hb.y * r.x);
      const int k = i % Q;
      s_pts[BITS == 6 ? c * PTS + k + (k >> 4) : i] =
          make_float2((s.x * hc.x - s.y * hc.y) * rscale, (s.x * hc.y + s.y * hc.x) * rscale);
    }
    if (d.hard_metric) {
      for (int i = tid; i < d.punct; i += DM_THREADS) s_rr[i] = 0;
      if (tid == 0) s_rr[zero_slot] = 0;
    }
    __syncthreads();
    float *lr0 = d.winner_only ? s_lr : d.lr + (size_t)f * NC * d.n_tx;
    if (BITS == 6 && d.grid64) {  // separable grid: a thread per symbol, all candidates at once
      if constexpr (BITS == 6) {
        for (int i = tid; i < d.n_sym; i += DM_THREADS) {
          unsigned int rr[6];
          demap_symbol_grid64<NC>(yf[i], hb, d.inv_var, d.levels, lr0 + i * BITS, lr_stride, rr);
          if (d.hard_metric) {
#pragma unroll
            for (int j = 0; j < 6; j++) s_rr[d.punct + i * BITS + j] = (unsigned char)rr[j];
          }
        }
      }
    } else if constexpr (BITS == 6) {  // a quad of lanes per symbol
      const int t = tid & 3;
      for (int i0 = 0; i0 < d.n_sym; i0 += DM_THREADS / 4) {
        const int i = i0 + (tid >> 2);
        const bool valid = i < d.n_sym;
        const float2 yy = yf[valid ? i : d.n_sym - 1];
        unsigned int rr_a, rr_b;
        demap_symbol_quad64<NC>(yy, s_pts, rscale, lr0 + i * BITS, lr_stride, t, valid, rr_a, rr_b);
        if (d.hard_metric && valid) {
          s_rr[d.punct + i * BITS + t] = (unsigned char)rr_a;
          if (t < 2) s_rr[d.punct + i * BITS + t + 4] = (unsigned char)rr_b;
        }
      }
    } else
#pragma unroll
    for (int u = 0; u < MAXS; u++) {
      const int i = u * DM_THREADS + tid;
      if (i < d.n_sym) {
        unsigned int rr[BITS];
        if constexpr (Q4 != 0xFFFFFFFFu) demap_symbol_q4<NC, Q4>(yreg[u], s_pts, rscale, lr0 + i * BITS, lr_stride, rr);
        else if constexpr (BITS == 4) {
          if (d.grid16) demap_symbol_grid16<NC>(yreg[u], hb, d.inv_var, d.levels, lr0 + i * BITS, lr_stride, rr);
          else demap_symbol<BITS, NC, SYM>(yreg[u], s_pts, rscale, lr0 + i * BITS, lr_stride, rr, stage, d);
        } else demap_symbol<BITS, NC, SYM>(yreg[u], s_pts, rscale, lr0 + i * BITS, lr_stride, rr, stage, d);
        if (d.hard_metric) {
#pragma unroll
          for (int j = 0; j < BITS; j++) s_rr[d.punct + i * BITS + j] = (unsigned char)rr[j];
        }
      }
    }
    for (int i = MAXS * DM_THREADS + tid; BITS != 6 && i < d.n_sym; i += DM_THREADS) {  // very long frames
      unsigned int rr[BITS];
      if constexpr (Q4 != 0xFFFFFFFFu) demap_symbol_q4<NC, Q4>(yf[i], s_pts, rscale, lr0 + i * BITS, lr_stride, rr);
      else if constexpr (BITS == 4) {
        if (d.grid16) demap_symbol_grid16<NC>(yf[i], hb, d.inv_var, d.levels, lr0 + i * BITS, lr_stride, rr);
        else demap_symbol<BITS, NC, SYM>(yf[i], s_pts, rscale, lr0 + i * BITS, lr_stride, rr, stage, d);
      } else demap_symbol<BITS, NC, SYM>(yf[i], s_pts, rscale, lr0 + i * BITS, lr_stride, rr, stage, d);
      if (d.hard_metric) {
#pragma unroll
        for (int j = 0; j < BITS; j++) s_rr[d.punct + i * BITS + j] = (unsigned char)rr[j];
      }
    }
    if (d.hard_metric) {
      __syncthreads();
      int bad[4] = {0, 0, 0, 0};
      // ELL table, transposed: consecutive rows read consecutive 16-bit words.  Per-candidate counts are kept packed,
      // one byte each (a thread sees at most m_rows / DM_THREADS <= 255 rows), and split once at the end.
      unsigned int bad4 = 0;
      if (d.ell_width == 6) {  // (3,6)-regular codes: unrolled, all six table words in flight
        const uint16_t *ce = d.col_ell;
        const int mr = d.m_rows;
        for (int rrow = tid; rrow < mr; rrow += DM_THREADS) {
          unsigned int idx[6], par = 0;
#pragma unroll
          for (int k = 0; k < 6; k++) idx[k] = __ldg(ce + k * mr + rrow);
#pragma unroll
          for (int k = 0; k < 6; k++) par ^= s_rr[idx[k]];
          bad4 += (par & 1u) | ((par & 2u) << 7) | ((par & 4u) << 14) | ((par & 8u) << 21);
        }
      } else {
        for (int rrow = tid; rrow < d.m_rows; rrow += DM_THREADS) {
          unsigned int par = 0;
          for (int k = 0; k < d.ell_width; k++) par ^= s_rr[__ldg(d.col_ell + (size_t)k * d.m_rows + rrow)];
          bad4 += (par & 1u) | ((par & 2u) << 7) | ((par & 4u) << 14) | ((par & 8u) << 21);
          if ((bad4 & 0x80808080u) != 0) {  // a byte is about to overflow (very long codes): spill into the wide counters
#pragma unroll
            for (int c = 0; c < 4; c++) bad[c] += (bad4 >> (8 * c)) & 0xFFu;
            bad4 = 0;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 4; c++) bad[c] += (bad4 >> (8 * c)) & 0xFFu;
#pragma unroll
      for (int c = 0; c < 4; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bad[c] += __shfl_xor_sync(0xffffffffu, bad[c], o);
        if ((tid & 31) == 0 && bad[c]) atomicAdd(&s_cnt[c], bad[c]);
      }
      __syncthreads();
      if (tid == 0) {
        int best = 0;
        for (int c = 0; c < NC; c++) {
          d.metric[(size_t)f * 4 + c] = (float)s_cnt[c];
          if (s_cnt[c] < s_cnt[best]) best = c;  // std::min_element: first minimum (kmcodec.cc:61-65)
        }
        for (int c = NC; c < 4; c++) d.metric[(size_t)f * 4 + c] = 0.f;
        d.kstar[f] = best;
        s_best = best;
      }
      if (d.winner_only) {
        __syncthreads();
        const int best = s_best;
        if (d.skip_decode && s_cnt[best] == 0) {
          // The chosen candidate's hard decisions already satisfy every check: the decoder would make exactly these
          // decisions in its first variable phase (all messages neutral → posterior = channel ratio), find a zero
          // syndrome and return 0 + (0 < max_iter) = 1 (binaryldpccodec.cc:175-232).  Its answer is written here and the
          // frame never enters the decoder's queue — nor do its ratios travel through HBM.  (All row degrees are even,
          // so the inverted decisions of the metric and their complements have the same syndrome.)
          for (int v = tid; v < d.words_n * 32; v += DM_THREADS) {
            const unsigned bit = v < d.n_tx ? (((unsigned)s_rr[v] >> best) & 1u) ^ 1u : 0u;
            const unsigned word = __ballot_sync(0xffffffffu, bit);
            if ((tid & 31) == 0) d.out_bits[(size_t)f * d.words_n + (v >> 5)] = word;
          }
          if (tid == 0) d.out_ret[f] = 1;
        } else {
          const float *w = s_lr + (size_t)best * d.n_tx;
          float *out = d.lr + (size_t)f * d.n_tx;
          for (int i = tid; i < d.n_tx; i += DM_THREADS) out[i] = w[i];
          if (tid == 0 && d.queue) {  // longest first: a frame this far from a codeword will run all its iterations
            if (s_cnt[best] > d.long_metric) d.queue[atomicAdd(d.queue_n, 1)] = f;
            else d.queue[d.queue_cap - 1 - atomicAdd(d.queue_n + 1, 1)] = f;
          }
        }
      }
    }
    __syncthreads();
  }
}

// BinaryLDPCCodec::ParityCheck (binaryldpccodec.cc:281-299) on bit-packed decisions; one CTA per frame.
__global__ void syndrome_weight_kernel(int F, const uint32_t *bits, int words_n, int m_rows, const int32_t *row_ptr,
                                       const int32_t *col_idx, float *metric) {
  __shared__ int s_cnt;
  for (int f = blockIdx.x; f < F; f += gridDim.x) {
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const uint32_t *b = bits + (size_t)f * words_n;
    int bad = 0;
    for (int r = threadIdx.x; r < m_rows; r += blockDim.x) {
      uint32_t par = 0;
      for (int e = __ldg(row_ptr + r); e < __ldg(row_ptr + r + 1); e++) {
        const int c = __ldg(col_idx + e);
        par ^= (__ldg(b + (c >> 5)) >> (c & 31));
      }
      bad += par & 1u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(&s_cnt, bad);
    __syncthreads();
    if (threadIdx.x == 0) metric[f] = (float)s_cnt;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ soft-syndrome chain
// The soft metric (kmcodec.cc:146-155) sums ln(syndrom_soft_[r]) AFTER Decoder(metric_iter) — but the decoder writes
// syndrom_soft_ only in its check-node phase (binaryldpccodec.cc:274), so a decode that leaves at iteration 0
// (binaryldpccodec.cc:231-232) leaves the array as the PREVIOUS Decoder call left it: the previous candidate of the
// frame, or the previous frame's final decode, or whatever that one inherited.  Only the sum is ever read, so the state
// is one number per codec.  This kernel walks that chain for a batch: frame f's candidate 0 needs the state after
// frame f-1, which needs f-1's final decode, which needs f-1's choice …  Frames whose candidate 0 wrote its own value
// depend on nothing and resolve at once; the rest resolve as their predecessor's state becomes known — immediately
// when that final decode leaves at iteration 0 too, else after the host has run it (one round per link of a chain of
// consecutive dependent frames).  One CTA; flags[f]: bit 0 = state known, bits 8.. = 1 + round the frame was queued in.
__global__ void __launch_bounds__(1024) soft_chain_kernel(SoftChainParams p, int round) {
  __shared__ int s_q, s_left;
  const int tid = threadIdx.x;
  if (tid == 0) { s_q = 0; s_left = 0; }
  __syncthreads();
  volatile int32_t *flags = p.flags;
  volatile double *state = p.state;
  volatile int32_t *kstar = p.kstar;
  for (bool again = true; again;) {
    int progress = 0;
    for (int f = tid; f < p.B; f += blockDim.x) {
      const int fl = flags[f];
      if (fl & 1) continue;
      if (kstar[f] >= 0) {  // chosen in an earlier round: its final decode has run by now
        if ((fl >> 8) != 0 && (fl >> 8) <= round) {
          if (p.fret[f] != 1) state[f] = p.fown[f];  // (else: the value after the fourth candidate, already there)
          __threadfence_block();
          flags[f] = fl | 1;
          progress = 1;
        }
        continue;
      }
      const int32_t *mr = p.mret + 4 * (size_t)f;
      double m[4];
      if (mr[0] == 1) {  // candidate 0 left at t = 0: inherits the state the previous frame left
        if (f == 0) m[0] = *p.carry;
        else if (flags[f - 1] & 1) { __threadfence_block(); m[0] = state[f - 1]; }
        else continue;
      } else {
        m[0] = p.own[4 * (size_t)f];
      }
      int best = 0;
#pragma unroll
      for (int c = 1; c < 4; c++) m[c] = mr[c] == 1 ? m[c - 1] : p.own[4 * (size_t)f + c];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        p.metric[4 * (size_t)f + c] = (float)fabs(m[c]);  // kmcodec.cc:137
        if (fabs(m[c]) < fabs(m[best])) best = c;          // first minimum (kmcodec.cc:61-65)
      }
      state[f] = m[3];
      kstar[f] = best;
      int nf = fl;
      if (p.final_decode) {
        p.queue[atomicAdd(&s_q, 1)] = f;
        nf |= (round + 1) << 8;
      }
      __threadfence_block();
      if (!p.final_decode || mr[best] == 1) nf |= 1;  // no final decode, or it will leave at t = 0 as well
      flags[f] = nf;
      progress = 1;
    }
    again = __syncthreads_or(progress) != 0;
  }
  int left = 0;
  for (int f = tid; f < p.B; f += blockDim.x) left += (flags[f] & 1) ? 0 : 1;
  if (left) atomicAdd(&s_left, left);
  __syncthreads();
  if (tid == 0) {
    p.counts[0] = s_q;
    p.counts[1] = s_left;
    if (s_left == 0 && p.B > 0) *p.carry = state[p.B - 1];
  }
}

__global__ void abs_kernel(int n, float *v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = fabsf(v[i]);
}

__global__ void argmin4_kernel(int B, const float *metric, int32_t *kstar) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  int best = 0;
  for (int c = 1; c < 4; c++)
    if (metric[f * 4 + c] < metric[f * 4 + best]) best = c;  // std::min_element: first minimum
  kstar[f] = best;
}

// Metric decodes that already hold the answer (5G-style metric: Decoder(metric_iter) on every candidate, kmcodec.cc:157-160).
// If the CHOSEN candidate's metric decode reached a zero syndrome, the final Decoder(max_iter) call on the same input
// (kmcodec.cc:70-71) repeats exactly those iterations and stops at the same one with the same decisions and the same
// return value (metric_iter <= max_iter): its result is taken from the metric decode, and only the other frames are
// queued for the final decoder.  One warp per frame.
__global__ void reuse_metric_kernel(int B, int nbits, int bit_offset, int words_n, const int32_t *kstar, const float *metric,
                                    const int32_t *mret, const uint32_t *cand_bits, uint32_t *uu_hat, int32_t *ret,
                                    int32_t *queue, int32_t *queue_n) {
  const int lane = threadIdx.x & 31, words = (nbits + 31) / 32;
  for (int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); f < B; f += gridDim.x * (blockDim.x >> 5)) {
    const int k = kstar[f];
    if (metric[(size_t)f * 4 + k] != 0.0f) {  // not converged: the final decoder starts from scratch, like the reference
      if (lane == 0) queue[atomicAdd(queue_n, 1)] = f;
      continue;
    }
    const uint32_t *src = cand_bits + ((size_t)f * 4 + k) * words_n;
    for (int w = lane; w < words; w += 32) {
      const int q = bit_offset + 32 * w, wi = q >> 5, sh = q & 31;
      uint32_t v = src[wi] >> sh;
      if (sh && wi + 1 < words_n) v |= src[wi + 1] << (32 - sh);
      const int valid = nbits - 32 * w;
      if (valid < 32) v &= (1u << valid) - 1u;
      uu_hat[(size_t)f * words + w] = v;
    }
    if (lane == 0) ret[f] = mret[(size_t)f * 4 + k];
  }
}

// extract_bits_kernel for the frames of a device-side list
__global__ void extract_bits_queue_kernel(const int32_t *queue, const int32_t *queue_n, int nbits, int bit_offset, int src_words,
                                          const uint32_t *src, uint32_t *dst) {
  const int words = (nbits + 31) / 32;
  const long long total = (long long)(*queue_n) * words;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = queue[i / words], w = (int)(i % words);
    const uint32_t *s = src + (size_t)f * src_words;
    const int q = bit_offset + 32 * w, wi = q >> 5, sh = q & 31;
    uint32_t v = s[wi] >> sh;
    if (sh && wi + 1 < src_words) v |= s[wi + 1] << (32 - sh);
    const int valid = nbits - 32 * w;
    if (valid < 32) v &= (1u << valid) - 1u;
    dst[(size_t)f * words + w] = v;
  }
}

// ------------------------------------------------------------------------------------------------ utilities
__global__ void pack_bits_kernel(int F, int nbits, const int32_t *bits, uint32_t *packed) {
  const int words = (nbits + 31) / 32;
  const long long total = (long long)F * words * 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long wi = i >> 5;
    const int f = (int)(wi / words), w = (int)(wi % words), t = w * 32 + (int)(i & 31);
    const int bit = (t < nbits) ? (bits[(size_t)f * nbits + t] != 0) : 0;
    const uint32_t word = __ballot_sync(0xffffffffu, bit);
    if ((i & 31) == 0) packed[wi] = word;
  }
}

__global__ void unpack_bits_kernel(int F, int nbits, int bit_offset, int src_words, const uint32_t *packed, int32_t *bits) {
  const long long total = (long long)F * nbits;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / nbits), t = (int)(i % nbits) + bit_offset;
    bits[i] = (int32_t)((packed[(size_t)f * src_words + (t >> 5)] >> (t & 31)) & 1u);
  }
}

__global__ void extract_bits_kernel(int F, int nbits, int bit_offset, int src_words, const uint32_t *src, uint32_t *dst) {
  const int words = (nbits + 31) / 32;
  const long long total = (long long)F * words;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / words), w = (int)(i % words);
    const uint32_t *s = src + (size_t)f * src_words;
    const int q = bit_offset + 32 * w, wi = q >> 5, sh = q & 31;
    uint32_t v = s[wi] >> sh;
    if (sh && wi + 1 < src_words) v |= s[wi + 1] << (32 - sh);
    const int valid = nbits - 32 * w;
    if (valid < 32) v &= (1u << valid) - 1u;
    dst[i] = v;
  }
}

__global__ void lr_to_llr_kernel(size_t n, const float *lr, float *llr) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    llr[i] = logf(lr[i]);
}

// SourceSink::CntErr (lib/lab/src/sourcesink.cc:29-47), 64-bit counters; one warp per frame.
__global__ void count_errors_kernel(int B, int k, int k_words, const uint32_t *u, const uint32_t *uh, const int32_t *ret,
                                    int max_iter, unsigned long long *counters) {
  __shared__ unsigned long long s_acc[3];
  if (threadIdx.x < 3) s_acc[threadIdx.x] = 0ull;
  __syncthreads();
  const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  unsigned long long err_blk = 0, err_bit = 0, iters = 0;
  for (int f = blockIdx.x * warps + (threadIdx.x >> 5); f < B; f += gridDim.x * warps) {
    int ne = 0;
    for (int w = lane; w < k_words; w += 32) ne += __popc(u[(size_t)f * k_words + w] ^ uh[(size_t)f * k_words + w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ne += __shfl_xor_sync(0xffffffffu, ne, o);
    if (lane == 0) {
      err_bit += ne;
      err_blk += ne > 0;
      if (ret) iters += min(ret[f], max_iter);
    }
  }
  if (lane == 0) {
    atomicAdd(&s_acc[0], err_blk);
    atomicAdd(&s_acc[1], err_bit);
    atomicAdd(&s_acc[2], iters);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      atomicAdd(counters + 0, (unsigned long long)B);
      atomicAdd(counters + 2, (unsigned long long)B * (unsigned long long)k);
    }
    if (s_acc[0]) atomicAdd(counters + 1, s_acc[0]);
    if (s_acc[1]) atomicAdd(counters + 3, s_acc[1]);
    if (s_acc[2]) atomicAdd(counters + 4, s_acc[2]);
  }
}

inline int grid_for(long long total, int threads, int cap = 148 * 16) {
  long long g = (total + threads - 1) / threads;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

cudaError_t launch_gen_bits(const GenParams &g, uint32_t *u_packed, cudaStream_t s) {
  const long long total = (long long)g.B * ((g.k_words + 3) / 4);
  gen_bits_kernel<<<grid_for(total, 256), 256, 0, s>>>(g, u_packed);
  return cudaGetLastError();
}

static cudaError_t launch_encode_impl(const GenParams &g, uint32_t *u_packed, uint32_t *c_packed, bool gen, cudaStream_t s) {
  const int chk_words = (g.n_chk + 31) / 32;
  const int smem = ((((ENC_FT * (g.k_words + 1) + ENC_FT * (chk_words + 1)) + 3) & ~3) + ENC_FT * g.k_words) * (int)sizeof(uint32_t);
  if (smem > 48 * 1024) {
    cudaError_t e = gen ? cudaFuncSetAttribute(encode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                        : cudaFuncSetAttribute(encode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  if (g.B < 1) return cudaSuccess;
  if (gen) encode_kernel<true><<<(g.B + ENC_FT - 1) / ENC_FT, ENC_THREADS, smem, s>>>(g, u_packed, c_packed);
  else encode_kernel<false><<<(g.B + ENC_FT - 1) / ENC_FT, ENC_THREADS, smem, s>>>(g, u_packed, c_packed);
  return cudaGetLastError();
}
cudaError_t launch_encode(const GenParams &g, const uint32_t *u_packed, uint32_t *c_packed, cudaStream_t s) {
  return launch_encode_impl(g, const_cast<uint32_t *>(u_packed), c_packed, false, s);
}
cudaError_t launch_gen_encode(const GenParams &g, uint32_t *u_packed, uint32_t *c_packed, cudaStream_t s) {
  return launch_encode_impl(g, u_packed, c_packed, true, s);
}

cudaError_t launch_channel(const GenParams &g, const uint32_t *c_packed, const float2 *h_in, const float2 *noise,
                           float2 *h_out, float2 *y, cudaStream_t s) {
  if (noise) {
    channel_replay_kernel<<<grid_for((long long)g.B * g.n_sym, 256), 256, 0, s>>>(g, c_packed, h_in, noise, y);
    return cudaGetLastError();
  }
  if (!h_out) return cudaErrorInvalidValue;  // the Philox channel always reports its fades
  channel_kernel<<<g.B < 148 * 10 ? g.B : 148 * 10, CH_THREADS, 0, s>>>(g, c_packed, h_out, y);
  return cudaGetLastError();
}

// the reference decoder's own input — double P(bit = 0) (binaryldpccodec.cc:165) — to the kernels' likelihood ratio
__global__ void p0_to_lr_kernel(size_t n, const double *p0, float *lr) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double p = p0[i], r = p / (1.0 - p);
    lr[i] = (float)fmin(fmax(r, 1.0e-12), 1.0e12);
  }
}
cudaError_t launch_p0_to_lr(size_t n, const double *p0, float *lr, cudaStream_t s) {
  if (n < 1) return cudaSuccess;
  p0_to_lr_kernel<<<grid_for((long long)n, 256), 256, 0, s>>>(n, p0, lr);
  return cudaGetLastError();
}

__global__ void f64_to_f32_kernel(size_t n, const double *in, float *out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (float)in[i];
}
cudaError_t launch_f64_to_f32(size_t n, const double *in, float *out, cudaStream_t s) {
  if (n < 1) return cudaSuccess;
  f64_to_f32_kernel<<<grid_for((long long)n, 256), 256, 0, s>>>(n, in, out);
  return cudaGetLastError();
}

template <bool F64IN>
static cudaError_t launch_kmeans_warp(int B, const void *y, int n_sym, const KmConst &kc, int iters, float2 *hhat,
                                      double2 *hhat64, int32_t *passes, float2 *y32_out, int num_sms, cudaStream_t s) {
  const int spl = (n_sym + 31) / 32;
  const int grid = std::min((B + KMW_WARPS - 1) / KMW_WARPS, num_sms * 16);
#define KMW(SPL, NB)                                                                                                          \
  do {                                                                                                                        \
    const int smem = F64IN ? 0 : KMW_WARPS * SPL * 32 * (int)sizeof(float2);                                                  \
    if (smem > 40 * 1024) cudaFuncSetAttribute(kmeans_warp_kernel<SPL, NB, F64IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    kmeans_warp_kernel<SPL, NB, F64IN><<<grid, KMW_WARPS * 32, smem, s>>>(B, y, n_sym, kc, iters, hhat, hhat64, passes, y32_out);      \
  } while (0)
  if (kc.n_nb <= 2) {
    if (spl <= 16) KMW(16, 2); else if (spl <= 24) KMW(24, 2); else if (spl <= 36) KMW(36, 2); else if (spl <= 48) KMW(48, 2); else KMW(64, 2);
  } else if (kc.n_nb <= 4) {
    if (spl <= 16) KMW(16, 4); else if (spl <= 24) KMW(24, 4); else if (spl <= 36) KMW(36, 4); else if (spl <= 48) KMW(48, 4); else KMW(64, 4);
  } else {
    if (spl <= 16) KMW(16, 8); else if (spl <= 24) KMW(24, 8); else if (spl <= 36) KMW(36, 8); else if (spl <= 48) KMW(48, 8); else KMW(64, 8);
  }
#undef KMW
  return cudaGetLastError();
}

cudaError_t launch_kmeans(int B, const void *y, int y_is_f64, int n_sym, const float2 *points, int q, const KmConst &kc,
                          int iters, float2 *hhat, double2 *hhat64, int32_t *passes, float2 *y32_out,
                          int num_sms, cudaStream_t s) {
  if (B < 1) return cudaSuccess;
  if (kc.n_nb >= 1 && kc.n_nb <= 8 && n_sym <= 32 * 64) {  // warp per frame, Voronoi-neighbour half-plane tests
    return y_is_f64 ? launch_kmeans_warp<true>(B, y, n_sym, kc, iters, hhat, hhat64, passes, y32_out, num_sms, s)
                    : launch_kmeans_warp<false>(B, y, n_sym, kc, iters, hhat, hhat64, passes, nullptr, num_sms, s);
  }
  // general fallback (constellations whose first point has more than 8 Voronoi neighbours, very long frames): one CTA
  // per frame, full distance comparison in fp32 against every constellation point
  const float2 *y32 = reinterpret_cast<const float2 *>(y);
  if (y_is_f64) {
    if (!y32_out) return cudaErrorInvalidValue;
    cudaError_t e = launch_f64_to_f32((size_t)B * n_sym * 2, reinterpret_cast<const double *>(y), reinterpret_cast<float *>(y32_out), s);
    if (e != cudaSuccess) return e;
    y32 = y32_out;
  }
  const int spt = (n_sym + KM_THREADS - 1) / KM_THREADS;
  const int grid = B < num_sms * 8 ? B : num_sms * 8;
  const int smem = q * (int)sizeof(float2);
  if (spt <= 4) kmeans_kernel<4><<<grid, KM_THREADS, smem, s>>>(B, y32, n_sym, points, q, iters, hhat, hhat64, passes);
  else if (spt <= 9) kmeans_kernel<9><<<grid, KM_THREADS, smem, s>>>(B, y32, n_sym, points, q, iters, hhat, hhat64, passes);
  else if (spt <= 16) kmeans_kernel<16><<<grid, KM_THREADS, smem, s>>>(B, y32, n_sym, points, q, iters, hhat, hhat64, passes);
  else if (spt <= 64) kmeans_kernel<64><<<grid, KM_THREADS, smem, s>>>(B, y32, n_sym, points, q, iters, hhat, hhat64, passes);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

template <int BITS>
static cudaError_t launch_demap_bits(const DemapParams &d, int grid, int smem, cudaStream_t s) {
  if constexpr (BITS == 2) {  // 4 points: the compiled partition codes (demap_symbol_q4); anything else takes the general path
    if (d.n_cand == 4 && d.symmetric && d.q4_code == kQ4CodeQpsk) {
      demap_kernel<2, 4, false, kQ4CodeQpsk><<<grid, DM_THREADS, smem, s>>>(d);
      return cudaGetLastError();
    }
    if (d.n_cand == 4 && d.symmetric && d.q4_code == kQ4Code4psk) {
      demap_kernel<2, 4, false, kQ4Code4psk><<<grid, DM_THREADS, smem, s>>>(d);
      return cudaGetLastError();
    }
    if (d.n_cand == 1 && d.q4_code != 0xFFFFFFFFu) {  // (candidate 0 reads 01|23 and 02|13 whatever the file)
      demap_kernel<2, 1, false, kQ4CodeQpsk><<<grid, DM_THREADS, smem, s>>>(d);
      return cudaGetLastError();
    }
  }
  if (d.n_cand == 4) {
    constexpr bool kCanSym = BITS <= 5;  // (64 points: the quad-of-lanes path keeps one softmax per candidate, see there)
    const bool sym = kCanSym && d.symmetric;
    if (sym) smem = ((smem + 15) & ~15) + (int)sizeof(float) * (1 << BITS) * DM_THREADS;  // one column of Q words per thread
    auto k = sym ? demap_kernel<BITS, 4, kCanSym> : demap_kernel<BITS, 4, false>;
    if (smem > 48 * 1024) {  // per device: set on every large launch
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return e;
    }
    k<<<grid, DM_THREADS, smem, s>>>(d);
  } else if (d.n_cand == 1) {
    demap_kernel<BITS, 1, false><<<grid, DM_THREADS, smem, s>>>(d);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_demap(const DemapParams &d, int num_sms, cudaStream_t s) {
  const int grid = d.B < num_sms * 8 ? d.B : num_sms * 8;
  if (grid < 1) return cudaSuccess;
  const int smem = (int)sizeof(float2) * (d.q == 64 ? 68 : d.q) * 4 + (d.winner_only ? (int)sizeof(float) * 4 * d.n_tx : 0) + d.n_tx + d.punct + 32;
  switch (d.bits_per_symbol) {
    case 1: return launch_demap_bits<1>(d, grid, smem, s);
    case 2: return launch_demap_bits<2>(d, grid, smem, s);
    case 3: return launch_demap_bits<3>(d, grid, smem, s);
    case 4: return launch_demap_bits<4>(d, grid, smem, s);
    case 5: return launch_demap_bits<5>(d, grid, smem, s);
    case 6: return launch_demap_bits<6>(d, grid, smem, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_syndrome_weight(int F, const uint32_t *bits, int words_n, int m_rows, const int32_t *row_ptr,
                                   const int32_t *col_idx, float *metric, cudaStream_t s) {
  if (F < 1) return cudaSuccess;
  syndrome_weight_kernel<<<F < 148 * 8 ? F : 148 * 8, 256, 0, s>>>(F, bits, words_n, m_rows, row_ptr, col_idx, metric);
  return cudaGetLastError();
}

cudaError_t launch_soft_chain(const SoftChainParams &p, int round, cudaStream_t s) {
  soft_chain_kernel<<<1, 1024, 0, s>>>(p, round);
  return cudaGetLastError();
}

cudaError_t launch_abs_inplace(int n, float *v, cudaStream_t s) {
  if (n < 1) return cudaSuccess;
  abs_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, v);
  return cudaGetLastError();
}

cudaError_t launch_argmin4(int B, const float *metric, int32_t *kstar, cudaStream_t s) {
  if (B < 1) return cudaSuccess;
  argmin4_kernel<<<(B + 255) / 256, 256, 0, s>>>(B, metric, kstar);
  return cudaGetLastError();
}

cudaError_t launch_reuse_metric(int B, int nbits, int bit_offset, int words_n, const int32_t *kstar, const float *metric,
                                const int32_t *mret, const uint32_t *cand_bits, uint32_t *uu_hat, int32_t *ret, int32_t *queue,
                                int32_t *queue_n, cudaStream_t s) {
  if (B < 1) return cudaSuccess;
  reuse_metric_kernel<<<std::min((B + 7) / 8, 148 * 8), 256, 0, s>>>(B, nbits, bit_offset, words_n, kstar, metric, mret, cand_bits,
                                                                  uu_hat, ret, queue, queue_n);
  return cudaGetLastError();
}

cudaError_t launch_extract_bits_queue(int max_frames, const int32_t *queue, const int32_t *queue_n, int nbits, int bit_offset,
                                      int src_words, const uint32_t *src, uint32_t *dst, cudaStream_t s) {
  const long long total = (long long)max_frames * ((nbits + 31) / 32);
  if (total < 1) return cudaSuccess;
  extract_bits_queue_kernel<<<grid_for(total, 256), 256, 0, s>>>(queue, queue_n, nbits, bit_offset, src_words, src, dst);
  return cudaGetLastError();
}

cudaError_t launch_pack_bits(int F, int nbits, const int32_t *bits, uint32_t *packed, cudaStream_t s) {
  const long long total = (long long)F * ((nbits + 31) / 32) * 32;
  if (total < 1) return cudaSuccess;
  pack_bits_kernel<<<grid_for(total, 256), 256, 0, s>>>(F, nbits, bits, packed);
  return cudaGetLastError();
}

cudaError_t launch_unpack_bits(int F, int nbits, int bit_offset, int src_words, const uint32_t *packed, int32_t *bits,
                               cudaStream_t s) {
  const long long total = (long long)F * nbits;
  if (total < 1) return cudaSuccess;
  unpack_bits_kernel<<<grid_for(total, 256), 256, 0, s>>>(F, nbits, bit_offset, src_words, packed, bits);
  return cudaGetLastError();
}

cudaError_t launch_extract_bits(int F, int nbits, int bit_offset, int src_words, const uint32_t *src, uint32_t *dst,
                                cudaStream_t s) {
  const long long total = (long long)F * ((nbits + 31) / 32);
  if (total < 1) return cudaSuccess;
  extract_bits_kernel<<<grid_for(total, 256), 256, 0, s>>>(F, nbits, bit_offset, src_words, src, dst);
  return cudaGetLastError();
}

cudaError_t launch_lr_to_llr(size_t n, const float *lr, float *llr, cudaStream_t s) {
  if (n < 1) return cudaSuccess;
  lr_to_llr_kernel<<<grid_for((long long)n, 256), 256, 0, s>>>(n, lr, llr);
  return cudaGetLastError();
}

cudaError_t launch_count_errors(int B, int k, int k_words, const uint32_t *u_packed, const uint32_t *uu_hat_packed,
                                const int32_t *ret, int max_iter, unsigned long long *counters, cudaStream_t s) {
  if (B < 1) return cudaSuccess;
  const int warps = 8;
  int grid = (B + warps - 1) / warps;
  if (grid > 148 * 4) grid = 148 * 4;
  count_errors_kernel<<<grid, warps * 32, 0, s>>>(B, k, k_words, u_packed, uu_hat_packed, ret, max_iter, counters);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ roofline probe
// What the SMs of THIS device sustain for conflict-free shared-memory loads (LDS.128, every lane its own 16 bytes, no
// two loads of the loop at one address): the denominator of the decoder's roofline.frac, measured instead of derived
// from 128 B/clk/SM x clock (tools/microbench.cu is the stand-alone version with the other pipes).
namespace {
constexpr int PROBE_T = 1024, PROBE_ITERS = 2048;  // ~2 ms per launch: the launch ramp is < 1 % of it
__global__ void __launch_bounds__(PROBE_T, 2) smem_probe_kernel(float *out) {
  extern __shared__ __align__(16) float psm[];
  for (int i = threadIdx.x; i < 8192; i += PROBE_T) psm[i] = (float)i;
  __syncthreads();
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const int base = (threadIdx.x * 4) & 8191;
  for (int it = 0; it < PROBE_ITERS; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int a = (base + (it * 8 + u) * 1056) & (8191 & ~3);
      float v, w, x, y;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v), "=f"(w), "=f"(x), "=f"(y)
                   : "r"((unsigned)__cvta_generic_to_shared(psm + a)));
      a0 += v; a1 += w; a2 += x; a3 += y;
    }
  }
  out[blockIdx.x * PROBE_T + threadIdx.x] = a0 + a1 + a2 + a3;
}
}  // namespace

cudaError_t measure_smem_bandwidth(int num_sms, double *gbs, cudaStream_t s) {
  const int grid = num_sms * 2;
  float *out = nullptr;
  cudaError_t e = cudaMalloc(&out, sizeof(float) * grid * PROBE_T);
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3 && e == cudaSuccess; r++) {  // first launch = warm-up
    cudaEventRecord(e0, s);
    smem_probe_kernel<<<grid, PROBE_T, 32768, s>>>(out);
    cudaEventRecord(e1, s);
    e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) *gbs = (double)grid * PROBE_T * PROBE_ITERS * 8 * 16 / (best * 1e-3) * 1e-9;
  return e;
}

}  // namespace kml
