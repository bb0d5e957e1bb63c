// Layered (row-serial) min-sum decoder for quasi-cyclic codes — THROUGHPUT mode (`algorithm = 3`), not reference-pinned.
//
// BASELINE config 2 names "5G LDPC BG2 … min-sum decoding"; the reference has only the flooding sum-product,
// so like bp_minsum.cu this kernel has no reference output to be compared with bit for bit: it is gated by BER/FER against
// the sum-product decoder on the same frames (tests/test_gpu_minsum.py).  What it keeps from the reference: the clip of the
// messages (±27.63), the punctured columns' neutral prior (binary5gldpccodec.cc:126-130), the return value convention
// iter + (iter < max_iter), and "stop when the decisions satisfy every check".
//
// Schedule.  A quasi-cyclic parity-check matrix is a grid of Z x Z blocks, each zero or a cyclically shifted identity
// (5G BG2 R1/2 K960: 12 block rows x 22 block columns, Z = 96).  The Z checks of one block row touch Z DIFFERENT variables
// in every block column, so a block row is a conflict-free LAYER: thread z of a frame's group owns check z of every layer
// and walks the layers in order, each check seeing the posteriors as the layers before it left them — which is why a
// layered decoder needs about half the iterations of a flooding one.  Per check of degree d:
//     v_k   = L[col_k] - c2v_old_k                     (L = posterior LLR, in shared memory; d loads)
//     c2v_k = sign * max(alpha * min_{k' != k} |v_k'| - beta, 0)   clipped to +-27.63
//     L[col_k] = v_k + c2v_k                           (d stores)
// The old messages of a check are kept COMPRESSED — (min1, min2) as one half2 word and (index of the minimum, sign bits)
// as another — two words per check and layer, thread-private in shared memory.  Per edge-iteration: one LDS + one STS of
// 4 bytes (plus a 16-bit table read for the variable's index), against four 4-byte accesses for the flooding kernels, and
// no variable-node phase at all.  What it costs instead is compare / select work on the half-rate ALU pipe (measured:
// profiles/r2s_layered_5G_full_B8192.txt), which is why it does not beat the flooding plan kernel per iteration.
//
// One CTA holds FPC frames (groups of Z threads, each with its own named barrier and its own place in the frame queue).
#include <cuda_fp16.h>

#include "kml_internal.h"
#include "kml_kernels.cuh"
#include "bp_minsum_nodes.cuh"

namespace kml {
namespace {

constexpr int LAY_DC = 10;   // maximum check degree handled (BG2: 10); kml_api checks it
constexpr int LAY_FPC = 4;   // frames per CTA
constexpr int LAY_MAX_EDGES = 256, LAY_MAX_LAYERS = 32;
constexpr int LAY_MAX_THREADS = LAY_FPC * 128;  // lifting sizes up to 128 (four warps per frame group)

__device__ __forceinline__ void group_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// One check of degree D: thread-private compressed state `st`, posteriors L.  Returns 1 if the check saw unsatisfied parity
// or moved a decision.
template <int D>
__device__ __forceinline__ int layer_check(float *L, uint32_t *st, int G, const uint16_t *tab_l, float alpha, float beta) {
  const __half2 mm = *reinterpret_cast<const __half2 *>(st);  // (min1, min2) of the last visit, already scaled
  const uint32_t meta = st[G];                                   // bits 0-15: sign of each old message, 16-19: index of min1
  const float om1 = __low2float(mm), om2 = __high2float(mm);
  const int oidx = (int)(meta >> 16);
  float v[D];
  int col[D];
  float m1 = 3.0e38f, m2 = 3.0e38f;
  int idx = 0;
  uint32_t sgn = 0, par = 0, hard = 0, dec = 0;
#pragma unroll
  for (int k = 0; k < D; k++) {
    col[k] = tab_l[k * G];  // this thread's variable in block k of the layer (CTA-wide table, see the kernel)
    const float lv = L[col[k]];
    const uint32_t bit = lv > 0.0f ? 0u : 1u;  // the decision this check sees (tie → 1, like the other decoders)
    hard ^= bit;
    dec |= bit << k;
    const float old = ((meta >> k) & 1u) ? -(k == oidx ? om2 : om1) : (k == oidx ? om2 : om1);
    v[k] = lv - old;
    const float a = fabsf(v[k]);
    const uint32_t neg = v[k] < 0.0f ? 1u : 0u;
    sgn |= neg << k;
    par ^= neg;
    if (a < m1) { m2 = m1; m1 = a; idx = k; }
    else if (a < m2) m2 = a;
  }
  const float s1 = fminf(fmaxf(fmaf(alpha, m1, -beta), 0.0f), msn::kLlrClip);
  const float s2 = fminf(fmaxf(fmaf(alpha, m2, -beta), 0.0f), msn::kLlrClip);
  // stored rounded to half: the next visit subtracts exactly what this one adds
  const __half2 nm = __floats2half2_rn(s1, s2);
  const float q1 = __low2float(nm), q2 = __high2float(nm);
  uint32_t nsgn = 0, moved = 0;
#pragma unroll
  for (int k = 0; k < D; k++) {
    const uint32_t neg = par ^ ((sgn >> k) & 1u);  // sign of the product of the OTHER messages
    const float mag = k == idx ? q2 : q1;
    nsgn |= neg << k;
    const float nl = v[k] + (neg ? -mag : mag);
    moved |= ((nl > 0.0f ? 0u : 1u) ^ (dec >> k)) & 1u;  // a decision changed during this iteration
    L[col[k]] = nl;
  }
  *reinterpret_cast<__half2 *>(st) = nm;
  st[G] = nsgn | ((uint32_t)idx << 16);
  return (int)(hard | moved);
}

// (3 CTAs per SM at 56 registers: measured no faster — the kernel is issue-bound)
__global__ void __launch_bounds__(LAY_MAX_THREADS, 2) ms_layered_kernel(const DecParams p) {
  extern __shared__ __align__(16) unsigned char lsm[];
  const LayeredTables &lt = p.lay;
  __shared__ uint32_t s_cs[LAY_MAX_EDGES];  // the block structure: 78 words for BG2
  __shared__ int s_ptr[LAY_MAX_LAYERS + 1];
  for (int i = threadIdx.x; i < lt.n_edges; i += blockDim.x) s_cs[i] = __ldg(lt.lay_cs + i);
  for (int i = threadIdx.x; i <= lt.n_layers; i += blockDim.x) s_ptr[i] = __ldg(lt.lay_ptr + i);
  __syncthreads();
  const int Z = lt.z, G = blockDim.x / LAY_FPC;  // G = threads per frame group (Z rounded up to a warp multiple)
  // tab[e][z] = the variable check z meets in block e: (block column) * Z + (z + shift) mod Z — the same for every frame, so it
  // is expanded once per CTA (one LDS.U16 per edge instead of seven integer instructions)
  uint16_t *tab = reinterpret_cast<uint16_t *>(lsm);
  for (int i = threadIdx.x; i < lt.n_edges * G; i += blockDim.x) {
    const uint32_t cs = s_cs[i / G];
    int zz = i % G + (int)(cs & 0xFFFFu);
    zz = zz >= Z ? zz - Z : zz;
    tab[i] = (uint16_t)((cs >> 16) + (zz < Z ? zz : 0));
  }
  __syncthreads();
  const int tab_words = (lt.n_edges * G + 1) / 2;
  const int grp = threadIdx.x / G, z = threadIdx.x % G, lane = threadIdx.x & 31;
  const int n = p.t.n, NL = lt.n_layers;
  // per group: posteriors L[n], then the compressed messages [NL][2][G]
  const int grp_words = ((n + 3) & ~3) + NL * G * 2;
  float *L = reinterpret_cast<float *>(lsm) + ((tab_words + 3) & ~3) + (size_t)grp * grp_words;
  uint32_t *cm = reinterpret_cast<uint32_t *>(L + ((n + 3) & ~3));
  __shared__ int s_frame[LAY_FPC];
  const bool active = z < Z;
  const int bar_id = 1 + grp;  // (barrier 0 is __syncthreads)

  while (true) {
    if (z == 0) s_frame[grp] = next_frame(p);
    group_barrier(bar_id, G);
    const int f = s_frame[grp];
    if (f < 0) break;
    const float *in = p.in + (size_t)(p.sel ? f * p.n_cand + __ldg(p.sel + f) : f) * p.t.n_tx;
    for (int v = z; v < n; v += G)  // punctured variables: prior (0.5, 0.5) → LLR 0
      L[v] = v < p.t.punct ? 0.0f : msn::load_channel_llr(in, v - p.t.punct, p.in_is_lr);
    for (int i = z; i < NL * G * 2; i += G) cm[i] = 0u;  // all messages 0
    group_barrier(bar_id, G);

    int ret = p.iters + (p.iters < p.max_iter);
    bool latched = false;
    for (int t = 0; t < p.iters; t++) {
      int fail = 0;
      for (int l = 0; l < NL; l++) {
        if (active) {
          const int e0 = s_ptr[l], d = s_ptr[l + 1] - e0;
          uint32_t *st = cm + (l * 2) * G + z;  // two planes per layer: conflict-free for the group's lanes
          const uint16_t *tab_l = tab + e0 * G + z;
          switch (d) {  // (uniform per layer; compiled per degree so that short rows do not issue ten predicated edges)
            case 3: fail |= layer_check<3>(L, st, G, tab_l, p.alpha, p.beta); break;
            case 4: fail |= layer_check<4>(L, st, G, tab_l, p.alpha, p.beta); break;
            case 5: fail |= layer_check<5>(L, st, G, tab_l, p.alpha, p.beta); break;
            case 6: fail |= layer_check<6>(L, st, G, tab_l, p.alpha, p.beta); break;
            case 7: fail |= layer_check<7>(L, st, G, tab_l, p.alpha, p.beta); break;
            case 8: fail |= layer_check<8>(L, st, G, tab_l, p.alpha, p.beta); break;
            case 9: fail |= layer_check<9>(L, st, G, tab_l, p.alpha, p.beta); break;
            case 10: fail |= layer_check<10>(L, st, G, tab_l, p.alpha, p.beta); break;
            case 2: fail |= layer_check<2>(L, st, G, tab_l, p.alpha, p.beta); break;
            default: fail |= layer_check<1>(L, st, G, tab_l, p.alpha, p.beta); break;
          }
        }
        group_barrier(bar_id, G);  // the next layer's checks read these posteriors
      }
      // every check saw satisfied parity and no decision moved afterwards → the decisions are a codeword
      int any = fail;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) any |= __shfl_xor_sync(0xffffffffu, any, o);
      __shared__ int s_fail[LAY_FPC][2];
      if (z == 0) s_fail[grp][t & 1] = 0;
      group_barrier(bar_id, G);
      if (lane == 0 && any) s_fail[grp][t & 1] = 1;
      group_barrier(bar_id, G);
      if (!s_fail[grp][t & 1] && !latched) {
        latched = true;
        ret = t + 1 + (t + 1 < p.max_iter);  // t + 1 iterations were executed (the flooding decoders test BEFORE the check phase)
        if (p.early_exit) break;
        for (int c0 = 0; c0 < n; c0 += G) {  // fixed-iteration mode: latch the decisions now
          const int vv = c0 + z;
          const uint32_t word = __ballot_sync(0xffffffffu, vv < n && !(L[vv] > 0.0f));
          if (lane == 0 && vv < n) p.out_bits[(size_t)f * p.words_n + (vv >> 5)] = word;
        }
      }
    }
    if (!latched || p.early_exit) {
      for (int c0 = 0; c0 < n; c0 += G) {
        const int vv = c0 + z;
        const uint32_t word = __ballot_sync(0xffffffffu, vv < n && !(L[vv] > 0.0f));
        if (lane == 0 && vv < n) p.out_bits[(size_t)f * p.words_n + (vv >> 5)] = word;
      }
    }
    if (z == 0) p.out_ret[f] = ret;
    group_barrier(bar_id, G);  // s_frame / L are reused
  }
}

}  // namespace

int layered_threads(int z) { return LAY_FPC * ((z + 31) & ~31); }
int layered_smem_bytes(int n, int n_layers, int n_edges, int z) {
  const int G = (z + 31) & ~31;
  const int tab_words = (n_edges * G + 1) / 2;
  return (((tab_words + 3) & ~3) + LAY_FPC * (((n + 3) & ~3) + n_layers * G * 2)) * 4;
}
int layered_max_degree() { return LAY_DC; }
int layered_max_threads() { return LAY_MAX_THREADS; }
int layered_max_edges() { return LAY_MAX_EDGES; }
int layered_max_layers() { return LAY_MAX_LAYERS; }
int layered_frames_per_cta() { return LAY_FPC; }
dec_kernel_t layered_kernel() { return ms_layered_kernel; }

}  // namespace kml
