// Kernel parameter blocks and launchers of libkmldpc_b200.so (sm_100a only).
#ifndef KML_KERNELS_CUH
#define KML_KERNELS_CUH
#include <cuda_runtime.h>

#include <cstdint>

namespace kml {

// ---- belief-propagation decoder -------------------------------------------------------------------------------
// Edge messages live in shared memory as one 32-bit word per edge at word address k * plane + slot(row), k = position
// of the edge inside its row: check-node threads (one row slot each) touch consecutive words → conflict free, and
// variable-node threads gather/scatter through per-variable address lists that layout_opt.cpp makes conflict free.
// (The regular sum-product kernels use a second, row-major table: 6 * slot + k, see bp_decode.cu.)
struct DecTables {
  const uint16_t *vn_addr;  // [n][dv_max] shared-memory word address of each edge of a variable, 0xFFFF = none
  const uint8_t *vn_deg;    // [n]
  const uint8_t *cn_deg;    // [m_pad] degree of the row held by a slot (0 = padding)
  int n, m_pad, plane, n_tx, punct, dv_max, dc_max;  // m_pad = row slots (multiple of 32), plane = m_pad + 1
  // generic sum-product kernel: per-item address lists and degree-balanced warp work lists
  const uint16_t *vn_addr_g;  // [n_pad / 32][dv_max][32]
  const uint32_t *vn_items;   // [vn_items_n] run of groups of 32 variables: first group | degree << 16 (0xFF = mixed) |
                              // number of groups << 24, laid out [round][warp]; 0xFFFFFFFF = none
  const uint32_t *cn_items;   // [cn_items_n] groups of 32 row slots, same layout
  int n_pad, vn_items_n, cn_items_n;
};

// Block structure of a quasi-cyclic code for the layered decoder (bp_layered.cu): layer l = block row l; its checks'
// edges e in [lay_ptr[l], lay_ptr[l + 1]) connect check z to variable (lay_cs[e] >> 16) + (z + (lay_cs[e] & 0xFFFF)) mod z
// (high half = block column * Z, low half = cyclic shift).
struct LayeredTables {
  const int32_t *lay_ptr;   // [n_layers + 1]
  const uint32_t *lay_cs;   // [n_edges]
  int z, n_layers, n_edges;
};

struct DecParams {
  DecTables t;
  LayeredTables lay;        // algorithm = 3 only
  const float *in;          // [B * n_cand][n_tx] natural-log LLR, or likelihood ratio P0/P1 when in_is_lr
  const int32_t *sel;       // optional [B]: candidate picked per frame (row f * n_cand + sel[f] of `in`)
  int n_cand, in_is_lr;
  int B, iters, max_iter, early_exit;
  uint32_t *out_bits;       // [B][words_n] hard decisions of all graph variables, bit-packed
  int32_t *out_ret;         // [B] reference return value: iter + (iter < max_iter)
  double *out_soft;         // optional [B]: sum over rows of ln(syndrom_soft) after the LAST check-node phase executed
                            // (accumulated with atomics: zero it first; untouched for a frame that leaves at t = 0)
  float *out_synd;          // optional [B]: unsatisfied checks of the final decisions (= ParityCheck(cc_hat)); only the
                            // kernels dec_has_synd_output() names fill it
  unsigned int *work_counter;  // dynamic frame scheduler (zeroed by the launcher)
  int words_n;
  float alpha, beta;        // min-sum: magnitude = max(alpha m - beta, 0) (algorithm = 1 | 2)
  // optional frame queue: entry i of the queue decodes frame frame_idx[i]; the queue length is *n_frames_dev (written by
  // an earlier kernel of the stream; B is then only the upper bound the grid is sized for).  Inputs and outputs stay
  // indexed by the frame number, so a queue is just a subset / an order of the batch.
  // queue_cap > 0: a TWO-ENDED queue in one array of queue_cap entries — n_frames_dev[0] entries filled from the front,
  // n_frames_dev[1] from the back (entry j of the back part at frame_idx[queue_cap - 1 - j]); the front part is served
  // first.  The demapper puts the frames that will probably run the full iteration count in front (longest first).
  const int32_t *frame_idx;
  const int32_t *n_frames_dev;
  int queue_cap;
};

#ifdef __CUDACC__
__device__ __forceinline__ int frame_count(const DecParams &p) {
  if (!p.n_frames_dev) return p.B;
  const int n = p.queue_cap ? __ldg(p.n_frames_dev) + __ldg(p.n_frames_dev + 1) : __ldg(p.n_frames_dev);
  return min(n, p.B);
}
// frame behind entry i of the queue (i < frame_count)
__device__ __forceinline__ int frame_at(const DecParams &p, int i) {
  if (!p.frame_idx) return i;
  if (p.queue_cap) {
    const int nf = __ldg(p.n_frames_dev);
    return __ldg(p.frame_idx + (i < nf ? i : p.queue_cap - 1 - (i - nf)));
  }
  return __ldg(p.frame_idx + i);
}
// thread 0 of a CTA: the next frame of the queue, or -1 when it is empty
__device__ __forceinline__ int next_frame(const DecParams &p) {
  const int i = (int)atomicAdd(p.work_counter, 1u);
  if (i >= frame_count(p)) return -1;
  return frame_at(p, i);
}
#endif

enum DecKernelKind { DEC_REG_6_3 = 0, DEC_REG_12_6 = 1, DEC_GEN_4_8 = 2, DEC_GEN_9_10 = 3, DEC_GEN_16_32 = 4 };

typedef void (*dec_kernel_t)(const DecParams);

struct DecLaunch {
  DecKernelKind kind;
  int alg;  // 0 = sum-product (reference semantics, bp_decode.cu), 1 = normalised min-sum fp32, 2 = min-sum fp16 x 2 frames,
            // 3 = layered min-sum (quasi-cyclic codes, bp_layered.cu)
  int threads;
  int smem_bytes;
  int ctas_per_sm;  // filled by dec_prepare (occupancy query)
  dec_kernel_t fn;  // filled by dec_prepare
  int soft;         // this record launches the kernel that also produces DecParams::out_soft
  int qc_plan;      // != 0: the graph matches a compiled quasi-cyclic plan (bp_qc_kernel); sum-product, no soft output
  int rowmajor;     // sum-product kernels: messages at row_stride * slot + k (their own DecTables) instead of planar
};

dec_kernel_t minsum_kernel_of(DecKernelKind k, int alg);
// bp_layered.cu (algorithm = 3: layered min-sum for quasi-cyclic codes)
dec_kernel_t layered_kernel();
int layered_threads(int z);
int layered_smem_bytes(int n, int n_layers, int n_edges, int z);
int layered_max_degree();
int layered_max_threads();
int layered_max_edges();
int layered_max_layers();
int layered_frames_per_cta();
inline bool dec_two_frames_per_cta(DecKernelKind k, int alg) { return alg == 2 && (k == DEC_REG_6_3 || k == DEC_REG_12_6); }
bool dec_wants_rowmajor(DecKernelKind k, int alg);
int dec_regular_threads(DecKernelKind k);
int dec_generic_max_threads();
int dec_generic_row_stride(DecKernelKind k);
int dec_match_qc_plan(int n, int m_pad, const uint8_t *vdeg, const uint8_t *cndeg, int dv_max, int row_stride);
// does the kernel behind `l` write DecParams::out_synd itself (else: launch_syndrome_weight on its packed decisions)
bool dec_has_synd_output(const DecLaunch &l, bool soft);
cudaError_t dec_prepare(DecLaunch &l);
cudaError_t dec_launch(const DecLaunch &l, const DecParams &p, int num_sms, cudaStream_t s);

// ---- generation: Philox bits, GF(2) encoder, mapping + block-fading AWGN channel --------------------------------
struct GenParams {
  int B, k, k_words, n_tx, tx_words, n_chk, punct, is_5g, encoder_active;
  int bits_per_symbol, n_sym, q;
  uint64_t seed, frame0;
  float sigma_over_sqrt2;
  const uint32_t *enc_t;  // [k_words][n_chk] transposed parity rows
  const float2 *points;   // [q]
};
cudaError_t launch_gen_bits(const GenParams &g, uint32_t *u_packed, cudaStream_t s);
cudaError_t launch_encode(const GenParams &g, const uint32_t *u_packed, uint32_t *c_packed, cudaStream_t s);
// launch_gen_bits + launch_encode in one kernel (same bits: the Philox stream is a function of the frame index only)
cudaError_t launch_gen_encode(const GenParams &g, uint32_t *u_packed, uint32_t *c_packed, cudaStream_t s);
// noise == nullptr → Philox noise and fading (h written to h_out); otherwise h is read from h_in, noise as given
cudaError_t launch_channel(const GenParams &g, const uint32_t *c_packed, const float2 *h_in, const float2 *noise,
                           float2 *h_out, float2 *y, cudaStream_t s);

// ---- k-means blind channel estimate ------------------------------------------------------------------------------
// fp64 constants of the warp kernel (host_code.cpp fills them from the constellation): s_0, 1 / s_0 and, for each Voronoi
// neighbour t of s_0, ds_t = s_t - s_0 and dn_t = (|s_t|^2 - |s_0|^2) / 2 — the half-plane of neighbour t under the
// estimate h is  Re(conj(ds_t h) y) <= dn_t |h|^2.
struct KmConst {
  double s0r, s0i, is0r, is0i;
  double dsr[8], dsi[8], dn[8];
  int n_nb;  // 0 → the general kernel (full distance comparison against every constellation point)
};
// y: float2 [B][n_sym], or double2 when y_is_f64 (then y32_out, if given, receives the fp32 copy the demapper reads).
// hhat64 (optional): the estimate as carried (fp64).
cudaError_t launch_kmeans(int B, const void *y, int y_is_f64, int n_sym, const float2 *points, int q, const KmConst &kc,
                          int iters, float2 *hhat, double2 *hhat64, int32_t *passes, float2 *y32_out,
                          int num_sms, cudaStream_t s);
cudaError_t launch_f64_to_f32(size_t n, const double *in, float *out, cudaStream_t s);
cudaError_t launch_p0_to_lr(size_t n, const double *p0, float *lr, cudaStream_t s);

// ---- soft demapper + candidate resolver ---------------------------------------------------------------------------
struct DemapParams {
  int B, n_sym, n_tx, bits_per_symbol, q, n_cand;  // n_cand = 4 (blind) or 1
  int hard_metric;                                 // compute syndrome weights of the inverted hard decisions
  int winner_only;                                 // (hard_metric, 4 candidates) keep the ratios in shared memory and
                                                   // write only the chosen candidate's: lr is then [B][n_tx]
  int m_rows, punct;
  const float2 *y;       // [B][n_sym]
  const float2 *h;       // [B] channel estimate (hhat or true h)
  float inv_var;
  float2 rot[4];         // exp(j (kPi/2) k)
  const float2 *points;  // [q]
  const int32_t *row_ptr, *col_idx;  // permuted graph
  const uint16_t *col_ell;           // [ell_width][m_rows] graph columns per row (hard metric), padded with n_tx + punct
  int ell_width;
  float *lr;             // [B][n_cand][n_tx] likelihood ratios P0/P1 in [1e-12, 1e12]
  float *metric;         // [B][4] (hard metric only)
  int32_t *kstar;        // [B]     (hard metric only; else untouched)
  // winner_only + skip_decode: a frame whose chosen candidate has syndrome weight 0 gets its decisions (out_bits[B][words_n],
  // out_ret = 1) written here — what the decoder would return at iteration 0 — and every OTHER frame is appended to the
  // decoder's queue (queue[B], *queue_n zeroed by the caller).  Needs punct == 0 and even row degrees (host checks).
  // 4 candidates: the constellation is invariant under quarter turns, s_k e^{j c pi/2} = s_{perm[c-1][k]} (host_code.cpp):
  // candidate c's point probabilities are candidate 0's, re-labelled
  int symmetric;
  int perm[3][64];
  // 64 points that form the square Gray-mapped grid compiled into demap_symbol_grid64 (host check): the 8 levels of an axis
  int grid64, grid16;  // (grid16: the 4 x 4 Gray grid of demap_symbol_grid16, levels[0..3])
  float levels[8];
  // 4 points: which 2 | 2 partition each (candidate, bit) reads (demap_symbol_q4); 0xFFFFFFFF: general path
  uint32_t q4_code;
  int skip_decode, words_n;
  uint32_t *out_bits;
  int32_t *out_ret;
  int32_t *queue, *queue_n;   // two-ended (see DecParams): queue[queue_cap], queue_n[2] = {front, back} lengths
  int queue_cap, long_metric; // a winner with more than long_metric unsatisfied checks goes to the front (it will hardly converge)
};
cudaError_t launch_demap(const DemapParams &d, int num_sms, cudaStream_t s);

// syndrome weight of packed decisions: bits [F][words_n] → metric[F] (float)
// frames whose chosen candidate's metric decode converged take its result; the others are appended to queue (*queue_n zeroed by the caller)
cudaError_t launch_reuse_metric(int B, int nbits, int bit_offset, int words_n, const int32_t *kstar, const float *metric,
                                const int32_t *mret, const uint32_t *cand_bits, uint32_t *uu_hat, int32_t *ret, int32_t *queue,
                                int32_t *queue_n, cudaStream_t s);
cudaError_t launch_extract_bits_queue(int max_frames, const int32_t *queue, const int32_t *queue_n, int nbits, int bit_offset,
                                      int src_words, const uint32_t *src, uint32_t *dst, cudaStream_t s);
cudaError_t launch_syndrome_weight(int F, const uint32_t *bits, int words_n, int m_rows, const int32_t *row_ptr,
                                   const int32_t *col_idx, float *metric, cudaStream_t s);
cudaError_t launch_abs_inplace(int n, float *v, cudaStream_t s);
// the reference's stale-syndrom_soft_ chain of the soft metric (see soft_chain_kernel)
struct SoftChainParams {
  int B, final_decode;    // final_decode = 0: GetMetrics only (histogram mode, kml_resolve)
  const double *own;      // [B][4] sum of ln(syndrom_soft) each metric decode produced itself
  const int32_t *mret;    // [B][4] its return value; 1 = left at iteration 0 = syndrom_soft_ untouched
  const double *fown;     // [B] the same of the frame's final decode (valid once it has been through the queue)
  const int32_t *fret;    // [B]
  double *carry;          // [1] in: the sum as the previous call left it; out: as this batch leaves it
  float *metric;          // [B][4] |metric| (kmcodec.cc:137)
  int32_t *kstar;         // [B] -1 = not chosen yet
  double *state;          // [B] value after the frame
  int32_t *flags;         // [B] zeroed by the caller
  int32_t *queue;         // [B] frames chosen in this round (to be decoded before the next round)
  int32_t *counts;        // [2] queue length of this round, frames whose state is still unknown
};
cudaError_t launch_soft_chain(const SoftChainParams &p, int round, cudaStream_t s);
cudaError_t launch_argmin4(int B, const float *metric, int32_t *kstar, cudaStream_t s);

// ---- small utilities ------------------------------------------------------------------------------------------------
cudaError_t launch_pack_bits(int F, int nbits, const int32_t *bits, uint32_t *packed, cudaStream_t s);
cudaError_t launch_unpack_bits(int F, int nbits, int bit_offset, int src_words, const uint32_t *packed, int32_t *bits,
                               cudaStream_t s);
cudaError_t launch_extract_bits(int F, int nbits, int bit_offset, int src_words, const uint32_t *src, uint32_t *dst,
                                cudaStream_t s);
cudaError_t launch_lr_to_llr(size_t n, const float *lr, float *llr, cudaStream_t s);
// counters[0..3] += tot_blk, err_blk, tot_bit, err_bit ; counters[4] += sum of min(ret, max_iter) when ret != nullptr
cudaError_t launch_count_errors(int B, int k, int k_words, const uint32_t *u_packed, const uint32_t *uu_hat_packed,
                                const int32_t *ret, int max_iter, unsigned long long *counters, cudaStream_t s);

// best-of-2 LDS.128 bandwidth of the device in GB/s (roofline denominator of the decoder)
cudaError_t measure_smem_bandwidth(int num_sms, double *gbs, cudaStream_t s);

}  // namespace kml
#endif
