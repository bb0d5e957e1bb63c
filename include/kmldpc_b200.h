/* kmldpc_b200 — C ABI of the B200-native (sm_100a) Monte-Carlo link path of trganda/kmldpc.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  Every entry point cites the
 * reference interface it replaces (paths relative to the reference's kmldpc/ directory).  All functions return 0 on
 * success or a negative kml_status; kml_last_error() gives the message.  No exceptions, no exit() (the reference
 * exit(-1)s on file errors, lib/lab/src/binaryldpccodec.cc:77-80).  There is NO CPU fallback: kml_create() fails
 * with KML_ERR_CUDA when no sm_100 device is usable.
 *
 * Layout conventions: batch-major, frame-contiguous; complex numbers interleaved (re, im) float32; bits as int32
 * (0/1) in the stage entry points — exactly the reference's `int *uu / *cc` — and bit-packed uint32 words
 * (bit t of a frame = word t/32, bit t%32) where a name says `_packed`.  Pointers are HOST pointers unless the
 * function name ends in `_dev`; `_dev` variants take device pointers plus a cudaStream_t passed as void*.
 * A context is owned by one host thread.  `_dev` calls are asynchronous on the caller's stream and share
 * one work space per context: calls issued on different streams are ordered on the device (each waits for the event the
 * previous one recorded), so they are safe but do not overlap — use one context per concurrent stream.
 */
#ifndef KMLDPC_B200_H
#define KMLDPC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  KML_OK = 0,
  KML_ERR_ARG = -1,     /* bad argument / unsupported configuration */
  KML_ERR_IO = -2,      /* file could not be opened or parsed */
  KML_ERR_CUDA = -3,    /* CUDA runtime error or no usable device */
  KML_ERR_STATE = -4,   /* call order / capacity problem */
  KML_ERR_NCCL = -5,    /* NCCL could not be loaded / initialised, or a collective failed */
} kml_status;

/* ------------------------------------------------------------------------------------------------------------
 * Host-side code / constellation descriptions (immutable once built; borrowed by kml_create, copied to the device)
 * ---------------------------------------------------------------------------------------------------------- */

/* LDPC code after the reference's Gaussian elimination + column permutation.
 * Replaces: BinaryLDPCCodec::BinaryLDPCCodec(toml) + SystemMatrixH (lib/lab/src/binaryldpccodec.cc:62-129,346-493)
 *           Binary5GLDPCCodec ctor + SystemMatrixH (lib/lab/src/binary5gldpccodec.cc:12-80,240-391). */
typedef struct kml_code {
  int32_t n_rows;        /* M: checks */
  int32_t n_graph;       /* columns of H (variables in the Tanner graph) */
  int32_t n_tx;          /* transmitted bits per frame (= n_graph - puncture) */
  int32_t k;             /* information bits: code_dim() */
  int32_t n_chk;         /* rank found by the elimination: code_chk_ */
  int32_t puncture;      /* 2Z leading graph columns that are never transmitted (5G), else 0 */
  int32_t info_offset;   /* uu_hat = cc_hat[info_offset .. info_offset+k): code_chk_ (PEG) or 0 (5G) */
  int32_t n_edges;       /* E */
  int32_t is_5g;
  int32_t encoder_active;/* [ldpc] active */
  int32_t enc_words;     /* uint32 words per row of enc_rows = ceil(k/32) */
  int32_t reserved;
  const int32_t *row_ptr;   /* [M+1]  CSR of the PERMUTED H, columns ascending within a row */
  const int32_t *col_idx;   /* [E] */
  const int32_t *perm;      /* [n_graph] new column j = original column perm[j] (tempP) */
  const uint32_t *enc_rows; /* [n_chk][enc_words] parity part of the reduced matrix: bit j of row t multiplies
                               info bit j; parity bit t = XOR_j u_j & enc[t][j].  NULL when !encoder_active */
} kml_code;

/* Parses an H file ("num_of_row--num_of_col--rank_of_H" format) and runs the bit-packed elimination.
 * The result is identical (permutation, reduced matrix, graph) to the reference's byte-matrix SystemMatrixH. */
int kml_code_load(const char *h_file, int is_5g, int encoder_active, kml_code **out);
void kml_code_free(kml_code *code);

/* Constellation. Replaces Modem::init (lib/lab/src/modem.cc:87-129): points scaled to unit mean energy;
 * the label of point i is i, MSB first. */
typedef struct kml_modem {
  int32_t bits_per_symbol; /* m */
  int32_t n_points;        /* Q = 2^m */
  const double *points;    /* [Q][2] (re, im) after normalisation */
} kml_modem;
int kml_modem_load(const char *modem_file, kml_modem **out);
void kml_modem_free(kml_modem *modem);

/* ------------------------------------------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct kml_opts {
  int32_t max_iter;      /* [ldpc] max_iter (binaryldpccodec.cc:70) */
  int32_t known_h;       /* [decoder] true_h_arg (simulator.cc:14-15) */
  int32_t metric_type;   /* [xcodec] metric_type: 0 hard syndrome weight, 1 soft syndrome (kmcodec.cc:146-155) */
  int32_t metric_iter;   /* [xcodec] metric_iter (kmcodec.cc:25) */
  int32_t kmeans_iter;   /* 20 (simulator.cc:140) */
  int32_t early_exit;    /* 1 = reference semantics (stop at the first zero syndrome, binaryldpccodec.cc:231);
                            0 = fixed-iteration timing mode: all max_iter iterations are executed but the decision
                            and return value are latched at the first zero syndrome, so RESULTS ARE IDENTICAL */
  int32_t max_batch;     /* frames per launch the workspaces are sized for (0 = default 16384) */
  int32_t algorithm;     /* 0 = flooding sum-product = the reference's decoder (parity mode);
                            1 = normalised min-sum, fp32 messages (throughput mode; NOT in the reference, hence not
                                reference-pinned: checked against oracle/minsum_ref.py and gated by BER/FER against 0);
                            2 = the same with fp16 messages, two frames per shared-memory word ((3,6)-regular codes and
                                the quasi-cyclic 5G BG2 plan; other graphs run algorithm 1);
                            3 = layered (row-serial) min-sum on the block structure of a quasi-cyclic code (the 5G
                                BG2 code; four frames per CTA); an error for codes without that structure */
} kml_opts;

typedef struct kml_ctx kml_ctx;

int kml_create(kml_ctx **out, int device, const kml_code *code, const kml_modem *modem, const kml_opts *opts);
void kml_destroy(kml_ctx *ctx);
const char *kml_last_error(const kml_ctx *ctx);  /* ctx may be NULL: message of the last failed create/load */
int kml_set_early_exit(kml_ctx *ctx, int early_exit);
/* Switches the decoder of every later call (also the 5G metric decodes); alpha = min-sum normalisation in (0, 1]. */
int kml_set_algorithm(kml_ctx *ctx, int algorithm, double alpha);
/* Check-node rule of the min-sum decoders (algorithm 1 | 2 | 3): |c2v| = max(alpha * min - beta, 0).  alpha = 0.8, beta = 0 is
 * the normalised min-sum kml_set_algorithm selects; alpha = 1, beta = 0.5 the offset min-sum.  Not in the reference. */
int kml_set_minsum(kml_ctx *ctx, double alpha, double beta);
/* info[0..7] = n_rows, n_graph, n_tx, k, bits_per_symbol, n_points, n_symbols per frame, max_batch */
int kml_info(const kml_ctx *ctx, int32_t info[8]);
/* Decoder launch facts: info[0..7] = kernel kind (bits 0-7: 0/1 = (3,6)-regular PEG2304 / PEG8064 shapes, 2-4 = run-time
 * graph; bits 8-15: id of the compile-time quasi-cyclic plan in use, 0 = none; bit 16: row-major messages), threads per CTA, dynamic shared memory bytes, CTAs per SM,
 * layout annealing residual, shared-memory wavefronts above one per variable-node gather (0 = conflict free),
 * variable-node gather instructions per iteration (the ideal wavefront count), row slots. */
int kml_decoder_info(const kml_ctx *ctx, int32_t info[8]);
/* Shared-memory load bandwidth this device sustains (GB/s; conflict-free LDS.128 on every SM, best of 2 launches):
 * the measured denominator of the decoder's roofline.  No reference counterpart — diagnostics. */
int kml_measure_smem_bandwidth(kml_ctx *ctx, double *gb_per_s);
/* Number of kernels launched by this context since creation (bench.py's gpu_launches). */
uint64_t kml_launch_count(const kml_ctx *ctx);

/* ------------------------------------------------------------------------------------------------------------
 * Stage entry points (host buffers; parity tests and the standalone roofline runs call these)
 * ---------------------------------------------------------------------------------------------------------- */

/* A2  BinaryLDPCCodec::Encoder (binaryldpccodec.cc:144-162) / Binary5GLDPCCodec::Encoder (binary5gldpccodec.cc:86-109)
 *     u[B][k] -> c[B][n_tx] (int32 0/1).  With encoder_active = 0 the reference zeroes BOTH u and c; here c = 0. */
int kml_encode(kml_ctx *ctx, int B, const int32_t *u, int32_t *c);

/* A1+A3+A5  SourceSink::GetBitStr (sourcesink.cc:5-9), the h draw (simulator.cc:120-128) and
 *     ModemLinearSystem::PartitionModemLSystem (modemlinearsystem.cc:30-48) with the LCG replaced by Philox4x32-10
 *     keyed by `seed`, counter = (stream, frame index, sample).  Frame indices are frame0 .. frame0+B-1, so results
 *     do not depend on batch size or GPU count.  Outputs (any may be NULL): u[B][k], c[B][n_tx] int32,
 *     h[B][2], y[B][n_sym][2] float32.  sigma^2 = 10^(-snr_db/10) (simulator.cc:74-77). */
int kml_generate(kml_ctx *ctx, int B, double snr_db, uint64_t seed, uint64_t frame0, int32_t *u, int32_t *c,
                 float *h, float *y);

/* A4+A5 with caller-supplied bits, fading and noise (deterministic: used to replay reference frames):
 *     y = h*x(c) + (sigma/sqrt2)*noise, noise[B][n_sym][2] standard normals or NULL for no noise. */
int kml_modulate(kml_ctx *ctx, int B, const int32_t *c, const float *h, const float *noise, double sigma, float *y);

/* A6  kmldpc::KMeans::Run + clusters()[0]/constellations[0] (src/kmeans.cc:15-84, simulator.cc:134-145),
 *     compiled semantics (cumulative sums, cluster 0 anchor).  y[B][n_sym][2] -> hhat[B][2]; passes[B] (may be NULL)
 *     = number of assignment passes executed. */
int kml_kmeans(kml_ctx *ctx, int B, const float *y, float *hhat, int32_t *passes);
/*     The same on the reference's own types: y = std::vector<std::complex<double>> received symbols of B frames
 *     (KMeans ctor, src/kmeans.cc:4-10; ModemLinearSystem::GetRecvSymbol) as double [B][n_sym][2]; hhat[B][2] double =
 *     clusters()[0] / constellations[0] as carried in fp64.  The samples are narrowed to fp32 on the device only for the
 *     kernel's first-pass filter; every assignment that the filter cannot decide, the anchor choice and all sums use
 *     the double values, so the estimate follows the reference to ~1e-12 relative instead of the ~1e-7 an fp32 input
 *     allows (tests/test_gpu_parity.py). */
int kml_kmeans_f64(kml_ctx *ctx, int B, const double *y, double *hhat, int32_t *passes);

/* A7+A8  ModemLinearSystem::DeMapping (modemlinearsystem.cc:51-98, modem.cc:23-79) with all bit priors 0.5
 *     (kmcodec.cc:92-103).  One channel estimate per frame: h[B][2].  Output llr[B][n_tx] = ln(P(bit=0)/P(bit=1))
 *     with the reference's clipping order, i.e. |llr| <= ln((1-1e-12)/1e-12). */
int kml_demap(kml_ctx *ctx, int B, const float *y, const float *h, double var, float *llr);

/* A9  KmCodec::GetMetrics / Metric / GetParityCheck + first argmin (kmcodec.cc:54-66,105-163): the four candidates
 *     hhat*exp(j*(kPi/2)*k).  Outputs metric[B][4] (float: syndrome weights, or |sum ln syndrom_soft|) and kstar[B]. */
int kml_resolve(kml_ctx *ctx, int B, const float *y, const float *hhat, double var, float *metric, int32_t *kstar);

/* A10 BinaryLDPCCodec::Decoder / Binary5GLDPCCodec::Decoder (binaryldpccodec.cc:165-278, binary5gldpccodec.cc:112-232).
 *     llr[B][n_tx] -> cc_hat[B][n_graph], uu_hat[B][k] (int32, either may be NULL), ret[B] = iter + (iter < max_iter)
 *     exactly like the reference's return value.  iter_count as in Decoder(M2V, uu_hat, iter_count). */
int kml_decode(kml_ctx *ctx, int B, const float *llr, int iter_count, int32_t *cc_hat, int32_t *uu_hat, int32_t *ret);
/*     The reference's own signature, BinaryLDPCCodec::Decoder(const double *M2V, int *uu_hat, int iter_count)
 *     (binaryldpccodec.cc:165): p0[B][n_tx] = P(bit = 0) in double, as KmCodec::DeMapping leaves it in bit_l_out_
 *     (kmcodec.cc:92-103).  The ratio P0 / (1 - P0) is formed in fp64 on the device. */
int kml_decode_p0(kml_ctx *ctx, int B, const double *p0, int iter_count, int32_t *cc_hat, int32_t *uu_hat, int32_t *ret);

/* KmCodec::Decoder (kmcodec.cc:54-72) preceded by the k-means block of Simulator::run_blocks (simulator.cc:131-148):
 *     the whole receiver for B frames.  y[B][n_sym][2]; true_h[B][2] is read only when opts.known_h.
 *     Outputs (any may be NULL): uu_hat_packed[B][ceil(k/32)], hhat[B][2], kstar[B], ret[B], metric[B][4] = the four
 *     candidate metrics GetMetrics returned (kmcodec.cc:122-139; blind detection only) — with hhat and kstar, everything
 *     the reference's per-frame debug lines print ("Hhat = … Metric = …", "hatIndex = …", kmcodec.cc:64,132-136). */
int kml_receive(kml_ctx *ctx, int B, const float *y, const float *true_h, double var, uint32_t *uu_hat_packed,
                float *hhat, int32_t *kstar, int32_t *ret, float *metric);

/*     Pipelined form of kml_receive for callers that hand over batch after batch: kml_receive_submit enqueues the batch
 *     (H2D copy of y from the caller's — ideally pinned — buffer, the receiver kernels, D2H copies of the outputs) and
 *     returns; kml_receive_wait(ctx, n) blocks until at most n submitted batches are still in flight (n = 0: all done).
 *     Input and output buffers of a batch belong to the library until it has been waited for.  Up to 4 batches may be in
 *     flight (a fifth submit first waits for the oldest); consecutive batches overlap their copies with each other's
 *     kernels, which the blocking call can only do inside one batch. */
int kml_receive_submit(kml_ctx *ctx, int B, const float *y, const float *true_h, double var, uint32_t *uu_hat_packed,
                       float *hhat, int32_t *kstar, int32_t *ret, float *metric);
int kml_receive_wait(kml_ctx *ctx, int max_outstanding);

/*     The same seam on the reference's types (the arguments of KmCodec::Decoder, kmcodec.cc:54-72, batched):
 *     y[B][n_sym][2] and true_h[B][2] are std::complex<double> as double pairs; hhat[B][2] double.  The conversion to the
 *     kernels' fp32 happens on the device, inside this call. */
int kml_receive_f64(kml_ctx *ctx, int B, const double *y, const double *true_h, double var, uint32_t *uu_hat_packed,
                    double *hhat, int32_t *kstar, int32_t *ret, float *metric);

/* Soft-syndrome metric ([xcodec] metric_type = true) only: the reference's syndrom_soft_ array is written by the
 * decoder's check-node phase alone (binaryldpccodec.cc:274), so a Decoder call that leaves at iteration 0
 * (binaryldpccodec.cc:231-232) leaves it as the PREVIOUS call did, and Metric() (kmcodec.cc:146-155) then sums stale
 * values.  The context carries that state — the sum of ln(syndrom_soft_[r]) after its last Decoder call — from frame to
 * frame and call to call, in frame order, exactly like one reference codec object does (a fresh context starts from
 * syndrom_soft_ = 1, i.e. 0; the reference's array starts uninitialised).  set = 0 reads it into *value, set = 1
 * writes it (e.g. to start an independent run). */
int kml_soft_syndrome_state(kml_ctx *ctx, int set, double *value);

/* A11 SourceSink::CntErr (sourcesink.cc:29-47) for B frames: counters[4] += {tot_blk, err_blk, tot_bit, err_bit}
 *     (64-bit; the reference's 32-bit counters wrap at 3.7 M frames).  u_packed / uu_hat_packed [B][ceil(k/32)]. */
int kml_count_errors(kml_ctx *ctx, int B, const uint32_t *u_packed, const uint32_t *uu_hat_packed,
                     uint64_t counters[4]);

/* ------------------------------------------------------------------------------------------------------------
 * Fused path: the body of Simulator::run / run_blocks (simulator.cc:70-168) for one SNR point
 * ---------------------------------------------------------------------------------------------------------- */

/* Runs frames [frame_begin, frame_begin+frame_count) of SNR point `snr_db` entirely on the device (Philox bits ->
 * encode -> map -> channel -> k-means -> resolve -> demap -> decode -> count) in batches of max_batch and ADDS to
 * counters[4] = {tot_blk, err_blk, tot_bit, err_bit}.  Stops early after the batch in which err_blk (including the
 * caller's incoming counters[1]) reaches max_err_blk (0 = never), mirroring simulator.cc:117 with batch granularity.
 * iters_sum (may be NULL) accumulates the decoder iterations executed. */
int kml_simulate(kml_ctx *ctx, double snr_db, uint64_t seed, uint64_t frame_begin, uint64_t frame_count,
                 uint64_t max_err_blk, uint64_t counters[4], uint64_t *iters_sum);

/* The same for ONE batch (count <= max_batch) with the per-frame values the reference writes to its log file for every
 * frame — "Generated H = …" (simulator.cc:124-126), "Hhat = … Metric = …" per candidate (kmcodec.cc:132-136),
 * "hatIndex = …" (kmcodec.cc:64): h[count][2] the fade drawn, hhat[count][2], metric[count][4], kstar[count] (blind
 * detection only) and ret[count]; any may be NULL.  kml_sweep_run prints those lines from it when [gpu] debug = true. */
int kml_simulate_frames(kml_ctx *ctx, double snr_db, uint64_t seed, uint64_t frame_begin, int count, uint64_t counters[4],
                        float *h, float *hhat, float *metric, int32_t *kstar, int32_t *ret);

/* Histogram mode of Simulator::run_blocks (simulator.cc:154-162) + KmCodec::GetHistogramData (kmcodec.cc:74-79): frames
 * [frame_begin, frame_begin+frame_count) are generated and detected but NOT finally decoded; metrics[frame][4] receives
 * the four candidate metrics (in candidate order 0°, 90°, 180°, 270°; the caller rotates them to start at the minimum
 * like simulator.cc:155-160).  counters[4] are accumulated exactly as the reference does in this mode: CntErr runs on a
 * uu_hat the final decoder never wrote — the last metric decode's decisions (5G / soft metric) or an untouched
 * (zero) buffer (hard metric). */
int kml_histogram(kml_ctx *ctx, double snr_db, uint64_t seed, uint64_t frame_begin, uint64_t frame_count, float *metrics,
                  uint64_t counters[4]);

/* The same on GIVEN frames — the KmCodec::GetHistogramData seam itself (kmcodec.cc:74-79) followed by the CntErr of
 * simulator.cc:167: y[B][n_sym][2] received symbols, u_packed[B][ceil(k/32)] the transmitted information bits.
 * Outputs: metrics[B][4]; kstar[B] (may be NULL) = index of the first minimum, where the line of histogram_<snr>.txt
 * starts; uu_hat_packed (may be NULL) = the buffer CntErr ran on; counters[4] accumulated. */
int kml_histogram_rx(kml_ctx *ctx, int B, const float *y, double var, const uint32_t *u_packed, float *metrics,
                     int32_t *kstar, uint32_t *uu_hat_packed, uint64_t counters[4]);

/* ------------------------------------------------------------------------------------------------------------
 * Device-pointer variants (inputs already resident in HBM; asynchronous on `stream`)
 * ---------------------------------------------------------------------------------------------------------- */
int kml_generate_dev(kml_ctx *ctx, int B, double snr_db, uint64_t seed, uint64_t frame0, uint32_t *u_packed,
                     float *h, float *y, void *stream);
int kml_kmeans_dev(kml_ctx *ctx, int B, const float *y, float *hhat, int32_t *passes, void *stream);
int kml_receive_dev(kml_ctx *ctx, int B, const float *y, const float *true_h, double var, uint32_t *uu_hat_packed,
                    int32_t *ret, void *stream);
/* y[B][n_sym][2], h[B][2] in HBM -> llr[B][n_tx] (natural log) in HBM; feeds kml_decode_dev for standalone runs. */
int kml_demap_dev(kml_ctx *ctx, int B, const float *y, const float *h, double var, float *llr, void *stream);
/* llr/lr in HBM -> packed decisions in HBM.  in_is_lr = 1 when the input already holds likelihood ratios P0/P1. */
int kml_decode_dev(kml_ctx *ctx, int B, const float *llr, int in_is_lr, int iter_count, uint32_t *cc_hat_packed,
                   int32_t *ret, void *stream);
/* counters_dev: 4 x uint64 in device memory, accumulated with atomics. */
int kml_count_errors_dev(kml_ctx *ctx, int B, const uint32_t *u_packed, const uint32_t *uu_hat_packed,
                         uint64_t *counters_dev, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Sweep driver: Simulator::Simulator + Simulator::Simulate (src/simulator.cc:3-67) on top of kml_simulate
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct kml_sweep_cfg {
  double min_snr, max_snr, step_snr;                 /* [range] */
  uint64_t max_err_blk, max_num_blk;                 /* [range] maximum_error_number / maximum_block_number */
  int32_t known_h, is_5g, metric_type, metric_iter;  /* [decoder] / [xcodec] */
  int32_t max_iter, encoder_active;                  /* [ldpc] */
  int32_t histogram_enable;                          /* [histogram] enable: writes histogram_<snr>.txt */
  int32_t reduce_on_host;                            /* [gpu] reduce = "host": sum the per-GPU counters on the host
                                                        instead of one ncclAllReduce per point (default 0 = NCCL) */
  char matrix_file[512];                             /* [ldpc] matrix_file */
  char modem_file[512];                              /* [modem] modem_file */
  /* optional [gpu] table (ignored by the reference binary) */
  uint64_t seed;
  int32_t n_gpus, max_batch, early_exit, algorithm;  /* [gpu] gpus / batch / early_exit / algorithm (0 SPA, 1-3 min-sum, see kml_opts) */
  int32_t debug_frames, reserved2;                   /* [gpu] debug = true: the reference's per-frame log lines ("Generated H",
                                                        "Current Block Number", "Hhat … Metric", "hatIndex") through log_cb, frames
                                                        in index order on GPU 0 — a debugging mode, orders of magnitude slower */
} kml_sweep_cfg;

/* Minimal TOML reader for exactly the keys above (the reference parses the same file with toml11, kmldpc.cpp:29-31). */
int kml_sweep_cfg_load(const char *config_toml, kml_sweep_cfg *cfg);
/* Runs the sweep on n_gpus devices (one host thread per GPU, frame ranges of every SNR point sharded; with n_gpus > 1 the
 * per-GPU counters of a point are summed by one ncclAllReduce of 4 x uint64 over NVLink — libnccl.so.2 is opened at run
 * time; if it cannot be had the call FAILS with KML_ERR_NCCL unless cfg->reduce_on_host asks for the host sum) and fills
 * ber[n_points], fer[n_points], counters[n_points][4].  data_dir is prepended to relative file names.
 * maximum_error_number = 0 runs no frame at all, like the reference (simulator.cc:117: err_blk >= 0 holds at once).
 * log_cb (may be NULL) receives the reference-format lines ("SNR = … Total blk = …", "BER Result", …). */
int kml_sweep_points(const kml_sweep_cfg *cfg);
int kml_sweep_run(const kml_sweep_cfg *cfg, const char *data_dir, double *ber, double *fer, uint64_t *counters,
                  void (*log_cb)(const char *line, void *user), void *user);
/* seconds[0] = setup (files, contexts, NCCL), seconds[1] = the SNR points, of this process's last kml_sweep_run
 * (the reference prints one "Total time cost" for both, kmldpc.cpp:44-53). */
void kml_sweep_last_timing(double seconds[2]);


/* ------------------------------------------------------------------------------------------------------------
 * Multi-GPU counter reduction (SURVEY 8(b) "comm_init / reduce_counters"; the reference sums under a mutex in
 * threadsafe_sourcesink.cc).  One process, GPUs 0 .. n_gpus-1, ncclCommInitAll; kml_sweep_run uses the same code.
 * per_gpu[n_gpus][count] uint64 host words are copied to the GPUs, summed by ONE ncclAllReduce over NVLink and the
 * result read back from GPU 0 into total[count] (count <= 1024).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct kml_comm kml_comm;
int kml_comm_init(int n_gpus, kml_comm **out);
int kml_reduce_counters(kml_comm *comm, const uint64_t *per_gpu, int count, uint64_t *total);
void kml_comm_destroy(kml_comm *comm);

#ifdef __cplusplus
}
#endif
#endif /* KMLDPC_B200_H */
