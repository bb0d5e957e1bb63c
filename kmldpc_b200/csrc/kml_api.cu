// C ABI of libkmldpc_b200.so: context, stage entry points, the chained receiver and the fused simulate loop.
// See include/kmldpc_b200.h for the contract and the reference functions each entry point replaces.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kml_internal.h"
#include "kml_kernels.cuh"

using namespace kml;

namespace {

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    n = count;
    return cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
  }
};

// Work space of one in-flight batch.  Two lanes let copies of one batch overlap kernels of the other.
struct Lane {
  cudaStream_t stream = nullptr;
  DevBuf<uint32_t> u_packed, c_packed, uu_hat_packed, cc_hat_packed;
  DevBuf<float2> h, y, hhat, noise;
  DevBuf<double2> y64, hhat64, h64;  // staging of the reference-typed entry points (complex<double>), allocated on demand
  DevBuf<double> p0_io;
  DevBuf<float> lr, metric, llr_io;
  DevBuf<double> soft, fsoft, chain_state;   // soft-syndrome metric: own sums of the metric / final decodes, chain values
  DevBuf<int32_t> kstar, ret, mret, passes, bits_io, chain_flags, chain_queue, chain_counts;
  DevBuf<int32_t> dec_queue, dec_queue_n;    // frames the demapper hands to the decoder (the rest left at iteration 0)
  DevBuf<unsigned int> work_counter;
  DevBuf<unsigned long long> counters;  // 5 x u64: what THIS lane's batches of the current kml_simulate call counted
  // The work space above belongs to ONE stream at a time.  Every entry point calls lane_acquire() before it enqueues
  // work that touches it and lane_release() after: a call on another stream first waits (on the device) for the event
  // the previous user recorded, so calls on different streams of one context serialise instead of racing.
  cudaEvent_t owner_ev = nullptr;
  cudaStream_t owner = nullptr;
  bool owned = false;
};

}  // namespace

struct kml_ctx {
  int device = 0, num_sms = 148;
  // code / modem
  int M = 0, N = 0, n_tx = 0, K = 0, n_chk = 0, punct = 0, info_offset = 0, E = 0, is_5g = 0, active = 0;
  int k_words = 0, tx_words = 0, words_n = 0, bits = 0, Q = 0, n_sym = 0;
  int even_rows = 0;  // every check has even degree: a word and its complement have the same syndrome
  kml_opts opts{};
  int max_batch = 0;
  float2 rot[4];
  // device tables
  DevBuf<uint32_t> enc_t;
  DevBuf<float2> points;
  DevBuf<int32_t> row_ptr, col_idx;
  KmConst km{};  // fp64 constants of the k-means kernel (n_nb = 0: general kernel)
  int rot_symmetric = 0, rot_perm[3][64];  // s_k e^{j c pi/2} = s_{rot_perm[c-1][k]} for every k (else rot_symmetric = 0)
  uint32_t q4_code = 0xFFFFFFFFu;          // 4 points: partition code of demap_symbol_q4 (all ones: general demapper)
  int grid64 = 0, grid16 = 0;              // the constellation is the square Gray grid compiled into demap_symbol_grid64 / grid16
  float grid_levels[8] = {};
  DevBuf<uint16_t> vn_addr, vn_addr_rm, vn_addr_g, col_ell;
  DevBuf<uint32_t> vn_items, cn_items;
  int ell_width = 0;
  DevBuf<uint8_t> vn_deg, cn_deg, cn_deg_rm;
  DecTables dt{}, dt_rm{};  // planar layout (every kernel) / row-major layout (regular sum-product kernels)
  DecLaunch dl{}, dl_alg[4]{};  // dl = the active algorithm's launch record
  DevBuf<int32_t> lay_ptr;      // block structure for the layered decoder (algorithm = 3); lay.z == 0: the code is not quasi-cyclic
  DevBuf<uint32_t> lay_cs;
  LayeredTables lay{};
  std::string lay_why;          // why algorithm = 3 is not available for this code
  DecLaunch dl_soft{};          // sum-product kernel that also produces the soft-syndrome sums (metric_type = true)
  float alpha = 0.8f, beta = 0.0f;  // min-sum: normalisation and offset
  static constexpr int kLanes = 4;  // kml_simulate rotates over all of them; the receiver entry points use the first two
  Lane lane[kLanes];
  DevBuf<unsigned long long> counters;  // 5 x u64
  unsigned long long *h_counters = nullptr;  // pinned
  static constexpr int kRxRing = 4;          // kml_receive_submit calls that may be outstanding
  cudaEvent_t rx_done[kRxRing][2] = {};      // completion of a submit on each lane
  uint64_t rx_submitted = 0, rx_waited = 0;
  DevBuf<double> soft_carry;  // [1] sum of ln(syndrom_soft_) as the last Decoder call of this context left it
  int32_t *h_chain_counts = nullptr;  // pinned [2]
  uint64_t launches = 0;
  int layout_residual = 0, layout_excess = 0, layout_excess_planar = 0;  // layout_opt.cpp: annealing cost left / wavefronts above the ideal
  std::string err;
};

namespace {

#define KML_CUDA(ctx, expr)                                                                      \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                          \
      return KML_ERR_CUDA;                                                                       \
    }                                                                                            \
  } while (0)
#define KML_LAUNCH(ctx, expr)   \
  do {                          \
    KML_CUDA(ctx, expr);        \
    (ctx)->launches++;          \
  } while (0)
#define KML_RC(expr)                  \
  do {                                \
    int rc__ = (expr);                \
    if (rc__ != KML_OK) return rc__;  \
  } while (0)

template <class T>
int ensure(kml_ctx *c, DevBuf<T> &b, size_t count) {
  if (b.n >= count && b.p) return KML_OK;
  b.release();
  KML_CUDA(c, b.alloc(count));
  return KML_OK;
}

int fail_arg(kml_ctx *ctx, const char *msg) {
  if (ctx) ctx->err = msg;
  else set_global_error(msg);
  return KML_ERR_ARG;
}

GenParams gen_params(const kml_ctx *c, int B, double snr_db, uint64_t seed, uint64_t frame0) {
  GenParams g{};
  g.B = B; g.k = c->K; g.k_words = c->k_words; g.n_tx = c->n_tx; g.tx_words = c->tx_words; g.n_chk = c->n_chk;
  g.punct = c->punct; g.is_5g = c->is_5g; g.encoder_active = c->active;
  g.bits_per_symbol = c->bits; g.n_sym = c->n_sym; g.q = c->Q;
  g.seed = seed; g.frame0 = frame0;
  const double var = std::pow(10.0, -0.1 * snr_db);  // simulator.cc:74-77
  g.sigma_over_sqrt2 = (float)(std::sqrt(var) / 1.4142135623730950488016);
  g.enc_t = c->enc_t.p;
  g.points = c->points.p;
  return g;
}

int alloc_lane(kml_ctx *c, Lane &l) {
  const size_t B = (size_t)c->max_batch;
  KML_CUDA(c, cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
  KML_CUDA(c, cudaEventCreateWithFlags(&l.owner_ev, cudaEventDisableTiming));
  KML_CUDA(c, l.u_packed.alloc(B * c->k_words));
  KML_CUDA(c, l.c_packed.alloc(B * c->tx_words));
  KML_CUDA(c, l.uu_hat_packed.alloc(B * c->k_words));
  KML_CUDA(c, l.cc_hat_packed.alloc(4 * B * c->words_n));
  KML_CUDA(c, l.h.alloc(B));
  KML_CUDA(c, l.y.alloc(B * c->n_sym));
  KML_CUDA(c, l.hhat.alloc(B));
  KML_CUDA(c, l.lr.alloc(4 * B * c->n_tx));
  KML_CUDA(c, l.metric.alloc(4 * B));
  KML_CUDA(c, l.kstar.alloc(B));
  KML_CUDA(c, l.ret.alloc(B));
  KML_CUDA(c, l.mret.alloc(4 * B));
  KML_CUDA(c, l.dec_queue.alloc(B));
  KML_CUDA(c, l.dec_queue_n.alloc(2));
  if (c->opts.metric_type) {
    KML_CUDA(c, l.soft.alloc(4 * B));
    KML_CUDA(c, l.fsoft.alloc(B));
    KML_CUDA(c, l.chain_state.alloc(B));
    KML_CUDA(c, l.chain_flags.alloc(B));
    KML_CUDA(c, l.chain_queue.alloc(B));
    KML_CUDA(c, l.chain_counts.alloc(2));
  }
  KML_CUDA(c, l.passes.alloc(B));
  KML_CUDA(c, l.work_counter.alloc(1));
  KML_CUDA(c, l.counters.alloc(5));
  return KML_OK;
}

void free_lane(Lane &l) {
  l.u_packed.release(); l.c_packed.release(); l.uu_hat_packed.release(); l.cc_hat_packed.release();
  l.h.release(); l.y.release(); l.hhat.release(); l.noise.release(); l.lr.release(); l.metric.release();
  l.soft.release(); l.fsoft.release(); l.chain_state.release(); l.chain_flags.release(); l.chain_queue.release();
  l.chain_counts.release(); l.mret.release(); l.dec_queue.release(); l.dec_queue_n.release();
  l.llr_io.release(); l.kstar.release(); l.ret.release(); l.passes.release(); l.bits_io.release();
  l.work_counter.release(); l.counters.release();
  l.y64.release(); l.hhat64.release(); l.h64.release(); l.p0_io.release();
  if (l.owner_ev) cudaEventDestroy(l.owner_ev);
  l.owner_ev = nullptr;
  if (l.stream) cudaStreamDestroy(l.stream);
  l.stream = nullptr;
}

// see Lane: serialises users of a lane's work space across streams (device-side wait, no host block)
int lane_acquire(kml_ctx *c, Lane &l, cudaStream_t s) {
  if (l.owned && l.owner != s) KML_CUDA(c, cudaStreamWaitEvent(s, l.owner_ev, 0));
  return KML_OK;
}
int lane_release(kml_ctx *c, Lane &l, cudaStream_t s) {
  KML_CUDA(c, cudaEventRecord(l.owner_ev, s));
  l.owner = s;
  l.owned = true;
  return KML_OK;
}

// Quasi-cyclic structure for the layered decoder (bp_layered.cu): finds the largest Z for which H is a grid of Z x Z
// blocks that are zero or ONE cyclically shifted identity (row l Z + z meets column c Z + (z + s) mod Z), and uploads the
// (block column, shift) list of every block row.  Codes without that structure simply have no algorithm 3.
int build_layered_tables(kml_ctx *c, const kml_code *code) {
  const int M = c->M, N = c->N;
  c->lay = LayeredTables{};
  c->dl_alg[3] = DecLaunch{};
  c->lay_why = "the parity-check matrix is not a grid of cyclically shifted identity blocks";
  int g = M, b = N;
  while (b) { const int t = g % b; g = b; b = t; }  // Z divides gcd(M, N)
  std::vector<int32_t> ptr;
  std::vector<uint32_t> cs;
  int z_found = 0;
  for (int Z = g; Z >= 8 && !z_found; Z--) {
    if (g % Z) continue;
    ptr.assign(1, 0);
    cs.clear();
    bool ok = true;
    std::vector<std::pair<int, int>> first, mine;
    for (int l = 0; l < M / Z && ok; l++) {
      for (int z = 0; z < Z && ok; z++) {
        const int r = l * Z + z;
        mine.clear();
        for (int e = code->row_ptr[r]; e < code->row_ptr[r + 1]; e++) {
          const int col = code->col_idx[e];
          mine.emplace_back(col / Z, ((col % Z) - z + Z) % Z);
        }
        std::sort(mine.begin(), mine.end());
        for (size_t i = 1; i < mine.size(); i++) ok = ok && mine[i].first != mine[i - 1].first;  // one shift per block
        if (z == 0) first = mine;
        else ok = ok && mine == first;
      }
      if (ok) {
        for (auto &bs : first) cs.push_back(((uint32_t)(bs.first * Z) << 16) | (uint32_t)bs.second);
        ptr.push_back((int32_t)cs.size());
      }
    }
    if (ok) z_found = Z;
  }
  if (!z_found) return KML_OK;
  const int NL = M / z_found;
  int dmax = 0;
  for (int l = 0; l < NL; l++) dmax = std::max(dmax, ptr[l + 1] - ptr[l]);
  if (dmax > layered_max_degree() || (int)cs.size() > layered_max_edges() || NL > layered_max_layers() || N > 0xFFFF) {
    c->lay_why = "block structure beyond the compiled layered kernel (check degree / block count)";
    return KML_OK;
  }
  DecLaunch dl{};
  dl.kind = c->dl_alg[0].kind;
  dl.alg = 3;
  dl.threads = layered_threads(z_found);
  dl.smem_bytes = layered_smem_bytes(N, NL, (int)cs.size(), z_found);
  if (dl.threads > layered_max_threads() || dl.smem_bytes > 227 * 1024) {
    c->lay_why = "lifting size too large for the layered kernel's frames-per-CTA tiling";
    return KML_OK;
  }
  KML_CUDA(c, c->lay_ptr.alloc(ptr.size()));
  KML_CUDA(c, cudaMemcpy(c->lay_ptr.p, ptr.data(), ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  KML_CUDA(c, c->lay_cs.alloc(cs.size()));
  KML_CUDA(c, cudaMemcpy(c->lay_cs.p, cs.data(), cs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  KML_CUDA(c, dec_prepare(dl));
  c->dl_alg[3] = dl;
  c->lay.lay_ptr = c->lay_ptr.p; c->lay.lay_cs = c->lay_cs.p;
  c->lay.z = z_found; c->lay.n_layers = NL; c->lay.n_edges = (int)cs.size();
  return KML_OK;
}

// Builds the shared-memory layout of the decoder: edge (row r, position k) at word k * plane + slot(r); slot() and the
// in-row order come from layout_opt.cpp so that variable-node gathers are (nearly) bank-conflict free.
int build_decoder_tables(kml_ctx *c, const kml_code *code) {
  const int M = c->M, N = c->N;
  const int mpad = (M + 31) & ~31;
  const int plane = mpad + 1;
  std::vector<int> rdeg(M), cdeg(N, 0);
  int dcm = 0, dvm = 0;
  for (int r = 0; r < M; r++) {
    rdeg[r] = code->row_ptr[r + 1] - code->row_ptr[r];
    dcm = std::max(dcm, rdeg[r]);
    for (int e = code->row_ptr[r]; e < code->row_ptr[r + 1]; e++) cdeg[code->col_idx[e]]++;
  }
  for (int v = 0; v < N; v++) dvm = std::max(dvm, cdeg[v]);
  if ((size_t)dcm * plane >= 0xFFFFu) return fail_arg(c, "code too large for 16-bit shared-memory edge addresses");
  const bool regular = std::all_of(rdeg.begin(), rdeg.end(), [](int d) { return d == 6; }) &&
                       std::all_of(cdeg.begin(), cdeg.end(), [](int d) { return d == 3; }) && c->punct == 0 &&
                       mpad == M && N == 2 * M;
  DecLaunch dl{};
  int dv_tab;
  if (regular && N == 6 * 384) {  // PEG2304: 384 threads x (6 variables, 3 checks), 3 CTAs per SM
    dl.kind = DEC_REG_6_3; dl.threads = 384; dv_tab = 3;
    dl.smem_bytes = 6 * plane * 4;
  } else if (regular && N == 12 * 672) {  // PEG8064: 672 threads x (12, 6), one CTA (97 KB of messages) per SM
    dl.kind = DEC_REG_12_6; dl.threads = 672; dv_tab = 3;
    dl.smem_bytes = 6 * plane * 4;
  } else {
    if (dvm <= 4 && dcm <= 8) { dl.kind = DEC_GEN_4_8; }
    else if (dvm <= 9 && dcm <= 10) { dl.kind = DEC_GEN_9_10; }
    else if (dvm <= 16 && dcm <= 16) { dl.kind = DEC_GEN_16_32; }
    else return fail_arg(c, "row/column degree beyond the compiled decoder kernels (max 16/16)");
    dv_tab = dvm;
    const int tmax = dec_generic_max_threads();
    int t = std::min(tmax, std::max(128, ((N + 5) / 6 + 31) & ~31));
    for (int cand = tmax; cand >= 256; cand -= 32)  // prefer a block size that tiles the row slots exactly
      if (mpad % cand == 0 && cand * 8 >= N) { t = cand; break; }
    dl.threads = t;
    dl.smem_bytes = (dcm * plane + N + 2 * c->words_n) * 4;
  }
  if (dl.smem_bytes > 227 * 1024) return fail_arg(c, "code too large for one frame per SM in shared memory");
  // every kernel hands 32 consecutive variables to one warp instruction
  std::vector<int> group(N);
  for (int v = 0; v < N; v++) group[v] = v / 32;
  std::vector<int> erow(code->n_edges);
  for (int r = 0; r < M; r++)
    for (int e = code->row_ptr[r]; e < code->row_ptr[r + 1]; e++) erow[e] = r;
  std::vector<uint8_t> vdeg(N, 0), cndeg(mpad, 0);
  // address list of one layout: word = pos * pos_stride + slot * slot_stride, compact and in colour order (no holes:
  // the kernels unroll on the exact degree)
  auto build_addr = [&](int pos_stride, int slot_stride, std::vector<uint16_t> &vaddr, int *residual, int *excess) {
    std::vector<int> slot, pos;
    std::vector<std::vector<int>> order;
    *residual = optimize_decoder_layout(M, N, mpad, pos_stride, slot_stride, code->row_ptr, code->col_idx, group,
                                        (N + 31) / 32, dv_tab, slot, pos, order, excess);
    vaddr.assign((size_t)N * dv_tab, 0xFFFFu);
    std::fill(vdeg.begin(), vdeg.end(), 0);
    std::fill(cndeg.begin(), cndeg.end(), 0);
    for (int r = 0; r < M; r++) cndeg[slot[r]] = (uint8_t)rdeg[r];
    for (int v = 0; v < N; v++)
      for (int i = 0; i < dv_tab; i++) {
        const int e = order[v][i];
        if (e >= 0) vaddr[(size_t)v * dv_tab + vdeg[v]++] = (uint16_t)(pos[e] * pos_stride + slot[erow[e]] * slot_stride);
      }
  };
  const bool regular_kind = dl.kind == DEC_REG_6_3 || dl.kind == DEC_REG_12_6;
  const int planar_smem = regular_kind ? 6 * plane * 4 : dl.smem_bytes;
  std::vector<uint16_t> vaddr;
  // ---- planar tables: the min-sum kernels (and the sum-product kernels under KML_DEC_PLANAR=1)
  build_addr(plane, 1, vaddr, &c->layout_residual, &c->layout_excess_planar);
  c->layout_excess = c->layout_excess_planar;
  KML_CUDA(c, c->vn_addr.alloc(vaddr.size()));
  KML_CUDA(c, cudaMemcpy(c->vn_addr.p, vaddr.data(), vaddr.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  KML_CUDA(c, c->vn_deg.alloc(N));
  KML_CUDA(c, cudaMemcpy(c->vn_deg.p, vdeg.data(), N, cudaMemcpyHostToDevice));
  KML_CUDA(c, c->cn_deg.alloc(mpad));
  KML_CUDA(c, cudaMemcpy(c->cn_deg.p, cndeg.data(), mpad, cudaMemcpyHostToDevice));
  c->dt.vn_addr = c->vn_addr.p; c->dt.vn_deg = c->vn_deg.p; c->dt.cn_deg = c->cn_deg.p;
  c->dt.n = N; c->dt.m_pad = mpad; c->dt.plane = plane; c->dt.n_tx = c->n_tx; c->dt.punct = c->punct;
  c->dt.dv_max = dv_tab; c->dt.dc_max = dcm;
  c->dt.n_pad = (N + 31) & ~31;
  c->dt_rm = c->dt;
  // ---- row-major tables of the sum-product kernels: word = row_stride * slot + k
  const bool rowmajor = dec_wants_rowmajor(dl.kind, 0);
  int rm_smem = planar_smem, qc_plan = 0;
  if (rowmajor && regular_kind) {  // check nodes read their six words with LDS.64
    int res_rm = 0;
    build_addr(1, 6, vaddr, &res_rm, &c->layout_excess);
    KML_CUDA(c, c->vn_addr_rm.alloc(vaddr.size()));
    KML_CUDA(c, cudaMemcpy(c->vn_addr_rm.p, vaddr.data(), vaddr.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    c->dt_rm.vn_addr = c->vn_addr_rm.p; c->dt_rm.plane = 1;
    rm_smem = 6 * mpad * 4;
  } else if (rowmajor) {
    // generic kernel: odd row stride; address lists per warp item as [edge k][lane] (a warp's 32 loads = one 64-byte
    // line at an immediate offset); warp work lists balanced by node degree — longest-processing-time first over the
    // CTA's warps, then laid out [round][warp]
    const int rs = dec_generic_row_stride(dl.kind), n_pad = c->dt.n_pad, W = dl.threads / 32;
    if ((size_t)mpad * rs >= 0xFFFFu) return fail_arg(c, "code too large for 16-bit shared-memory edge addresses");
    int res_rm = 0;
    build_addr(1, rs, vaddr, &res_rm, &c->layout_excess);  // also refills vdeg / cndeg for THIS slot assignment
    std::vector<uint16_t> vg((size_t)n_pad * dv_tab, 0xFFFFu);
    for (int v = 0; v < N; v++)
      for (int k = 0; k < dv_tab; k++) vg[((size_t)(v / 32) * dv_tab + k) * 32 + (v & 31)] = vaddr[(size_t)v * dv_tab + k];
    // Work lists: a warp item is a RUN of consecutive groups with one degree (quasi-cyclic codes: Z / 32 groups per block
    // column), so the per-item dispatch overhead is paid once per run.  Runs are cut into chunks of at most half a
    // warp's share, and the chunks are placed longest-processing-time first.  Cost model (issue slots, from the SASS):
    // 30 per item + 17 per group + 22 per edge of a lane.
    auto balance = [&](int n_groups, auto group_deg, std::vector<uint32_t> &items) {
      struct Run { int g, deg, cnt; long cost; };
      auto gcost = [](int deg) { return 17L + 22L * std::max(deg == 0xFF ? 16 : deg, 1); };
      std::vector<Run> runs;
      long total = 0;
      for (int g = 0; g < n_groups; g++) {
        int lo, hi;
        group_deg(g, &lo, &hi);
        if (hi < 0) continue;  // nothing to do in this group
        const int deg = lo == hi ? hi : 0xFF;  // mixed degrees: looked up per lane
        total += gcost(deg);
        if (!runs.empty() && deg != 0xFF && runs.back().deg == deg && runs.back().g + runs.back().cnt == g) runs.back().cnt++;
        else runs.push_back({g, deg, 1, 0});
      }
      const long target = std::max<long>(total / W, 1);
      std::vector<Run> chunks;
      for (const Run &r : runs) {
        const int mc = (int)std::min<long>(255, std::max<long>(1, target / (2 * gcost(r.deg))));
        for (int g = r.g, left = r.cnt; left > 0;) {
          const int k = std::min(left, mc);
          chunks.push_back({g, r.deg, k, 30 + k * gcost(r.deg)});
          g += k;
          left -= k;
        }
      }
      std::stable_sort(chunks.begin(), chunks.end(), [](const Run &a, const Run &b) { return a.cost > b.cost; });
      std::vector<std::vector<uint32_t>> mine(W);
      std::vector<long> load(W, 0);
      for (const Run &ch : chunks) {
        const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        mine[w].push_back((uint32_t)ch.g | ((uint32_t)ch.deg << 16) | ((uint32_t)ch.cnt << 24));
        load[w] += ch.cost;
      }
      size_t rounds = 0;
      for (auto &m : mine) rounds = std::max(rounds, m.size());
      items.assign(rounds * W, 0xFFFFFFFFu);
      for (int ww = 0; ww < W; ww++)
        for (size_t i = 0; i < mine[ww].size(); i++) items[i * W + ww] = mine[ww][i];
    };
    std::vector<uint32_t> vi, ci;
    balance(n_pad / 32, [&](int g, int *lo, int *hi) {
      *lo = 255; *hi = 0;
      for (int v = g * 32; v < std::min(N, g * 32 + 32); v++) { *lo = std::min<int>(*lo, vdeg[v]); *hi = std::max<int>(*hi, vdeg[v]); }
      // ragged last group and degree-0 variables (they still report a decision): per-lane dispatch
      if (g * 32 + 32 > N || *lo == 0) { *lo = 0; *hi = std::max(*hi, 1); }
    }, vi);
    balance(mpad / 32, [&](int g, int *lo, int *hi) {
      *lo = 255; *hi = 0;
      for (int sl = g * 32; sl < g * 32 + 32; sl++) { *lo = std::min<int>(*lo, cndeg[sl]); *hi = std::max<int>(*hi, cndeg[sl]); }
      if (*hi == 0) *hi = -1;  // padding slots only
    }, ci);
    KML_CUDA(c, c->vn_addr_g.alloc(vg.size()));
    KML_CUDA(c, cudaMemcpy(c->vn_addr_g.p, vg.data(), vg.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    KML_CUDA(c, c->cn_deg_rm.alloc(mpad));
    KML_CUDA(c, cudaMemcpy(c->cn_deg_rm.p, cndeg.data(), mpad, cudaMemcpyHostToDevice));
    KML_CUDA(c, c->vn_items.alloc(vi.size()));
    KML_CUDA(c, cudaMemcpy(c->vn_items.p, vi.data(), vi.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    KML_CUDA(c, c->cn_items.alloc(ci.size()));
    KML_CUDA(c, cudaMemcpy(c->cn_items.p, ci.data(), ci.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    KML_CUDA(c, c->vn_addr_rm.alloc(vaddr.size()));  // [v][dv_max] in the same layout: the quasi-cyclic kernel's lists
    KML_CUDA(c, cudaMemcpy(c->vn_addr_rm.p, vaddr.data(), vaddr.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    qc_plan = dec_match_qc_plan(N, mpad, vdeg.data(), cndeg.data(), dv_tab, rs);
    c->dt_rm.vn_addr = c->vn_addr_rm.p; c->dt_rm.plane = 1; c->dt_rm.cn_deg = c->cn_deg_rm.p;
    c->dt_rm.vn_addr_g = c->vn_addr_g.p; c->dt_rm.vn_items = c->vn_items.p; c->dt_rm.cn_items = c->cn_items.p;
    c->dt_rm.vn_items_n = (int)vi.size(); c->dt_rm.cn_items_n = (int)ci.size();
    rm_smem = (mpad * rs + N + 2 * c->words_n) * 4;
    if (rm_smem > 227 * 1024) return fail_arg(c, "code too large for one frame per SM in shared memory");
  }
  for (int alg = 0; alg < 3; alg++) {
    dl.alg = alg;
    dl.qc_plan = qc_plan;  // the quasi-cyclic plan kernel has a sum-product and a min-sum variant, both on the row-major tables
    dl.rowmajor = ((alg == 0 && rowmajor) || qc_plan) ? 1 : 0;
    if (dl.kind == DEC_REG_12_6) dl.threads = dl.rowmajor ? dec_regular_threads(dl.kind) : 672;
    dl.smem_bytes = dl.rowmajor ? rm_smem : planar_smem;
    KML_CUDA(c, dec_prepare(dl));
    c->dl_alg[alg] = dl;
  }
  KML_RC(build_layered_tables(c, code));
  if (c->opts.algorithm == 3 && !c->lay.z) return fail_arg(c, ("algorithm = 3 (layered min-sum): " + c->lay_why).c_str());
  c->dl = c->dl_alg[c->opts.algorithm];
  {  // soft-syndrome twin: regular codes → scalar planar kernel on the planar tables; other graphs → run-time-graph kernel
    DecLaunch ds = c->dl_alg[0];
    ds.soft = 1;
    ds.qc_plan = 0;
    if (regular_kind) {
      ds.rowmajor = 0;
      ds.threads = dl.kind == DEC_REG_12_6 ? 672 : 384;
      ds.smem_bytes = planar_smem;
    }
    KML_CUDA(c, dec_prepare(ds));
    c->dl_soft = ds;
  }
  return KML_OK;
}

DecParams dec_params(kml_ctx *c, Lane &l, int B, const float *in, const int32_t *sel, int n_cand, int in_is_lr, int iters,
                     uint32_t *out_bits, int32_t *out_ret, double *out_soft) {
  DecParams p{};
  p.t = (out_soft ? c->dl_soft.rowmajor : c->dl.rowmajor) ? c->dt_rm : c->dt;
  p.in = in; p.sel = sel; p.n_cand = n_cand; p.in_is_lr = in_is_lr;
  p.B = B; p.iters = iters; p.max_iter = c->opts.max_iter; p.early_exit = c->opts.early_exit;
  p.out_bits = out_bits; p.out_ret = out_ret; p.out_soft = out_soft;
  p.work_counter = l.work_counter.p; p.words_n = c->words_n; p.alpha = c->alpha; p.beta = c->beta;
  p.lay = c->lay;
  return p;
}

// Metric() for the decode-based metrics (kmcodec.cc:140-163): Decoder(metric_iter) on the four candidates of every frame.
// Hard metric of the 5G codec: l.metric / l.kstar are final on return.  Soft metric: the decodes' own ln-sums land in
// l.soft and their return values in l.mret; the caller runs the syndrom_soft_ chain (soft_chain) next.
int metric_decodes(kml_ctx *c, Lane &l, cudaStream_t s, int B) {
  const bool soft = c->opts.metric_type != 0;
  if (soft) KML_CUDA(c, cudaMemsetAsync(l.soft.p, 0, sizeof(double) * 4 * (size_t)B, s));
  DecParams p = dec_params(c, l, 4 * B, l.lr.p, nullptr, 1, 1, c->opts.metric_iter, l.cc_hat_packed.p, l.mret.p, soft ? l.soft.p : nullptr);
  p.early_exit = 1;
  if (!soft && dec_has_synd_output(c->dl, false)) p.out_synd = l.metric.p;  // syndrome weight straight from the decoder
  KML_LAUNCH(c, dec_launch(soft ? c->dl_soft : c->dl, p, c->num_sms, s));
  if (soft) return KML_OK;
  if (!p.out_synd)
    KML_LAUNCH(c, launch_syndrome_weight(4 * B, l.cc_hat_packed.p, c->words_n, c->M, c->row_ptr.p, c->col_idx.p, l.metric.p, s));
  KML_LAUNCH(c, launch_argmin4(B, l.metric.p, l.kstar.p, s));
  return KML_OK;
}

// The soft metric's candidate choice with the reference's stale syndrom_soft_ semantics (see soft_chain_kernel), and —
// when final_decode — the final decodes themselves (their ln-sums feed the chain): l.metric, l.kstar, and with
// final_decode l.cc_hat_packed[B][words_n] / l.ret.  Blocks the host once per round (this mode is sequential by nature).
int soft_chain(kml_ctx *c, Lane &l, cudaStream_t s, int B, bool final_decode) {
  if (B < 1) return KML_OK;
  KML_CUDA(c, cudaMemsetAsync(l.kstar.p, 0xFF, sizeof(int32_t) * (size_t)B, s));
  KML_CUDA(c, cudaMemsetAsync(l.chain_flags.p, 0, sizeof(int32_t) * (size_t)B, s));
  KML_CUDA(c, cudaMemsetAsync(l.fsoft.p, 0, sizeof(double) * (size_t)B, s));
  KML_CUDA(c, cudaMemsetAsync(l.ret.p, 0, sizeof(int32_t) * (size_t)B, s));
  SoftChainParams q{};
  q.B = B; q.final_decode = final_decode ? 1 : 0;
  q.own = l.soft.p; q.mret = l.mret.p; q.fown = l.fsoft.p; q.fret = l.ret.p; q.carry = c->soft_carry.p;
  q.metric = l.metric.p; q.kstar = l.kstar.p; q.state = l.chain_state.p; q.flags = l.chain_flags.p;
  q.queue = l.chain_queue.p; q.counts = l.chain_counts.p;
  for (int round = 0; round <= B + 1; round++) {
    KML_LAUNCH(c, launch_soft_chain(q, round, s));
    KML_CUDA(c, cudaMemcpyAsync(c->h_chain_counts, l.chain_counts.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    KML_CUDA(c, cudaStreamSynchronize(s));
    const int queued = c->h_chain_counts[0], left = c->h_chain_counts[1];
    if (queued > 0) {  // the frames chosen in this round: final Decoder(max_iter) on the chosen candidate (kmcodec.cc:70-71)
      DecParams p = dec_params(c, l, queued, l.lr.p, l.kstar.p, 4, 1, c->opts.max_iter, l.cc_hat_packed.p, l.ret.p, l.fsoft.p);
      p.frame_idx = l.chain_queue.p;
      KML_LAUNCH(c, dec_launch(c->dl_soft, p, c->num_sms, s));
    }
    if (left == 0) return KML_OK;
    if (queued == 0) break;  // no progress: cannot happen (frame 0 never waits)
  }
  c->err = "soft-syndrome chain did not resolve";
  return KML_ERR_STATE;
}

// k-means → 4 candidates → metric → argmin → demap → decode (simulator.cc:131-148 + kmcodec.cc:54-72), all on stream s.
// y (float2, or double2 when y_is_f64 — then l.y receives the fp32 copy) and, for known_h, true_h are device pointers;
// decisions land in l.cc_hat_packed / l.ret / l.uu_hat_packed, the estimate in l.hhat (and l.hhat64 when want_h64).
int receive_on_lane(kml_ctx *c, Lane &l, cudaStream_t s, int B, const void *y, int y_is_f64, const float2 *true_h,
                    double var, bool want_h64 = false) {
  const float2 *y32 = y_is_f64 ? l.y.p : reinterpret_cast<const float2 *>(y);
  DemapParams d{};
  d.B = B; d.n_sym = c->n_sym; d.n_tx = c->n_tx; d.bits_per_symbol = c->bits; d.q = c->Q;
  d.m_rows = c->M; d.punct = c->punct; d.y = y32; d.inv_var = (float)(1.0 / var);
  for (int k = 0; k < 4; k++) d.rot[k] = c->rot[k];
  d.points = c->points.p; d.row_ptr = c->row_ptr.p; d.col_idx = c->col_idx.p; d.col_ell = c->col_ell.p; d.ell_width = c->ell_width;
  d.lr = l.lr.p; d.metric = l.metric.p; d.kstar = l.kstar.p;
  d.symmetric = c->rot_symmetric;
  std::memcpy(d.perm, c->rot_perm, sizeof d.perm);
  d.grid64 = c->grid64; d.grid16 = c->grid16;
  std::memcpy(d.levels, c->grid_levels, sizeof d.levels);
  d.q4_code = c->q4_code;
  const int32_t *sel = nullptr;
  int n_cand = 1;
  if (c->opts.known_h) {
    if (y_is_f64) KML_LAUNCH(c, launch_f64_to_f32((size_t)B * c->n_sym * 2, reinterpret_cast<const double *>(y), reinterpret_cast<float *>(l.y.p), s));
    d.h = true_h; d.n_cand = 1; d.hard_metric = 0;
    KML_LAUNCH(c, launch_demap(d, c->num_sms, s));
  } else {
    KML_LAUNCH(c, launch_kmeans(B, y, y_is_f64, c->n_sym, c->points.p, c->Q, c->km, c->opts.kmeans_iter, l.hhat.p,
                               want_h64 ? l.hhat64.p : nullptr, l.passes.p, y_is_f64 ? l.y.p : nullptr, c->num_sms, s));
    const bool decode_metric = c->is_5g || c->opts.metric_type;
    d.h = l.hhat.p; d.n_cand = 4; d.hard_metric = decode_metric ? 0 : 1;
    // hard metric: the four ratio vectors stay in shared memory and only the winner's reaches HBM — as long as that
    // leaves room for >= 4 CTAs per SM (PEG2304: 39 KB); long frames (PEG8064: 137 KB) write all four instead
    d.winner_only = (!decode_metric && 16 * (size_t)c->n_tx + c->n_tx + 64 * (size_t)c->Q <= 56 * 1024) ? 1 : 0;
    // … and a frame whose winner already satisfies every check is finished by the demapper itself (what the decoder
    // returns at iteration 0); only the others are queued for the decoder.  Reference semantics unchanged; off in the
    // fixed-iteration timing mode (early_exit = 0 promises the full iteration count for every frame).
    d.skip_decode = (d.winner_only && c->opts.early_exit && c->even_rows && c->punct == 0) ? 1 : 0;
    d.words_n = c->words_n; d.out_bits = l.cc_hat_packed.p; d.out_ret = l.ret.p;
    if (d.skip_decode) {
      d.queue = l.dec_queue.p; d.queue_n = l.dec_queue_n.p; d.queue_cap = c->max_batch;
      d.long_metric = c->M / 6;  // (a frame that converges has a handful of wrong bits, three unsatisfied checks each)
      KML_CUDA(c, cudaMemsetAsync(l.dec_queue_n.p, 0, 2 * sizeof(int32_t), s));
    }
    KML_LAUNCH(c, launch_demap(d, c->num_sms, s));
    if (d.winner_only) {
      DecParams p = dec_params(c, l, B, l.lr.p, nullptr, 1, 1, c->opts.max_iter, l.cc_hat_packed.p, l.ret.p, nullptr);
      if (d.skip_decode) { p.frame_idx = l.dec_queue.p; p.n_frames_dev = l.dec_queue_n.p; p.queue_cap = c->max_batch; }
      KML_LAUNCH(c, dec_launch(c->dl, p, c->num_sms, s));
      KML_LAUNCH(c, launch_extract_bits(B, c->K, c->info_offset, c->words_n, l.cc_hat_packed.p, l.uu_hat_packed.p, s));
      return KML_OK;
    }
    if (decode_metric) KML_RC(metric_decodes(c, l, s, B));  // Metric(): Decoder(metric_iter) on every candidate (kmcodec.cc:146-160)
    if (c->opts.metric_type) {          // soft metric: choice and final decodes are interleaved (stale syndrom_soft_ chain)
      KML_RC(soft_chain(c, l, s, B, true));
      KML_LAUNCH(c, launch_extract_bits(B, c->K, c->info_offset, c->words_n, l.cc_hat_packed.p, l.uu_hat_packed.p, s));
      return KML_OK;
    }
    sel = l.kstar.p;
    n_cand = 4;
    // A chosen candidate whose metric decode already reached a zero syndrome needs no final decode: Decoder(max_iter) on the
    // same input repeats those iterations and returns the same word and value (reuse_metric_kernel).  Off in the
    // fixed-iteration timing mode, like skip_decode above.
    if (decode_metric && c->opts.early_exit && c->opts.metric_iter <= c->opts.max_iter) {
      KML_CUDA(c, cudaMemsetAsync(l.dec_queue_n.p, 0, 2 * sizeof(int32_t), s));
      KML_LAUNCH(c, launch_reuse_metric(B, c->K, c->info_offset, c->words_n, l.kstar.p, l.metric.p, l.mret.p, l.cc_hat_packed.p,
                                        l.uu_hat_packed.p, l.ret.p, l.dec_queue.p, l.dec_queue_n.p, s));
      DecParams p = dec_params(c, l, B, l.lr.p, sel, n_cand, 1, c->opts.max_iter, l.cc_hat_packed.p, l.ret.p, nullptr);
      p.frame_idx = l.dec_queue.p; p.n_frames_dev = l.dec_queue_n.p; p.queue_cap = 0;
      KML_LAUNCH(c, dec_launch(c->dl, p, c->num_sms, s));
      KML_LAUNCH(c, launch_extract_bits_queue(B, l.dec_queue.p, l.dec_queue_n.p, c->K, c->info_offset, c->words_n,
                                              l.cc_hat_packed.p, l.uu_hat_packed.p, s));
      return KML_OK;
    }
  }
  DecParams p = dec_params(c, l, B, l.lr.p, sel, n_cand, 1, c->opts.max_iter, l.cc_hat_packed.p, l.ret.p, nullptr);
  KML_LAUNCH(c, dec_launch(c->dl, p, c->num_sms, s));
  KML_LAUNCH(c, launch_extract_bits(B, c->K, c->info_offset, c->words_n, l.cc_hat_packed.p, l.uu_hat_packed.p, s));
  return KML_OK;
}

int check_batch(kml_ctx *c, int B) {
  if (!c) return KML_ERR_ARG;
  if (B < 0) return fail_arg(c, "negative batch");
  return KML_OK;
}

}  // namespace

// ================================================================================================ context
extern "C" int kml_create(kml_ctx **out, int device, const kml_code *code, const kml_modem *modem, const kml_opts *opts) {
  if (!out || !code || !modem || !opts) return fail_arg(nullptr, "kml_create: null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_global_error("kml_create: no CUDA device — kmldpc_b200 has no CPU fallback");
    return KML_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) return fail_arg(nullptr, "kml_create: bad device index");
  cudaDeviceProp prop{};
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    set_global_error("kml_create: cannot select device");
    return KML_ERR_CUDA;
  }
  if (prop.major != 10) {
    set_global_error("kml_create: kernels are built for sm_100a only (found sm_" + std::to_string(prop.major) +
                     std::to_string(prop.minor) + ")");
    return KML_ERR_CUDA;
  }
  if (code->n_tx % modem->bits_per_symbol != 0)  // modemlinearsystem.cc:7-13
    return fail_arg(nullptr, "kml_create: n_tx is not a multiple of bits_per_symbol");
  if (modem->bits_per_symbol > 6) return fail_arg(nullptr, "kml_create: constellations above 64 points are not built");
  if (opts->max_iter < 1) return fail_arg(nullptr, "kml_create: max_iter < 1");
  if (opts->algorithm < 0 || opts->algorithm > 3) return fail_arg(nullptr, "kml_create: unknown algorithm");
  if (opts->algorithm != 0 && opts->metric_type)
    return fail_arg(nullptr, "kml_create: the soft-syndrome metric needs the sum-product decoder (algorithm = 0)");
  auto *c = new kml_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->M = code->n_rows; c->N = code->n_graph; c->n_tx = code->n_tx; c->K = code->k; c->n_chk = code->n_chk;
  c->punct = code->puncture; c->info_offset = code->info_offset; c->E = code->n_edges; c->is_5g = code->is_5g;
  c->active = code->encoder_active;
  c->k_words = (c->K + 31) / 32; c->tx_words = (c->n_tx + 31) / 32; c->words_n = (c->N + 31) / 32;
  c->even_rows = 1;
  for (int r = 0; r < c->M; r++) c->even_rows &= ((code->row_ptr[r + 1] - code->row_ptr[r]) & 1) ? 0 : 1;
  c->bits = modem->bits_per_symbol; c->Q = modem->n_points; c->n_sym = c->n_tx / c->bits;
  c->opts = *opts;
  if (c->opts.kmeans_iter <= 0) c->opts.kmeans_iter = 20;
  if (c->opts.metric_iter <= 0) c->opts.metric_iter = 5;
  c->max_batch = opts->max_batch > 0 ? opts->max_batch : 16384;
  for (int k = 0; k < 4; k++) {  // exp(j (kPi/2) k) with the reference's truncated pi (simulator.cc:146-148)
    const double a = (kRefPi / 2) * k;
    c->rot[k] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
  auto fail = [&](int rc) {
    set_global_error(c->err);
    kml_destroy(c);
    return rc;
  };
#define KML_TRY(expr)               \
  do {                              \
    int rc__ = (expr);              \
    if (rc__ != KML_OK) return fail(rc__); \
  } while (0)
  auto upload = [&]() -> int {
    // encoder matrix, transposed to [word][row] so that consecutive threads read consecutive addresses
    if (c->active) {
      const int W = code->enc_words;
      std::vector<uint32_t> t((size_t)W * c->n_chk);
      for (int r = 0; r < c->n_chk; r++)
        for (int w = 0; w < W; w++) t[(size_t)w * c->n_chk + r] = code->enc_rows[(size_t)r * W + w];
      KML_CUDA(c, c->enc_t.alloc(t.size()));
      KML_CUDA(c, cudaMemcpy(c->enc_t.p, t.data(), t.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    std::vector<float2> pts(c->Q);
    for (int i = 0; i < c->Q; i++) pts[i] = make_float2((float)modem->points[2 * i], (float)modem->points[2 * i + 1]);
    KML_CUDA(c, c->points.alloc(c->Q));
    KML_CUDA(c, cudaMemcpy(c->points.p, pts.data(), sizeof(float2) * c->Q, cudaMemcpyHostToDevice));
    {  // quarter-turn symmetry of the constellation (demapper: one set of exponentials serves the four candidates)
      c->rot_symmetric = c->Q <= 64 ? 1 : 0;
      for (int cc = 1; cc <= 3 && c->rot_symmetric; cc++) {
        const double cr = cc == 2 ? -1.0 : 0.0, ci = cc == 1 ? 1.0 : (cc == 3 ? -1.0 : 0.0);  // e^{j cc pi/2}
        std::vector<char> used(c->Q, 0);
        for (int k = 0; k < c->Q && c->rot_symmetric; k++) {
          const double xr = modem->points[2 * k] * cr - modem->points[2 * k + 1] * ci;
          const double xi = modem->points[2 * k] * ci + modem->points[2 * k + 1] * cr;
          int hit = -1;
          for (int m = 0; m < c->Q; m++)
            if (!used[m] && std::hypot(modem->points[2 * m] - xr, modem->points[2 * m + 1] - xi) < 1e-9) { hit = m; break; }
          if (hit < 0) c->rot_symmetric = 0;
          else { used[hit] = 1; c->rot_perm[cc - 1][k] = hit; }
        }
      }
    }
    {  // 4 points: for every (candidate c, bit j) the pair of point indices whose probabilities add up to P(bit = 0)
       // (demap_symbol_q4's partition code); candidate c reads point perm_c[k] under label k
      c->q4_code = 0xFFFFFFFFu;
      const char *e4 = knob("KML_DEMAP_NO_Q4");  // shipped fallback (announced on stderr): the general demapper
      if (c->Q == 4 && !(e4 && atoi(e4))) {
        uint32_t code = 0;
        for (int cc = 0; cc < 4; cc++)
          for (int j = 0; j < 2; j++) {
            int z0[2], n0 = 0;
            for (int k = 0; k < 4; k++)
              if (((k >> (1 - j)) & 1) == 0) z0[n0++] = (cc == 0 || !c->rot_symmetric) ? k : c->rot_perm[cc - 1][k];
            const int a = std::min(z0[0], z0[1]), b = std::max(z0[0], z0[1]);
            int idx = 0;
            if (a == 0 && b == 1) idx = 0; else if (a == 2 && b == 3) idx = 1; else if (a == 0 && b == 2) idx = 2;
            else if (a == 1 && b == 3) idx = 3; else if (a == 0 && b == 3) idx = 4; else idx = 5;
            code |= (uint32_t)idx << (3 * (2 * cc + j));
          }
        c->q4_code = code;
      }
    }
    {  // 64 points on a square grid with the Gray labelling demap_symbol_grid64 has compiled in (link_kernels.cu): levels
       // ascending l_0..l_7 = -l_7..-l_0 on both axes, point k at (level i, level j) with label bits (MSB first)
       // i >= 4, i in 2..5, i in {1,2,5,6}, j < 4, j in 2..5, j in {1,2,5,6}.  Anything else keeps the general 64-point path.
      c->grid64 = 0;
      const char *e = knob("KML_DEMAP_NO_GRID");  // shipped fallback (announced on stderr): the quad-of-lanes demapper
      if (c->Q == 64 && c->rot_symmetric && !(e && atoi(e))) {
        std::vector<double> lv;
        for (int k = 0; k < 64; k++) {
          bool seen = false;
          for (double v : lv) seen = seen || std::fabs(v - modem->points[2 * k]) < 1e-9;
          if (!seen) lv.push_back(modem->points[2 * k]);
        }
        std::sort(lv.begin(), lv.end());
        bool ok = lv.size() == 8;
        for (int i = 0; ok && i < 8; i++) ok = std::fabs(lv[i] + lv[7 - i]) < 1e-9;
        auto level_of = [&](double v) {
          for (int i = 0; i < 8; i++)
            if (std::fabs(lv[i] - v) < 1e-9) return i;
          return -1;
        };
        for (int k = 0; ok && k < 64; k++) {
          const int i = level_of(modem->points[2 * k]), j = level_of(modem->points[2 * k + 1]);
          if (i < 0 || j < 0) { ok = false; break; }
          const int mid_i = (i >= 2 && i <= 5), low_i = (i == 1 || i == 2 || i == 5 || i == 6);
          const int mid_j = (j >= 2 && j <= 5), low_j = (j == 1 || j == 2 || j == 5 || j == 6);
          const int label = ((i >= 4) << 5) | (mid_i << 4) | (low_i << 3) | ((j < 4) << 2) | (mid_j << 1) | low_j;
          ok = label == k;
        }
        if (ok) {
          c->grid64 = 1;
          for (int i = 0; i < 8; i++) c->grid_levels[i] = (float)lv[i];
        }
      }
    }
    {  // 16 points on the 4 x 4 Gray grid demap_symbol_grid16 has compiled in: label bits (MSB first) j < 2, j in {1,2}, i < 2,
       // i in {1,2} for the point at (in-phase level i, quadrature level j)
      c->grid16 = 0;
      const char *e = knob("KML_DEMAP_NO_GRID");
      if (c->Q == 16 && c->rot_symmetric && !(e && atoi(e))) {
        std::vector<double> lv;
        for (int k = 0; k < 16; k++) {
          bool seen = false;
          for (double v : lv) seen = seen || std::fabs(v - modem->points[2 * k]) < 1e-9;
          if (!seen) lv.push_back(modem->points[2 * k]);
        }
        std::sort(lv.begin(), lv.end());
        bool ok = lv.size() == 4;
        for (int i = 0; ok && i < 4; i++) ok = std::fabs(lv[i] + lv[3 - i]) < 1e-9;
        auto level_of = [&](double v) {
          for (int i = 0; i < (int)lv.size(); i++)
            if (std::fabs(lv[i] - v) < 1e-9) return i;
          return -1;
        };
        for (int k = 0; ok && k < 16; k++) {
          const int i = level_of(modem->points[2 * k]), j = level_of(modem->points[2 * k + 1]);
          if (i < 0 || j < 0) { ok = false; break; }
          const int label = ((j < 2) << 3) | ((j == 1 || j == 2) << 2) | ((i < 2) << 1) | (i == 1 || i == 2);
          ok = label == k;
        }
        if (ok) {
          c->grid16 = 1;
          for (int i = 0; i < 4; i++) c->grid_levels[i] = (float)lv[i];
        }
      }
    }
    {  // k-means: only "nearest centroid is cluster 0" matters, decided by the Voronoi neighbours of s_0 (host_code.cpp)
      const std::vector<int> nb = voronoi_neighbours_of_first(modem->points, c->Q);
      const double s0r = modem->points[0], s0i = modem->points[1], s0n = s0r * s0r + s0i * s0i;
      c->km.s0r = s0r; c->km.s0i = s0i;
      c->km.is0r = s0r / s0n; c->km.is0i = -s0i / s0n;  // 1 / s_0 = conj(s_0) / |s_0|^2
      c->km.n_nb = nb.size() <= 8 ? (int)nb.size() : 0;
      for (int t = 0; t < c->km.n_nb; t++) {
        const double sr = modem->points[2 * nb[t]], si = modem->points[2 * nb[t] + 1];
        c->km.dsr[t] = sr - s0r; c->km.dsi[t] = si - s0i;
        c->km.dn[t] = 0.5 * ((sr * sr + si * si) - s0n);
      }
    }
    KML_CUDA(c, c->row_ptr.alloc(c->M + 1));
    KML_CUDA(c, cudaMemcpy(c->row_ptr.p, code->row_ptr, sizeof(int32_t) * (c->M + 1), cudaMemcpyHostToDevice));
    KML_CUDA(c, c->col_idx.alloc(c->E));
    KML_CUDA(c, cudaMemcpy(c->col_idx.p, code->col_idx, sizeof(int32_t) * c->E, cudaMemcpyHostToDevice));
    {  // ELL copy of the graph for the demapper's syndrome pass: [max row degree][M] uint16, padding → an always-zero byte
      int w = 0;
      for (int r = 0; r < c->M; r++) w = std::max(w, code->row_ptr[r + 1] - code->row_ptr[r]);
      std::vector<uint16_t> ell((size_t)w * c->M, (uint16_t)(c->N));
      for (int r = 0; r < c->M; r++)
        for (int e = code->row_ptr[r], k = 0; e < code->row_ptr[r + 1]; e++, k++) ell[(size_t)k * c->M + r] = (uint16_t)code->col_idx[e];
      c->ell_width = w;
      KML_CUDA(c, c->col_ell.alloc(ell.size()));
      KML_CUDA(c, cudaMemcpy(c->col_ell.p, ell.data(), ell.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
    KML_CUDA(c, c->counters.alloc(5));
    KML_CUDA(c, cudaMemset(c->counters.p, 0, 5 * sizeof(unsigned long long)));
    KML_CUDA(c, cudaMallocHost(&c->h_counters, 5 * sizeof(unsigned long long)));
    KML_CUDA(c, cudaMallocHost(&c->h_chain_counts, 2 * sizeof(int32_t)));
    // syndrom_soft_ before the first Decoder call: the reference leaves the array uninitialised
    // (binaryldpccodec.cc:88); here, and in oracle/ref/ref_harness.cc, it starts as all ones → ln-sum 0
    KML_CUDA(c, c->soft_carry.alloc(1));
    KML_CUDA(c, cudaMemset(c->soft_carry.p, 0, sizeof(double)));
    return KML_OK;
  };
  KML_TRY(upload());
  KML_TRY(build_decoder_tables(c, code));
  for (int k = 0; k < kml_ctx::kLanes; k++) KML_TRY(alloc_lane(c, c->lane[k]));
#undef KML_TRY
  *out = c;
  return KML_OK;
}

extern "C" void kml_destroy(kml_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int k = 0; k < kml_ctx::kLanes; k++) free_lane(c->lane[k]);
  c->enc_t.release(); c->points.release(); c->row_ptr.release(); c->col_idx.release();
  c->vn_addr.release(); c->vn_addr_rm.release(); c->vn_addr_g.release(); c->vn_items.release(); c->cn_items.release(); c->cn_deg_rm.release(); c->col_ell.release(); c->vn_deg.release(); c->cn_deg.release(); c->counters.release();
  c->lay_ptr.release(); c->lay_cs.release();
  if (c->h_counters) cudaFreeHost(c->h_counters);
  if (c->h_chain_counts) cudaFreeHost(c->h_chain_counts);
  for (auto &slot : c->rx_done)
    for (cudaEvent_t e : slot)
      if (e) cudaEventDestroy(e);
  c->soft_carry.release();
  delete c;
}

extern "C" const char *kml_last_error(const kml_ctx *c) { return c ? c->err.c_str() : global_error(); }

extern "C" int kml_set_early_exit(kml_ctx *c, int early_exit) {
  if (!c) return KML_ERR_ARG;
  c->opts.early_exit = early_exit ? 1 : 0;
  return KML_OK;
}

extern "C" int kml_set_algorithm(kml_ctx *c, int algorithm, double alpha) {
  if (!c) return KML_ERR_ARG;
  if (algorithm < 0 || algorithm > 3)
    return fail_arg(c, "kml_set_algorithm: 0 = sum-product, 1 = normalised min-sum fp32, 2 = min-sum fp16 x 2 frames, "
                       "3 = layered min-sum (quasi-cyclic codes)");
  if (algorithm == 3 && !c->lay.z) return fail_arg(c, ("kml_set_algorithm: layered min-sum: " + c->lay_why).c_str());
  if (algorithm != 0 && c->opts.metric_type) return fail_arg(c, "kml_set_algorithm: soft-syndrome metric needs sum-product");
  if (algorithm != 0 && !(alpha > 0.0 && alpha <= 1.0)) return fail_arg(c, "kml_set_algorithm: alpha must be in (0, 1]");
  c->opts.algorithm = algorithm;
  if (algorithm != 0) { c->alpha = (float)alpha; c->beta = 0.0f; }
  c->dl = c->dl_alg[algorithm];
  return KML_OK;
}

extern "C" int kml_set_minsum(kml_ctx *c, double alpha, double beta) {
  if (!c) return KML_ERR_ARG;
  if (!(alpha > 0.0 && alpha <= 1.0) || !(beta >= 0.0 && beta < 27.0))
    return fail_arg(c, "kml_set_minsum: alpha must be in (0, 1], beta in [0, 27)");
  c->alpha = (float)alpha;
  c->beta = (float)beta;
  return KML_OK;
}

extern "C" int kml_info(const kml_ctx *c, int32_t info[8]) {
  if (!c || !info) return KML_ERR_ARG;
  info[0] = c->M; info[1] = c->N; info[2] = c->n_tx; info[3] = c->K;
  info[4] = c->bits; info[5] = c->Q; info[6] = c->n_sym; info[7] = c->max_batch;
  return KML_OK;
}

extern "C" int kml_measure_smem_bandwidth(kml_ctx *c, double *gb_per_s) {
  if (!c || !gb_per_s) return KML_ERR_ARG;
  KML_CUDA(c, cudaSetDevice(c->device));
  KML_CUDA(c, measure_smem_bandwidth(c->num_sms, gb_per_s, c->lane[0].stream));
  c->launches += 3;
  return KML_OK;
}

extern "C" int kml_decoder_info(const kml_ctx *c, int32_t info[8]) {
  if (!c || !info) return KML_ERR_ARG;
  info[0] = (int)c->dl.kind | (c->dl.qc_plan << 8) | (c->dl.rowmajor << 16); info[1] = c->dl.threads; info[2] = c->dl.smem_bytes; info[3] = c->dl.ctas_per_sm;
  info[4] = c->layout_residual; info[5] = c->layout_excess; info[6] = ((c->N + 31) / 32) * c->dt.dv_max; info[7] = c->dt.m_pad;
  return KML_OK;
}

extern "C" uint64_t kml_launch_count(const kml_ctx *c) { return c ? c->launches : 0; }

// ================================================================================================ stage entry points
// Host-pointer stages run batch by batch on lane 0; temporaries for the int32 <-> packed conversions are (re)allocated
// on demand.  They exist for parity tests and standalone measurements, not for peak throughput.
namespace {
// begin / end of an entry point that uses lane `l` on stream `s`
#define KML_ENTER(c, l, s)                      \
  KML_CUDA(c, cudaSetDevice((c)->device));      \
  KML_RC(lane_acquire(c, l, s))
#define KML_LEAVE(c, l, s) KML_RC(lane_release(c, l, s))
}  // namespace

extern "C" int kml_encode(kml_ctx *c, int B, const int32_t *u, int32_t *cw) {
  KML_RC(check_batch(c, B));
  if (!u || !cw) return fail_arg(c, "kml_encode: null buffer");
  Lane &l = c->lane[0];
  KML_ENTER(c, l, l.stream);
  for (int b0 = 0; b0 < B; b0 += c->max_batch) {
    const int nb = std::min(c->max_batch, B - b0);
    KML_RC(ensure(c, l.bits_io, (size_t)nb * std::max(c->K, c->n_tx)));
    KML_CUDA(c, cudaMemcpyAsync(l.bits_io.p, u + (size_t)b0 * c->K, sizeof(int32_t) * nb * c->K, cudaMemcpyHostToDevice, l.stream));
    KML_LAUNCH(c, launch_pack_bits(nb, c->K, l.bits_io.p, l.u_packed.p, l.stream));
    GenParams g = gen_params(c, nb, 0.0, 0, 0);
    KML_LAUNCH(c, launch_encode(g, l.u_packed.p, l.c_packed.p, l.stream));
    KML_LAUNCH(c, launch_unpack_bits(nb, c->n_tx, 0, c->tx_words, l.c_packed.p, l.bits_io.p, l.stream));
    KML_CUDA(c, cudaMemcpyAsync(cw + (size_t)b0 * c->n_tx, l.bits_io.p, sizeof(int32_t) * nb * c->n_tx, cudaMemcpyDeviceToHost, l.stream));
    KML_CUDA(c, cudaStreamSynchronize(l.stream));
  }
  KML_LEAVE(c, l, l.stream);
  return KML_OK;
}

extern "C" int kml_generate(kml_ctx *c, int B, double snr_db, uint64_t seed, uint64_t frame0, int32_t *u, int32_t *cw,
                            float *h, float *y) {
  KML_RC(check_batch(c, B));
  Lane &l = c->lane[0];
  KML_ENTER(c, l, l.stream);
  for (int b0 = 0; b0 < B; b0 += c->max_batch) {
    const int nb = std::min(c->max_batch, B - b0);
    GenParams g = gen_params(c, nb, snr_db, seed, frame0 + b0);
    KML_LAUNCH(c, launch_gen_bits(g, l.u_packed.p, l.stream));
    KML_LAUNCH(c, launch_encode(g, l.u_packed.p, l.c_packed.p, l.stream));
    KML_LAUNCH(c, launch_channel(g, l.c_packed.p, nullptr, nullptr, l.h.p, l.y.p, l.stream));
    KML_RC(ensure(c, l.bits_io, (size_t)nb * std::max(c->K, c->n_tx)));
    if (u) {
      KML_LAUNCH(c, launch_unpack_bits(nb, c->K, 0, c->k_words, l.u_packed.p, l.bits_io.p, l.stream));
      KML_CUDA(c, cudaMemcpyAsync(u + (size_t)b0 * c->K, l.bits_io.p, sizeof(int32_t) * nb * c->K, cudaMemcpyDeviceToHost, l.stream));
      KML_CUDA(c, cudaStreamSynchronize(l.stream));
    }
    if (cw) {
      KML_LAUNCH(c, launch_unpack_bits(nb, c->n_tx, 0, c->tx_words, l.c_packed.p, l.bits_io.p, l.stream));
      KML_CUDA(c, cudaMemcpyAsync(cw + (size_t)b0 * c->n_tx, l.bits_io.p, sizeof(int32_t) * nb * c->n_tx, cudaMemcpyDeviceToHost, l.stream));
    }
    if (h) KML_CUDA(c, cudaMemcpyAsync(h + (size_t)b0 * 2, l.h.p, sizeof(float2) * nb, cudaMemcpyDeviceToHost, l.stream));
    if (y) KML_CUDA(c, cudaMemcpyAsync(y + (size_t)b0 * c->n_sym * 2, l.y.p, sizeof(float2) * nb * c->n_sym, cudaMemcpyDeviceToHost, l.stream));
    KML_CUDA(c, cudaStreamSynchronize(l.stream));
  }
  KML_LEAVE(c, l, l.stream);
  return KML_OK;
}

extern "C" int kml_modulate(kml_ctx *c, int B, const int32_t *cw, const float *h, const float *noise, double sigma, float *y) {
  KML_RC(check_batch(c, B));
  if (!cw || !h || !y) return fail_arg(c, "kml_modulate: null buffer");
  Lane &l = c->lane[0];
  KML_ENTER(c, l, l.stream);
  for (int b0 = 0; b0 < B; b0 += c->max_batch) {
    const int nb = std::min(c->max_batch, B - b0);
    KML_RC(ensure(c, l.bits_io, (size_t)nb * std::max(c->K, c->n_tx)));
    KML_RC(ensure(c, l.noise, (size_t)nb * c->n_sym));
    KML_CUDA(c, cudaMemcpyAsync(l.bits_io.p, cw + (size_t)b0 * c->n_tx, sizeof(int32_t) * nb * c->n_tx, cudaMemcpyHostToDevice, l.stream));
    KML_LAUNCH(c, launch_pack_bits(nb, c->n_tx, l.bits_io.p, l.c_packed.p, l.stream));
    KML_CUDA(c, cudaMemcpyAsync(l.h.p, h + (size_t)b0 * 2, sizeof(float2) * nb, cudaMemcpyHostToDevice, l.stream));
    if (noise) KML_CUDA(c, cudaMemcpyAsync(l.noise.p, noise + (size_t)b0 * c->n_sym * 2, sizeof(float2) * nb * c->n_sym, cudaMemcpyHostToDevice, l.stream));
    else KML_CUDA(c, cudaMemsetAsync(l.noise.p, 0, sizeof(float2) * nb * c->n_sym, l.stream));
    GenParams g = gen_params(c, nb, 0.0, 0, 0);
    g.sigma_over_sqrt2 = (float)(sigma / 1.4142135623730950488016);
    KML_LAUNCH(c, launch_channel(g, l.c_packed.p, l.h.p, l.noise.p, nullptr, l.y.p, l.stream));
    KML_CUDA(c, cudaMemcpyAsync(y + (size_t)b0 * c->n_sym * 2, l.y.p, sizeof(float2) * nb * c->n_sym, cudaMemcpyDeviceToHost, l.stream));
    KML_CUDA(c, cudaStreamSynchronize(l.stream));
  }
  KML_LEAVE(c, l, l.stream);
  return KML_OK;
}

namespace {
// y: float [B][n_sym][2] or, with y_is_f64, double (the reference's std::complex<double> received symbols)
int kmeans_host(kml_ctx *c, int B, const void *y, int y_is_f64, float *hhat, double *hhat64, int32_t *passes) {
  Lane &l = c->lane[0];
  KML_ENTER(c, l, l.stream);
  const size_t ysz = y_is_f64 ? sizeof(double2) : sizeof(float2);
  for (int b0 = 0; b0 < B; b0 += c->max_batch) {
    const int nb = std::min(c->max_batch, B - b0);
    void *ydev = l.y.p;
    if (y_is_f64) {
      KML_RC(ensure(c, l.y64, (size_t)c->max_batch * c->n_sym));
      ydev = l.y64.p;
    }
    if (hhat64) KML_RC(ensure(c, l.hhat64, (size_t)c->max_batch));
    KML_CUDA(c, cudaMemcpyAsync(ydev, (const char *)y + (size_t)b0 * c->n_sym * ysz, ysz * nb * c->n_sym, cudaMemcpyHostToDevice, l.stream));
    KML_LAUNCH(c, launch_kmeans(nb, ydev, y_is_f64, c->n_sym, c->points.p, c->Q, c->km, c->opts.kmeans_iter, l.hhat.p,
                               hhat64 ? l.hhat64.p : nullptr, l.passes.p, y_is_f64 ? l.y.p : nullptr, c->num_sms, l.stream));
    if (hhat) KML_CUDA(c, cudaMemcpyAsync(hhat + (size_t)b0 * 2, l.hhat.p, sizeof(float2) * nb, cudaMemcpyDeviceToHost, l.stream));
    if (hhat64) KML_CUDA(c, cudaMemcpyAsync(hhat64 + (size_t)b0 * 2, l.hhat64.p, sizeof(double2) * nb, cudaMemcpyDeviceToHost, l.stream));
    if (passes) KML_CUDA(c, cudaMemcpyAsync(passes + b0, l.passes.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, l.stream));
    KML_CUDA(c, cudaStreamSynchronize(l.stream));
  }
  KML_LEAVE(c, l, l.stream);
  return KML_OK;
}
}  // namespace

extern "C" int kml_kmeans(kml_ctx *c, int B, const float *y, float *hhat, int32_t *passes) {
  KML_RC(check_batch(c, B));
  if (!y || !hhat) return fail_arg(c, "kml_kmeans: null buffer");
  return kmeans_host(c, B, y, 0, hhat, nullptr, passes);
}

extern "C" int kml_kmeans_f64(kml_ctx *c, int B, const double *y, double *hhat, int32_t *passes) {
  KML_RC(check_batch(c, B));
  if (!y || !hhat) return fail_arg(c, "kml_kmeans_f64: null buffer");
  return kmeans_host(c, B, y, 1, nullptr, hhat, passes);
}

namespace {
DemapParams demap_params(kml_ctx *c, Lane &l, int B, double var, int n_cand, int hard_metric) {
  DemapParams d{};
  d.B = B; d.n_sym = c->n_sym; d.n_tx = c->n_tx; d.bits_per_symbol = c->bits; d.q = c->Q; d.n_cand = n_cand;
  d.hard_metric = hard_metric; d.m_rows = c->M; d.punct = c->punct; d.y = l.y.p; d.h = l.hhat.p;
  d.inv_var = (float)(1.0 / var);
  for (int k = 0; k < 4; k++) d.rot[k] = c->rot[k];
  d.points = c->points.p; d.row_ptr = c->row_ptr.p; d.col_idx = c->col_idx.p; d.col_ell = c->col_ell.p; d.ell_width = c->ell_width;
  d.lr = l.lr.p; d.metric = l.metric.p; d.kstar = l.kstar.p;
  d.symmetric = c->rot_symmetric;
  std::memcpy(d.perm, c->rot_perm, sizeof d.perm);
  d.grid64 = c->grid64; d.grid16 = c->grid16;
  std::memcpy(d.levels, c->grid_levels, sizeof d.levels);
  d.q4_code = c->q4_code;
  return d;
}

// GetMetrics on the four candidates of nb frames whose y / hhat sit in l.y / l.hhat (kmcodec.cc:122-139): l.metric, l.kstar
int resolve_on_lane(kml_ctx *c, Lane &l, cudaStream_t s, int nb, double var) {
  const bool decode_metric = c->is_5g || c->opts.metric_type;
  DemapParams d = demap_params(c, l, nb, var, 4, decode_metric ? 0 : 1);
  KML_LAUNCH(c, launch_demap(d, c->num_sms, s));
  if (decode_metric) {
    KML_RC(metric_decodes(c, l, s, nb));
    if (c->opts.metric_type) KML_RC(soft_chain(c, l, s, nb, false));
  }
  return KML_OK;
}
}  // namespace

extern "C" int kml_demap(kml_ctx *c, int B, const float *y, const float *h, double var, float *llr) {
  KML_RC(check_batch(c, B));
  if (!y || !h || !llr || !(var > 0)) return fail_arg(c, "kml_demap: bad argument");
  Lane &l = c->lane[0];
  KML_ENTER(c, l, l.stream);
  for (int b0 = 0; b0 < B; b0 += c->max_batch) {
    const int nb = std::min(c->max_batch, B - b0);
    KML_CUDA(c, cudaMemcpyAsync(l.y.p, y + (size_t)b0 * c->n_sym * 2, sizeof(float2) * nb * c->n_sym, cudaMemcpyHostToDevice, l.stream));
    KML_CUDA(c, cudaMemcpyAsync(l.hhat.p, h + (size_t)b0 * 2, sizeof(float2) * nb, cudaMemcpyHostToDevice, l.stream));
    DemapParams d = demap_params(c, l, nb, var, 1, 0);
    KML_LAUNCH(c, launch_demap(d, c->num_sms, l.stream));
    KML_LAUNCH(c, launch_lr_to_llr((size_t)nb * c->n_tx, l.lr.p, l.lr.p, l.stream));
    KML_CUDA(c, cudaMemcpyAsync(llr + (size_t)b0 * c->n_tx, l.lr.p, sizeof(float) * nb * c->n_tx, cudaMemcpyDeviceToHost, l.stream));
    KML_CUDA(c, cudaStreamSynchronize(l.stream));
  }
  KML_LEAVE(c, l, l.stream);
  return KML_OK;
}

extern "C" int kml_resolve(kml_ctx *c, int B, const float *y, const float *hhat, double var, float *metric, int32_t *kstar) {
  KML_RC(check_batch(c, B));
  if (!y || !hhat || !(var > 0)) return fail_arg(c, "kml_resolve: bad argument");
  Lane &l = c->lane[0];
  cudaStream_t s = l.stream;
  KML_ENTER(c, l, s);
  for (int b0 = 0; b0 < B; b0 += c->max_batch) {
    const int nb = std::min(c->max_batch, B - b0);
    KML_CUDA(c, cudaMemcpyAsync(l.y.p, y + (size_t)b0 * c->n_sym * 2, sizeof(float2) * nb * c->n_sym, cudaMemcpyHostToDevice, s));
    KML_CUDA(c, cudaMemcpyAsync(l.hhat.p, hhat + (size_t)b0 * 2, sizeof(float2) * nb, cudaMemcpyHostToDevice, s));
    KML_RC(resolve_on_lane(c, l, s, nb, var));
    if (metric) KML_CUDA(c, cudaMemcpyAsync(metric + (size_t)b0 * 4, l.metric.p, sizeof(float) * 4 * nb, cudaMemcpyDeviceToHost, s));
    if (kstar) KML_CUDA(c, cudaMemcpyAsync(kstar + b0, l.kstar.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, s));
    KML_CUDA(c, cudaStreamSynchronize(s));
  }
  KML_LEAVE(c, l, s);
  return KML_OK;
}

namespace {
// in: float natural-log LLR [B][n_tx] (p0_is_f64 = 0) or the reference's own decoder input, double P(bit = 0) [B][n_tx]
int decode_host(kml_ctx *c, int B, const void *in, int p0_is_f64, int iter_count, int32_t *cc_hat, int32_t *uu_hat, int32_t *ret) {
  Lane &l = c->lane[0];
  cudaStream_t s = l.stream;
  KML_ENTER(c, l, s);
  for (int b0 = 0; b0 < B; b0 += c->max_batch) {
    const int nb = std::min(c->max_batch, B - b0);
    const size_t n = (size_t)nb * c->n_tx;
    if (p0_is_f64) {  // ratio P0 / (1 - P0) formed in fp64 on the device, then narrowed: the decoder's own input format
      KML_RC(ensure(c, l.p0_io, (size_t)c->max_batch * c->n_tx));
      KML_CUDA(c, cudaMemcpyAsync(l.p0_io.p, (const double *)in + (size_t)b0 * c->n_tx, sizeof(double) * n, cudaMemcpyHostToDevice, s));
      KML_LAUNCH(c, launch_p0_to_lr(n, l.p0_io.p, l.lr.p, s));
    } else {
      KML_CUDA(c, cudaMemcpyAsync(l.lr.p, (const float *)in + (size_t)b0 * c->n_tx, sizeof(float) * n, cudaMemcpyHostToDevice, s));
    }
    DecParams p = dec_params(c, l, nb, l.lr.p, nullptr, 1, p0_is_f64 ? 1 : 0, iter_count, l.cc_hat_packed.p, l.ret.p, nullptr);
    KML_LAUNCH(c, dec_launch(c->dl, p, c->num_sms, s));
    KML_RC(ensure(c, l.bits_io, (size_t)nb * c->N));
    if (cc_hat) {
      KML_LAUNCH(c, launch_unpack_bits(nb, c->N, 0, c->words_n, l.cc_hat_packed.p, l.bits_io.p, s));
      KML_CUDA(c, cudaMemcpyAsync(cc_hat + (size_t)b0 * c->N, l.bits_io.p, sizeof(int32_t) * nb * c->N, cudaMemcpyDeviceToHost, s));
      KML_CUDA(c, cudaStreamSynchronize(s));
    }
    if (uu_hat) {
      KML_LAUNCH(c, launch_unpack_bits(nb, c->K, c->info_offset, c->words_n, l.cc_hat_packed.p, l.bits_io.p, s));
      KML_CUDA(c, cudaMemcpyAsync(uu_hat + (size_t)b0 * c->K, l.bits_io.p, sizeof(int32_t) * nb * c->K, cudaMemcpyDeviceToHost, s));
    }
    if (ret) KML_CUDA(c, cudaMemcpyAsync(ret + b0, l.ret.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, s));
    KML_CUDA(c, cudaStreamSynchronize(s));
  }
  KML_LEAVE(c, l, s);
  return KML_OK;
}
}  // namespace

extern "C" int kml_decode(kml_ctx *c, int B, const float *llr, int iter_count, int32_t *cc_hat, int32_t *uu_hat, int32_t *ret) {
  KML_RC(check_batch(c, B));
  if (!llr || iter_count < 1) return fail_arg(c, "kml_decode: bad argument");
  return decode_host(c, B, llr, 0, iter_count, cc_hat, uu_hat, ret);
}

extern "C" int kml_decode_p0(kml_ctx *c, int B, const double *p0, int iter_count, int32_t *cc_hat, int32_t *uu_hat, int32_t *ret) {
  KML_RC(check_batch(c, B));
  if (!p0 || iter_count < 1) return fail_arg(c, "kml_decode_p0: bad argument");
  return decode_host(c, B, p0, 1, iter_count, cc_hat, uu_hat, ret);
}

namespace {
// Host-buffer receiver: batches alternate between the two lanes so the H2D copy of batch i+1 and the D2H copy of batch
// i-1 overlap the kernels of batch i (true overlap needs pinned host buffers; pageable ones still work).
int receive_host(kml_ctx *c, int B, const void *y, int y_is_f64, const void *true_h, double var, uint32_t *uu_hat_packed,
                 float *hhat, double *hhat64, int32_t *kstar, int32_t *ret, float *metric, bool wait = true) {
  KML_CUDA(c, cudaSetDevice(c->device));
  // sub-batches of ~2048 frames (measured best on B200: the first H2D and the last D2H are the only exposed copies,
  // and the other lane's kernels fill the tail of each decoder launch); never fewer than ~1 frame per resident CTA
  int step = c->max_batch;
  if (B > 2 * c->num_sms * 8) step = std::min(step, std::max(c->num_sms * 8, std::min(2048, (B + 1) / 2)));
  int slow = 2;
  if (!wait) {  // pipelined (kml_receive_submit): the previous batch hides this one's first copy — two halves, one per lane,
    step = std::min(c->max_batch, std::max(c->num_sms * 8, (B + 1) / 2));  // no slow start (measured on B200, 16384 frames
    slow = 0;                                                              // per batch: 5.01 ms per step against 5.43 ms)
  }
#ifdef KML_TUNING
  if (const char *e = tuning_knob("KML_RX_CHUNK")) step = std::max(1, std::min(c->max_batch, atoi(e)));
  if (const char *e = tuning_knob("KML_RX_SLOW")) slow = std::max(0, std::min(6, atoi(e)));
#endif
  // the soft metric's syndrom_soft_ chain runs through the frames in order: one lane, one batch at a time
  const bool sequential = c->opts.metric_type && !c->opts.known_h;
  const size_t ysz = y_is_f64 ? sizeof(double2) : sizeof(float2), hsz = y_is_f64 ? sizeof(double2) : sizeof(float2);
  int li = 0, chunk_no = 0;
  for (int b0 = 0, nb = 0; b0 < B; b0 += nb, li ^= sequential ? 0 : 1, chunk_no++) {
    Lane &l = c->lane[li];
    cudaStream_t s = l.stream;
    KML_RC(lane_acquire(c, l, s));
    // slow start: the first copy is the only one nothing can hide, so the first two sub-batches are a quarter / a half
    const int want = (B > 4 * step && chunk_no < slow) ? std::max(c->num_sms, step >> (slow - chunk_no)) : step;
    nb = std::min(want, B - b0);
    void *ydev = l.y.p;
    if (y_is_f64) {
      KML_RC(ensure(c, l.y64, (size_t)c->max_batch * c->n_sym));
      ydev = l.y64.p;
      if (hhat64) KML_RC(ensure(c, l.hhat64, (size_t)c->max_batch));
    }
    KML_CUDA(c, cudaMemcpyAsync(ydev, (const char *)y + (size_t)b0 * c->n_sym * ysz, ysz * nb * c->n_sym, cudaMemcpyHostToDevice, s));
    if (c->opts.known_h) {
      if (y_is_f64) {
        KML_RC(ensure(c, l.h64, (size_t)c->max_batch));
        KML_CUDA(c, cudaMemcpyAsync(l.h64.p, (const char *)true_h + (size_t)b0 * hsz, hsz * nb, cudaMemcpyHostToDevice, s));
        KML_LAUNCH(c, launch_f64_to_f32((size_t)nb * 2, reinterpret_cast<const double *>(l.h64.p), reinterpret_cast<float *>(l.h.p), s));
      } else {
        KML_CUDA(c, cudaMemcpyAsync(l.h.p, (const char *)true_h + (size_t)b0 * hsz, hsz * nb, cudaMemcpyHostToDevice, s));
      }
    }
    KML_RC(receive_on_lane(c, l, s, nb, ydev, y_is_f64, l.h.p, var, hhat64 != nullptr));
    if (uu_hat_packed)
      KML_CUDA(c, cudaMemcpyAsync(uu_hat_packed + (size_t)b0 * c->k_words, l.uu_hat_packed.p, sizeof(uint32_t) * nb * c->k_words, cudaMemcpyDeviceToHost, s));
    if (hhat && !c->opts.known_h)
      KML_CUDA(c, cudaMemcpyAsync(hhat + (size_t)b0 * 2, l.hhat.p, sizeof(float2) * nb, cudaMemcpyDeviceToHost, s));
    if (hhat64 && !c->opts.known_h)
      KML_CUDA(c, cudaMemcpyAsync(hhat64 + (size_t)b0 * 2, l.hhat64.p, sizeof(double2) * nb, cudaMemcpyDeviceToHost, s));
    if (kstar && !c->opts.known_h)
      KML_CUDA(c, cudaMemcpyAsync(kstar + b0, l.kstar.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, s));
    if (ret) KML_CUDA(c, cudaMemcpyAsync(ret + b0, l.ret.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, s));
    if (metric && !c->opts.known_h)
      KML_CUDA(c, cudaMemcpyAsync(metric + (size_t)b0 * 4, l.metric.p, sizeof(float) * 4 * nb, cudaMemcpyDeviceToHost, s));
    KML_RC(lane_release(c, l, s));
    if (sequential) KML_CUDA(c, cudaStreamSynchronize(s));
  }
  if (!wait) return KML_OK;
  KML_CUDA(c, cudaStreamSynchronize(c->lane[0].stream));
  KML_CUDA(c, cudaStreamSynchronize(c->lane[1].stream));
  return KML_OK;
}

// completion of the oldest outstanding kml_receive_submit
int rx_retire_one(kml_ctx *c) {
  const int slot = (int)(c->rx_waited % kml_ctx::kRxRing);
  for (int k = 0; k < 2; k++) KML_CUDA(c, cudaEventSynchronize(c->rx_done[slot][k]));
  c->rx_waited++;
  return KML_OK;
}
}  // namespace

extern "C" int kml_receive_submit(kml_ctx *c, int B, const float *y, const float *true_h, double var, uint32_t *uu_hat_packed,
                                  float *hhat, int32_t *kstar, int32_t *ret, float *metric) {
  KML_RC(check_batch(c, B));
  if (!y || !(var > 0) || (c->opts.known_h && !true_h)) return fail_arg(c, "kml_receive_submit: bad argument");
  KML_CUDA(c, cudaSetDevice(c->device));
  while (c->rx_submitted - c->rx_waited >= (uint64_t)kml_ctx::kRxRing) KML_RC(rx_retire_one(c));  // ring full: oldest first
  KML_RC(receive_host(c, B, y, 0, true_h, var, uu_hat_packed, hhat, nullptr, kstar, ret, metric, false));
  const int slot = (int)(c->rx_submitted % kml_ctx::kRxRing);
  for (int k = 0; k < 2; k++) {
    if (!c->rx_done[slot][k]) KML_CUDA(c, cudaEventCreateWithFlags(&c->rx_done[slot][k], cudaEventDisableTiming));
    KML_CUDA(c, cudaEventRecord(c->rx_done[slot][k], c->lane[k].stream));
  }
  c->rx_submitted++;
  return KML_OK;
}

extern "C" int kml_receive_wait(kml_ctx *c, int max_outstanding) {
  if (!c) return KML_ERR_ARG;
  if (max_outstanding < 0) max_outstanding = 0;
  KML_CUDA(c, cudaSetDevice(c->device));
  while (c->rx_submitted - c->rx_waited > (uint64_t)max_outstanding) KML_RC(rx_retire_one(c));
  return KML_OK;
}

extern "C" int kml_receive(kml_ctx *c, int B, const float *y, const float *true_h, double var, uint32_t *uu_hat_packed,
                           float *hhat, int32_t *kstar, int32_t *ret, float *metric) {
  KML_RC(check_batch(c, B));
  if (!y || !(var > 0) || (c->opts.known_h && !true_h)) return fail_arg(c, "kml_receive: bad argument");
  return receive_host(c, B, y, 0, true_h, var, uu_hat_packed, hhat, nullptr, kstar, ret, metric);
}

extern "C" int kml_receive_f64(kml_ctx *c, int B, const double *y, const double *true_h, double var, uint32_t *uu_hat_packed,
                               double *hhat, int32_t *kstar, int32_t *ret, float *metric) {
  KML_RC(check_batch(c, B));
  if (!y || !(var > 0) || (c->opts.known_h && !true_h)) return fail_arg(c, "kml_receive_f64: bad argument");
  return receive_host(c, B, y, 1, true_h, var, uu_hat_packed, nullptr, hhat, kstar, ret, metric);
}

extern "C" int kml_soft_syndrome_state(kml_ctx *c, int set, double *value) {
  if (!c || !value) return KML_ERR_ARG;
  KML_CUDA(c, cudaSetDevice(c->device));
  KML_CUDA(c, cudaDeviceSynchronize());
  if (set) KML_CUDA(c, cudaMemcpy(c->soft_carry.p, value, sizeof(double), cudaMemcpyHostToDevice));
  else KML_CUDA(c, cudaMemcpy(value, c->soft_carry.p, sizeof(double), cudaMemcpyDeviceToHost));
  return KML_OK;
}

extern "C" int kml_count_errors(kml_ctx *c, int B, const uint32_t *u_packed, const uint32_t *uu_hat_packed, uint64_t counters[4]) {
  KML_RC(check_batch(c, B));
  if (!u_packed || !uu_hat_packed || !counters) return fail_arg(c, "kml_count_errors: null buffer");
  Lane &l = c->lane[0];
  cudaStream_t s = l.stream;
  KML_ENTER(c, l, s);
  KML_CUDA(c, cudaMemsetAsync(c->counters.p, 0, 5 * sizeof(unsigned long long), s));
  for (int b0 = 0; b0 < B; b0 += c->max_batch) {
    const int nb = std::min(c->max_batch, B - b0);
    KML_CUDA(c, cudaMemcpyAsync(l.u_packed.p, u_packed + (size_t)b0 * c->k_words, sizeof(uint32_t) * nb * c->k_words, cudaMemcpyHostToDevice, s));
    KML_CUDA(c, cudaMemcpyAsync(l.uu_hat_packed.p, uu_hat_packed + (size_t)b0 * c->k_words, sizeof(uint32_t) * nb * c->k_words, cudaMemcpyHostToDevice, s));
    KML_LAUNCH(c, launch_count_errors(nb, c->K, c->k_words, l.u_packed.p, l.uu_hat_packed.p, nullptr, c->opts.max_iter, c->counters.p, s));
  }
  KML_CUDA(c, cudaMemcpyAsync(c->h_counters, c->counters.p, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  KML_CUDA(c, cudaStreamSynchronize(s));
  KML_LEAVE(c, l, s);
  for (int k = 0; k < 4; k++) counters[k] += c->h_counters[k];
  return KML_OK;
}

// ================================================================================================ fused path
// Each lane accumulates into its OWN device counters, read back only when that lane is idle: the stop rule then depends
// on completed batches alone (same seed → same stopping point), and nothing reads a counter another stream is adding to.
extern "C" int kml_simulate(kml_ctx *c, double snr_db, uint64_t seed, uint64_t frame_begin, uint64_t frame_count,
                            uint64_t max_err_blk, uint64_t counters[4], uint64_t *iters_sum) {
  if (!c || !counters) return KML_ERR_ARG;
  KML_CUDA(c, cudaSetDevice(c->device));
  const double var = std::pow(10.0, -0.1 * snr_db);
  const bool sequential = c->opts.metric_type && !c->opts.known_h;  // see receive_host
  // four lanes keep a few more kernels in flight than two (C1 at 15 dB 20.7 -> 21.2 M frames/s, C5 8.17 -> 8.07 s); a tight error
  // budget keeps two, so that the stop rule lags by one batch as before
  constexpr int NLmax = kml_ctx::kLanes;
  const int NL = (max_err_blk != 0 && max_err_blk < 4 * (uint64_t)c->max_batch) ? 2 : NLmax;
  for (int li = 0; li < NL; li++) {
    Lane &l = c->lane[li];
    KML_RC(lane_acquire(c, l, l.stream));
    KML_CUDA(c, cudaMemsetAsync(l.counters.p, 0, 5 * sizeof(unsigned long long), l.stream));
  }
  unsigned long long seen[NLmax][5] = {};  // each lane's counters as of its last completed batch
  uint64_t done = 0;
  int li = 0;
  bool pending[NLmax] = {};
  auto read_lane = [&](int k) -> int {  // lane k is idle: its counters are final for everything it was given
    Lane &l = c->lane[k];
    KML_CUDA(c, cudaMemcpyAsync(c->h_counters, l.counters.p, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, l.stream));
    KML_CUDA(c, cudaStreamSynchronize(l.stream));
    for (int j = 0; j < 5; j++) seen[k][j] = c->h_counters[j];
    return KML_OK;
  };
  auto run = [&]() -> int {
    while (done < frame_count) {
      Lane &l = c->lane[li];
      if (pending[li]) {  // the lane's previous batch must be finished before its buffers are reused
        KML_RC(read_lane(li));
        pending[li] = false;
        // stop rule with one-batch lag (simulator.cc:117 checks before every frame)
        uint64_t errs = counters[1];
        for (int k = 0; k < NL; k++) errs += seen[k][1];
        if (max_err_blk && errs >= max_err_blk) break;
      }
      const int nb = (int)std::min<uint64_t>((uint64_t)c->max_batch, frame_count - done);
      GenParams g = gen_params(c, nb, snr_db, seed, frame_begin + done);
      KML_LAUNCH(c, launch_gen_encode(g, l.u_packed.p, l.c_packed.p, l.stream));
      KML_LAUNCH(c, launch_channel(g, l.c_packed.p, nullptr, nullptr, l.h.p, l.y.p, l.stream));
      KML_RC(receive_on_lane(c, l, l.stream, nb, l.y.p, 0, l.h.p, var));
      KML_LAUNCH(c, launch_count_errors(nb, c->K, c->k_words, l.u_packed.p, l.uu_hat_packed.p, l.ret.p, c->opts.max_iter,
                                       l.counters.p, l.stream));
      pending[li] = true;
      done += nb;
      if (!sequential) li = (li + 1) % NL;
    }
    return KML_OK;
  };
  int rc = run();
  for (int k = 0; k < NL && rc == KML_OK; k++) rc = read_lane(k);
  for (int k = 0; k < NL; k++) {
    cudaStreamSynchronize(c->lane[k].stream);
    lane_release(c, c->lane[k], c->lane[k].stream);
  }
  if (rc != KML_OK) return rc;
  for (int j = 0; j < NL; j++) {
    for (int k = 0; k < 4; k++) counters[k] += seen[j][k];
    if (iters_sum) *iters_sum += seen[j][4];
  }
  return KML_OK;
}

// Per-frame view of the fused path (the reference's per-frame log lines, simulator.cc:124-126,149-152, kmcodec.cc:64,132-136)
extern "C" int kml_simulate_frames(kml_ctx *c, double snr_db, uint64_t seed, uint64_t frame_begin, int count,
                                   uint64_t counters[4], float *h, float *hhat, float *metric, int32_t *kstar, int32_t *ret) {
  KML_RC(check_batch(c, count));
  if (!counters) return fail_arg(c, "kml_simulate_frames: null counters");
  if (count > c->max_batch) return fail_arg(c, "kml_simulate_frames: count exceeds max_batch");
  const double var = std::pow(10.0, -0.1 * snr_db);
  Lane &l = c->lane[0];
  cudaStream_t s = l.stream;
  KML_ENTER(c, l, s);
  KML_CUDA(c, cudaMemsetAsync(l.counters.p, 0, 5 * sizeof(unsigned long long), s));
  GenParams g = gen_params(c, count, snr_db, seed, frame_begin);
  KML_LAUNCH(c, launch_gen_encode(g, l.u_packed.p, l.c_packed.p, s));
  KML_LAUNCH(c, launch_channel(g, l.c_packed.p, nullptr, nullptr, l.h.p, l.y.p, s));
  KML_RC(receive_on_lane(c, l, s, count, l.y.p, 0, l.h.p, var));
  KML_LAUNCH(c, launch_count_errors(count, c->K, c->k_words, l.u_packed.p, l.uu_hat_packed.p, l.ret.p, c->opts.max_iter, l.counters.p, s));
  if (h) KML_CUDA(c, cudaMemcpyAsync(h, l.h.p, sizeof(float2) * count, cudaMemcpyDeviceToHost, s));
  if (ret) KML_CUDA(c, cudaMemcpyAsync(ret, l.ret.p, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, s));
  if (!c->opts.known_h) {
    if (hhat) KML_CUDA(c, cudaMemcpyAsync(hhat, l.hhat.p, sizeof(float2) * count, cudaMemcpyDeviceToHost, s));
    if (metric) KML_CUDA(c, cudaMemcpyAsync(metric, l.metric.p, sizeof(float) * 4 * count, cudaMemcpyDeviceToHost, s));
    if (kstar) KML_CUDA(c, cudaMemcpyAsync(kstar, l.kstar.p, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, s));
  }
  KML_CUDA(c, cudaMemcpyAsync(c->h_counters, l.counters.p, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  KML_CUDA(c, cudaStreamSynchronize(s));
  KML_LEAVE(c, l, s);
  for (int k = 0; k < 4; k++) counters[k] += c->h_counters[k];
  return KML_OK;
}

namespace {
// GetHistogramData + CntErr for nb frames whose symbols and information bits sit in l.y / l.u_packed
// (kmcodec.cc:74-79, simulator.cc:154-167): l.metric[nb][4], c->counters += CntErr on a uu_hat no final decoder wrote.
int histogram_on_lane(kml_ctx *c, Lane &l, cudaStream_t s, int nb, double var) {
  const bool decode_metric = c->is_5g || c->opts.metric_type;
  KML_LAUNCH(c, launch_kmeans(nb, l.y.p, 0, c->n_sym, c->points.p, c->Q, c->km, c->opts.kmeans_iter, l.hhat.p, nullptr,
                             l.passes.p, nullptr, c->num_sms, s));
  KML_RC(resolve_on_lane(c, l, s, nb, var));
  if (decode_metric) {
    // uu_hat as the reference leaves it: written by the LAST candidate's metric decode (kmcodec.cc:126-131,148,157)
    KML_LAUNCH(c, launch_extract_bits(nb, c->K, c->info_offset, 4 * c->words_n, l.cc_hat_packed.p + 3 * c->words_n,
                                     l.uu_hat_packed.p, s));
  } else {  // hard metric: nothing ever writes cdata.uu_hat_ (uninitialised in the reference, zeros here)
    KML_CUDA(c, cudaMemsetAsync(l.uu_hat_packed.p, 0, sizeof(uint32_t) * (size_t)nb * c->k_words, s));
  }
  KML_LAUNCH(c, launch_count_errors(nb, c->K, c->k_words, l.u_packed.p, l.uu_hat_packed.p, nullptr, c->opts.max_iter,
                                   c->counters.p, s));
  return KML_OK;
}
}  // namespace

extern "C" int kml_histogram(kml_ctx *c, double snr_db, uint64_t seed, uint64_t frame_begin, uint64_t frame_count,
                             float *metrics, uint64_t counters[4]) {
  if (!c || !metrics || !counters) return KML_ERR_ARG;
  if (c->opts.known_h) return fail_arg(c, "kml_histogram: needs the four blind candidates (true_h_arg = false)");
  const double var = std::pow(10.0, -0.1 * snr_db);
  Lane &l = c->lane[0];
  cudaStream_t s = l.stream;
  KML_ENTER(c, l, s);
  KML_CUDA(c, cudaMemsetAsync(c->counters.p, 0, 5 * sizeof(unsigned long long), s));
  for (uint64_t done = 0; done < frame_count;) {
    const int nb = (int)std::min<uint64_t>((uint64_t)c->max_batch, frame_count - done);
    GenParams g = gen_params(c, nb, snr_db, seed, frame_begin + done);
    KML_LAUNCH(c, launch_gen_bits(g, l.u_packed.p, s));
    KML_LAUNCH(c, launch_encode(g, l.u_packed.p, l.c_packed.p, s));
    KML_LAUNCH(c, launch_channel(g, l.c_packed.p, nullptr, nullptr, l.h.p, l.y.p, s));
    KML_RC(histogram_on_lane(c, l, s, nb, var));
    KML_CUDA(c, cudaMemcpyAsync(metrics + done * 4, l.metric.p, sizeof(float) * 4 * nb, cudaMemcpyDeviceToHost, s));
    KML_CUDA(c, cudaStreamSynchronize(s));
    done += nb;
  }
  KML_CUDA(c, cudaMemcpyAsync(c->h_counters, c->counters.p, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  KML_CUDA(c, cudaStreamSynchronize(s));
  KML_LEAVE(c, l, s);
  for (int k = 0; k < 4; k++) counters[k] += c->h_counters[k];
  return KML_OK;
}

extern "C" int kml_histogram_rx(kml_ctx *c, int B, const float *y, double var, const uint32_t *u_packed, float *metrics,
                                int32_t *kstar, uint32_t *uu_hat_packed, uint64_t counters[4]) {
  KML_RC(check_batch(c, B));
  if (!y || !u_packed || !metrics || !counters || !(var > 0)) return fail_arg(c, "kml_histogram_rx: bad argument");
  if (c->opts.known_h) return fail_arg(c, "kml_histogram_rx: needs the four blind candidates (true_h_arg = false)");
  Lane &l = c->lane[0];
  cudaStream_t s = l.stream;
  KML_ENTER(c, l, s);
  KML_CUDA(c, cudaMemsetAsync(c->counters.p, 0, 5 * sizeof(unsigned long long), s));
  for (int b0 = 0; b0 < B; b0 += c->max_batch) {
    const int nb = std::min(c->max_batch, B - b0);
    KML_CUDA(c, cudaMemcpyAsync(l.y.p, y + (size_t)b0 * c->n_sym * 2, sizeof(float2) * nb * c->n_sym, cudaMemcpyHostToDevice, s));
    KML_CUDA(c, cudaMemcpyAsync(l.u_packed.p, u_packed + (size_t)b0 * c->k_words, sizeof(uint32_t) * nb * c->k_words, cudaMemcpyHostToDevice, s));
    KML_RC(histogram_on_lane(c, l, s, nb, var));
    KML_CUDA(c, cudaMemcpyAsync(metrics + (size_t)b0 * 4, l.metric.p, sizeof(float) * 4 * nb, cudaMemcpyDeviceToHost, s));
    if (kstar) KML_CUDA(c, cudaMemcpyAsync(kstar + b0, l.kstar.p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, s));
    if (uu_hat_packed)
      KML_CUDA(c, cudaMemcpyAsync(uu_hat_packed + (size_t)b0 * c->k_words, l.uu_hat_packed.p, sizeof(uint32_t) * nb * c->k_words, cudaMemcpyDeviceToHost, s));
    KML_CUDA(c, cudaStreamSynchronize(s));
  }
  KML_CUDA(c, cudaMemcpyAsync(c->h_counters, c->counters.p, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  KML_CUDA(c, cudaStreamSynchronize(s));
  KML_LEAVE(c, l, s);
  for (int k = 0; k < 4; k++) counters[k] += c->h_counters[k];
  return KML_OK;
}

// ================================================================================================ device-pointer variants
// Asynchronous on the caller's stream; scratch comes from lane 0.  Calls on different streams of one context are
// ordered on the device by the lane's ownership event (lane_acquire), so they are safe — and serial.
namespace {
int check_dev_batch(kml_ctx *c, int B, const char *who) {
  KML_RC(check_batch(c, B));
  if (B > c->max_batch) {
    c->err = std::string(who) + ": B exceeds max_batch";
    return KML_ERR_ARG;
  }
  return KML_OK;
}
}  // namespace

extern "C" int kml_generate_dev(kml_ctx *c, int B, double snr_db, uint64_t seed, uint64_t frame0, uint32_t *u_packed,
                                float *h, float *y, void *stream) {
  KML_RC(check_dev_batch(c, B, "kml_generate_dev"));
  if (!u_packed || !h || !y) return fail_arg(c, "kml_generate_dev: null buffer");
  cudaStream_t s = (cudaStream_t)stream;
  Lane &l = c->lane[0];
  KML_ENTER(c, l, s);
  GenParams g = gen_params(c, B, snr_db, seed, frame0);
  KML_LAUNCH(c, launch_gen_bits(g, u_packed, s));
  KML_LAUNCH(c, launch_encode(g, u_packed, l.c_packed.p, s));
  KML_LAUNCH(c, launch_channel(g, l.c_packed.p, nullptr, nullptr, (float2 *)h, (float2 *)y, s));
  KML_LEAVE(c, l, s);
  return KML_OK;
}

extern "C" int kml_kmeans_dev(kml_ctx *c, int B, const float *y, float *hhat, int32_t *passes, void *stream) {
  KML_RC(check_batch(c, B));
  if (!y || !hhat) return fail_arg(c, "kml_kmeans_dev: null buffer");
  KML_CUDA(c, cudaSetDevice(c->device));  // (no scratch: reads y, writes hhat / passes)
  KML_LAUNCH(c, launch_kmeans(B, y, 0, c->n_sym, c->points.p, c->Q, c->km, c->opts.kmeans_iter, (float2 *)hhat, nullptr,
                             passes, nullptr, c->num_sms, (cudaStream_t)stream));
  return KML_OK;
}

extern "C" int kml_receive_dev(kml_ctx *c, int B, const float *y, const float *true_h, double var, uint32_t *uu_hat_packed,
                               int32_t *ret, void *stream) {
  KML_RC(check_dev_batch(c, B, "kml_receive_dev"));
  if (!y || !(var > 0) || (c->opts.known_h && !true_h)) return fail_arg(c, "kml_receive_dev: bad argument");
  Lane &l = c->lane[0];
  cudaStream_t s = (cudaStream_t)stream;
  KML_ENTER(c, l, s);
  KML_RC(receive_on_lane(c, l, s, B, y, 0, (const float2 *)true_h, var));
  if (uu_hat_packed)
    KML_CUDA(c, cudaMemcpyAsync(uu_hat_packed, l.uu_hat_packed.p, sizeof(uint32_t) * (size_t)B * c->k_words, cudaMemcpyDeviceToDevice, s));
  if (ret) KML_CUDA(c, cudaMemcpyAsync(ret, l.ret.p, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToDevice, s));
  KML_LEAVE(c, l, s);
  return KML_OK;
}

extern "C" int kml_demap_dev(kml_ctx *c, int B, const float *y, const float *h, double var, float *llr, void *stream) {
  KML_RC(check_dev_batch(c, B, "kml_demap_dev"));
  if (!y || !h || !llr || !(var > 0)) return fail_arg(c, "kml_demap_dev: bad argument");
  Lane &l = c->lane[0];
  KML_CUDA(c, cudaSetDevice(c->device));  // (no scratch: the ratios go straight to the caller's buffer)
  DemapParams d = demap_params(c, l, B, var, 1, 0);
  d.y = (const float2 *)y;
  d.h = (const float2 *)h;
  d.lr = llr;
  KML_LAUNCH(c, launch_demap(d, c->num_sms, (cudaStream_t)stream));
  KML_LAUNCH(c, launch_lr_to_llr((size_t)B * c->n_tx, llr, llr, (cudaStream_t)stream));
  return KML_OK;
}

extern "C" int kml_decode_dev(kml_ctx *c, int B, const float *llr, int in_is_lr, int iter_count, uint32_t *cc_hat_packed,
                              int32_t *ret, void *stream) {
  KML_RC(check_dev_batch(c, B, "kml_decode_dev"));
  if (!llr || !cc_hat_packed || !ret || iter_count < 1) return fail_arg(c, "kml_decode_dev: bad argument");
  Lane &l = c->lane[0];
  cudaStream_t s = (cudaStream_t)stream;
  KML_ENTER(c, l, s);  // (the frame queue's counter is lane scratch)
  DecParams p = dec_params(c, l, B, llr, nullptr, 1, in_is_lr, iter_count, cc_hat_packed, ret, nullptr);
  KML_LAUNCH(c, dec_launch(c->dl, p, c->num_sms, s));
  KML_LEAVE(c, l, s);
  return KML_OK;
}

extern "C" int kml_count_errors_dev(kml_ctx *c, int B, const uint32_t *u_packed, const uint32_t *uu_hat_packed,
                                    uint64_t *counters_dev, void *stream) {
  KML_RC(check_batch(c, B));
  if (!u_packed || !uu_hat_packed || !counters_dev) return fail_arg(c, "kml_count_errors_dev: null buffer");
  KML_CUDA(c, cudaSetDevice(c->device));
  KML_LAUNCH(c, launch_count_errors(B, c->K, c->k_words, u_packed, uu_hat_packed, nullptr, c->opts.max_iter,
                                   (unsigned long long *)counters_dev, (cudaStream_t)stream));
  return KML_OK;
}
