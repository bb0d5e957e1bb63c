// TEST INFRASTRUCTURE ONLY — never linked into, imported by or executed from the product path.
//
// Deterministic driver around the UNMODIFIED reference classes (trganda/kmldpc), compiled from the
// sources where they lie under /root/reference by oracle/Makefile into oracle/_ref/ref_harness.
// It mirrors the reference's per-frame loop (kmldpc/src/simulator.cc:118-166) and the candidate
// resolver (kmldpc/src/kmcodec.cc:54-72,105-163) through public members only, with the global LCG
// seeded by CLCRandNum::SetSeed(-1) (state 17, lib/lab/src/randnum.cc:10-11) and a single thread,
// and dumps every intermediate tensor so the C restatement (oracle/kml_oracle.c) and the CUDA path
// can be pinned against the reference itself.
//
// Modes
//   dump  : per-frame tensors + graph/encoder export → raw little-endian files in <out>/
//   time  : stage timings of the reference's own code path (bench.py --impl reference, cpu_baseline)
//
// Usage: ref_harness <dump|time> key=value ...
//   cfgdir=<dir with H + constellation files>  matrix=<file> modem=<file> g5=0|1 active=0|1
//   known_h=0|1 metric_type=0|1 metric_iter=5 max_iter=50 snr=10 frames=8 out=<dir> skip=<frames> seed=<LCG state>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <unistd.h>
#include <vector>

#include "binary5gldpccodec.h"
#include "binaryldpccodec.h"
#include "kmcodec.h"
#include "kmeans.h"
#include "log.h"
#include "modemlinearsystem.h"
#include "randnum.h"
#include "sourcesink.h"
#include "toml.hpp"

namespace {

using cplx = std::complex<double>;

struct Args {
  std::map<std::string, std::string> kv;
  std::string s(const std::string &k, const std::string &d) const {
    auto it = kv.find(k);
    return it == kv.end() ? d : it->second;
  }
  double d(const std::string &k, double dflt) const {
    auto it = kv.find(k);
    return it == kv.end() ? dflt : atof(it->second.c_str());
  }
  long i(const std::string &k, long dflt) const {
    auto it = kv.find(k);
    return it == kv.end() ? dflt : atol(it->second.c_str());
  }
};

// Subclasses only widen access to protected members for the one-off graph/encoder export.
struct PegExport : public lab::BinaryLDPCCodec {
  explicit PegExport(const toml::value &a) : lab::BinaryLDPCCodec(a) {}
  int chk() const { return code_chk_; }
  int rows() const { return num_row_; }
  int cols() const { return num_col_; }
  bool active() const { return encoder_active_; }
  lab::Edge *rh() const { return row_head_; }
  lab::Edge *ch() const { return col_head_; }
  char **enc() const { return enc_h_; }
};
struct G5Export : public lab::Binary5GLDPCCodec {
  explicit G5Export(const toml::value &a) : lab::Binary5GLDPCCodec(a) {}
  int chk() const { return code_chk_; }
  int rows() const { return num_row_; }
  int cols() const { return num_col_; }
  bool active() const { return encoder_active_; }
  lab::Edge *rh() const { return row_head_; }
  lab::Edge *ch() const { return col_head_; }
  char **enc() const { return enc_h_; }
};

template <class T>
void put(FILE *f, const T *p, size_t n) {
  if (fwrite(p, sizeof(T), n, f) != n) {
    perror("fwrite");
    exit(2);
  }
}

struct Out {
  std::string dir;
  std::map<std::string, FILE *> files;
  FILE *get(const std::string &name) {
    auto it = files.find(name);
    if (it != files.end()) return it->second;
    FILE *f = fopen((dir + "/" + name).c_str(), "wb");
    if (!f) {
      perror(name.c_str());
      exit(2);
    }
    files[name] = f;
    return f;
  }
  void close_all() {
    for (auto &kv : files) fclose(kv.second);
    files.clear();
  }
};

void put_bits(FILE *f, const int *b, int n) {
  std::vector<int8_t> t(n);
  for (int i = 0; i < n; i++) t[i] = (int8_t)b[i];
  put(f, t.data(), t.size());
}

template <class C>
void export_code(const C &c, Out &out, int n_tx, int k) {
  // Row adjacency in the traversal order of THIS object (head-inserted → descending columns),
  // column adjacency likewise (descending rows).  CSR with int32.
  std::vector<int32_t> rp(1, 0), ci, cp(1, 0), ri;
  for (int r = 0; r < c.rows(); r++) {
    for (lab::Edge *e = (c.rh() + r)->right; e->m_col_no != -1; e = e->right) ci.push_back(e->m_col_no);
    rp.push_back((int32_t)ci.size());
  }
  for (int v = 0; v < c.cols(); v++) {
    for (lab::Edge *e = (c.ch() + v)->down; e->m_row_no != -1; e = e->down) ri.push_back(e->m_row_no);
    cp.push_back((int32_t)ri.size());
  }
  put(out.get("row_ptr.i32"), rp.data(), rp.size());
  put(out.get("col_idx.i32"), ci.data(), ci.size());
  put(out.get("col_ptr.i32"), cp.data(), cp.size());
  put(out.get("row_idx.i32"), ri.data(), ri.size());
  if (c.active()) {
    FILE *f = out.get("enc_h.i8");
    for (int r = 0; r < c.rows(); r++) put(f, (const int8_t *)c.enc()[r], c.cols());
  }
  int32_t meta[8] = {c.rows(), c.cols(), c.chk(), c.code_dim(), n_tx, k, c.active() ? 1 : 0, 0};
  put(out.get("code_meta.i32"), meta, 8);
}

double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}// namespace

int
main(int argc, char **argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <dump|time> key=value ...\n", argv[0]);
    return 1;
  }
  std::string mode = argv[1];
  Args a;
  for (int i = 2; i < argc; i++) {
    std::string s = argv[i];
    auto p = s.find('=');
    if (p == std::string::npos) continue;
    a.kv[s.substr(0, p)] = s.substr(p + 1);
  }
  const std::string cfgdir = a.s("cfgdir", ".");
  const std::string matrix = a.s("matrix", "PEG2304regular0.5.txt");
  const std::string modem = a.s("modem", "2bits_QPSK.txt");
  const bool g5 = a.i("g5", 0) != 0;
  const bool active = a.i("active", 1) != 0;
  const bool known_h = a.i("known_h", 0) != 0;
  const bool metric_type = a.i("metric_type", 0) != 0;
  const int metric_iter = (int)a.i("metric_iter", 5);
  const int max_iter = (int)a.i("max_iter", 50);
  const double snr = a.d("snr", 10.0);
  const long frames = a.i("frames", 8);
  const long skip = a.i("skip", 0);
  const bool hist = a.i("hist", 0) != 0;  // [histogram] enable = true: KmCodec::GetHistogramData instead of Decoder
  std::string outdir = a.s("out", "");
  if (!outdir.empty() && outdir[0] != '/') {
    char cwd[4096];
    if (getcwd(cwd, sizeof cwd)) outdir = std::string(cwd) + "/" + outdir;
  }
  // The reference opens its data files by bare relative name (binaryldpccodec.cc:73-80, modem.cc:88-92).
  if (chdir(cfgdir.c_str()) != 0) {
    perror("chdir cfgdir");
    return 2;
  }

  // Logger needs a TeeStream (log.cc:71-75 casts unconditionally).
  std::ofstream devnull("/dev/null");
  lab::logger::TeeStream tee(devnull, devnull);
  lab::logger::Log::get().set_log_stream(tee);
  lab::logger::Log::get().set_log_level(lab::logger::Error);

  lab::CLCRandNum::Get().SetSeed(-1);// state = 17
  const long seed = a.i("seed", 17);
  if (seed != 17 && seed > 0) {
    // any other start state: SetSeed(flag > 0) reads it from stdin (lib/lab/src/randnum.cc:19-25)
    char tmpl[] = "/tmp/kml_seed_XXXXXX";
    int fd = mkstemp(tmpl);
    if (fd >= 0) {
      dprintf(fd, "%ld\n", seed);
      close(fd);
      if (freopen(tmpl, "r", stdin)) lab::CLCRandNum::Get().SetSeed(1);
      unlink(tmpl);
      printf("\n");
    }
  }
  lab::CWHRandNum::Get().SetSeed(-1);

  std::ostringstream ts;
  ts << "[range]\nminimum_snr = " << snr << "\nmaximum_snr = " << snr
     << "\nstep_snr = 1.0\nmaximum_error_number = 1000000\nmaximum_block_number = " << frames
     << "\nthread_block_number = 1\n[decoder]\ntrue_h_arg = " << (known_h ? "true" : "false")
     << "\n[xcodec]\n5gldpc = " << (g5 ? "true" : "false") << "\nmetric_type = " << (metric_type ? "true" : "false")
     << "\nmetric_iter = " << metric_iter << "\n[histogram]\nenable = false\n[ldpc]\nmax_iter = " << max_iter
     << "\nactive = " << (active ? "true" : "false") << "\nmatrix_file = \"" << matrix << "\"\n[modem]\nmodem_file = \""
     << modem << "\"\n";
  std::istringstream tis(ts.str());
  const toml::value args = toml::parse(tis, "harness.toml");

  double t_setup = now();
  std::unique_ptr<lab::BinaryLDPCCodec> codec;
  int n_tx, k, n_graph, two_z = 0;
  PegExport *peg = nullptr;
  G5Export *g5c = nullptr;
  if (g5) {
    g5c = new G5Export(args);
    codec.reset(g5c);
    n_tx = g5c->code_len_puncture();
    n_graph = g5c->cols();
    two_z = n_graph - n_tx;
  } else {
    peg = new PegExport(args);
    codec.reset(peg);
    n_tx = peg->code_len();
    n_graph = n_tx;
  }
  k = codec->code_dim();
  KmCodec kmcodec(args);// the reference's own resolver + decoder glue (second, independent codec instance)
  lab::ModemLinearSystem mls(args, n_tx);
  t_setup = now() - t_setup;
  const double var = pow(10.0, -0.1 * snr);// simulator.cc:74-77
  mls.set_sigma(sqrt(var));
  mls.set_var(var);
  auto cons = mls.constellations();
  const int q = (int)cons.size();
  int m = 0;
  while ((1 << m) < q) m++;
  const int n_sym = n_tx / m;

  Out out;
  out.dir = outdir;
  const bool dump = (mode == "dump");
  if (dump) {
    if (outdir.empty()) {
      fprintf(stderr, "dump needs out=<dir>\n");
      return 1;
    }
    if (g5) export_code(*g5c, out, n_tx, k);
    else
      export_code(*peg, out, n_tx, k);
    put(out.get("constellation.f64"), (const double *)cons.data(), 2 * cons.size());
    int32_t meta[8] = {n_tx, n_graph, k, m, q, n_sym, two_z, (int32_t)frames};
    put(out.get("run_meta.i32"), meta, 8);
    double dm[2] = {snr, var};
    put(out.get("run_meta.f64"), dm, 2);
  }

  std::vector<int> uu(k), uu_hat(k), uu_hat2(k), cc(n_tx), rr(n_tx);
  std::vector<double> bit_in(n_tx), bit_out(n_tx);
  lab::SourceSink ssink;
  ssink.ClrCnt();
  double t_src = 0, t_chan = 0, t_km = 0, t_res = 0, t_dem = 0, t_dec = 0, t_ref_decoder = 0;
  long iters_total = 0;
  long mismatch_kmcodec = 0;

  // syndrom_soft_ is `new double[...]`, never initialised (binaryldpccodec.cc:88), and only the check-node phase writes it
  // (binaryldpccodec.cc:274): the soft metric of a decode that leaves at iteration 0 reads whatever the previous Decoder
  // call left — for the very first call that is heap garbage.  Give it a defined start (all ones) through the public
  // accessor; from then on the reference's own stale-value chain runs unchanged.
  for (int j = 0; j < codec->num_row(); j++) codec->syndrom_soft()[j] = 1.0;

  for (long f = 0; f < skip + frames; f++) {
    const bool rec = dump && f >= skip;
    double t0 = now();
    ssink.GetBitStr(uu.data(), k);                  // simulator.cc:118
    codec->Encoder(uu.data(), cc.data());           // simulator.cc:119
    double t1 = now();
    cplx true_h;
    lab::CLCRandNum::Get().Normal(true_h);          // simulator.cc:121-123
    true_h *= sqrt(0.5);
    std::vector<cplx> gen_h(1, true_h);
    mls.PartitionModemLSystem(cc.data(), gen_h);    // simulator.cc:130
    double t2 = now();
    std::vector<cplx> h_hats;
    auto yy = mls.GetRecvSymbol();
    std::vector<cplx> clusters(q, cplx(0, 0));
    cplx h_hat(0, 0);
    if (known_h) {
      h_hats.push_back(true_h);
    } else {
      kmldpc::KMeans km(yy, cons, 20);             // simulator.cc:140-142
      km.Run();
      clusters = km.clusters();
      h_hat = clusters[0] / cons[0];
      for (size_t j = 0; j < 4; j++) h_hats.push_back(h_hat * exp(cplx(0, (lab::kPi / 2) * j)));
    }
    double t3 = now();
    if (hist) {  // simulator.cc:154-167 through the reference's own KmCodec::GetHistogramData (kmcodec.cc:74-79)
      if (f == 0) std::fill(uu_hat.begin(), uu_hat.end(), 0);  // cdata.uu_hat_ is never initialised (simulator.h:30-37)
      std::vector<double> met = kmcodec.GetHistogramData(mls, h_hats, uu_hat.data());
      const int lo = (int)(std::min_element(met.begin(), met.end()) - met.begin());
      double line[4];
      for (int j = 0; j < 4; j++) line[j] = met[(lo + j) % 4];  // the rotated line of histogram_<snr>.txt (:157-160)
      ssink.CntErr(uu.data(), uu_hat.data(), k, 1);
      if (rec) {
        put(out.get("h.f64"), (const double *)&true_h, 2);
        put(out.get("hhat.f64"), (const double *)&h_hat, 2);
        put(out.get("metric.f64"), met.data(), 4);
        put(out.get("hist_line.f64"), line, 4);
        int32_t ks = lo;
        put(out.get("kstar.i32"), &ks, 1);
        put_bits(out.get("uu_hat.i8"), uu_hat.data(), k);
        int nerr = 0;
        for (int t = 0; t < k; t++) nerr += (uu[t] != uu_hat[t]);
        int32_t ne = nerr;
        put(out.get("nerr.i32"), &ne, 1);
      }
      continue;
    }
    // ---- mirror of KmCodec::Decoder / GetMetrics / Metric / GetParityCheck (kmcodec.cc:54-163)
    double metrics[4] = {0, 0, 0, 0};
    int kstar = 0;
    if (h_hats.size() > 1) {
      for (size_t i = 0; i < h_hats.size(); i++) {
        std::vector<std::pair<int, cplx>> th = {{0, h_hats[i]}};
        for (int t = 0; t < n_tx; t++) bit_in[t] = 0.5;
        mls.DeMapping(th, bit_in.data(), bit_out.data());
        double metric;
        if (metric_type) {
          codec->Decoder(bit_out.data(), uu_hat.data(), metric_iter);
          metric = 0.0;
          for (int j = 0; j < codec->num_row(); j++) metric += log(codec->syndrom_soft()[j]);
        } else {
          if (g5) {
            codec->Decoder(bit_out.data(), uu_hat.data(), metric_iter);
            metric = codec->ParityCheck(codec->cc_hat());
          } else {
            for (int t = 0; t < n_tx; t++) rr[t] = bit_out[t] > 0.5 ? 1 : 0;
            metric = codec->ParityCheck(rr.data());
          }
        }
        metrics[i] = std::abs(metric);
      }
      kstar = (int)(std::min_element(metrics, metrics + 4) - metrics);
    }
    double t4 = now();
    std::vector<std::pair<int, cplx>> th = {{0, h_hats[kstar]}};
    for (int t = 0; t < n_tx; t++) bit_in[t] = 0.5;
    mls.DeMapping(th, bit_in.data(), bit_out.data());
    double t5 = now();
    int ret = codec->Decoder(bit_out.data(), uu_hat.data(), codec->max_iter());
    double t6 = now();
    // ---- the reference's own glue on the same inputs (must agree bit for bit)
    kmcodec.Decoder(mls, h_hats, uu_hat2.data());
    double t7 = now();
    for (int t = 0; t < k; t++)
      if (uu_hat[t] != uu_hat2[t]) {
        mismatch_kmcodec++;
        break;
      }
    ssink.CntErr(uu.data(), uu_hat.data(), k, 1);
    if (f >= skip) {
      t_src += t1 - t0;
      t_chan += t2 - t1;
      t_km += t3 - t2;
      t_res += t4 - t3;
      t_dem += t5 - t4;
      t_dec += t6 - t5;
      t_ref_decoder += t7 - t6;
      iters_total += (ret > max_iter ? max_iter : ret);
    }
    if (rec) {
      put_bits(out.get("u.i8"), uu.data(), k);
      put_bits(out.get("c.i8"), cc.data(), n_tx);
      put(out.get("h.f64"), (const double *)&true_h, 2);
      put(out.get("y.f64"), (const double *)yy.data(), 2 * yy.size());
      put(out.get("clusters.f64"), (const double *)clusters.data(), 2 * clusters.size());
      put(out.get("hhat.f64"), (const double *)&h_hat, 2);
      put(out.get("metric.f64"), metrics, 4);
      int32_t ks = kstar;
      put(out.get("kstar.i32"), &ks, 1);
      put(out.get("p0.f64"), bit_out.data(), n_tx);
      put_bits(out.get("cc_hat.i8"), codec->cc_hat(), n_graph);
      put_bits(out.get("uu_hat.i8"), uu_hat.data(), k);
      int32_t r32 = ret;
      put(out.get("ret.i32"), &r32, 1);
      int nerr = 0;
      for (int t = 0; t < k; t++) nerr += (uu[t] != uu_hat[t]);
      int32_t ne = nerr;
      put(out.get("nerr.i32"), &ne, 1);
    }
  }
  out.close_all();
  printf("{\"mode\":\"%s\",\"frames\":%ld,\"k\":%d,\"n_tx\":%d,\"n_graph\":%d,\"q\":%d,\"snr\":%.6f,"
         "\"tot_blk\":%u,\"err_blk\":%u,\"ber\":%.14f,\"fer\":%.14f,\"avg_ret\":%.4f,\"kmcodec_mismatch\":%ld,"
         "\"t_setup\":%.6f,\"t_src_enc\":%.6f,\"t_chan\":%.6f,\"t_kmeans\":%.6f,\"t_resolve\":%.6f,\"t_demap\":%.6f,"
         "\"t_decode\":%.6f,\"t_kmcodec_decoder\":%.6f}\n",
         mode.c_str(), frames, k, n_tx, n_graph, q, snr, ssink.tot_blk(), ssink.err_blk(), ssink.ber(), ssink.fer(),
         frames ? (double)iters_total / frames : 0.0, mismatch_kmcodec, t_setup, t_src, t_chan, t_km, t_res, t_dem,
         t_dec, t_ref_decoder);
  return 0;
}
