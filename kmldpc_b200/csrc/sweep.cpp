// Host-side sweep driver: what Simulator::Simulator + Simulator::Simulate + Simulator::run do around the per-frame loop
// (src/simulator.cc:3-109), on top of kml_simulate.  One host thread per GPU; frames of an SNR point are handed out in
// chunks from a shared cursor; with more than one GPU the per-GPU 64-bit counters of a point are summed by ONE
// ncclAllReduce over NVLink (single process, ncclCommInitAll — the path's only collective); lines are
// formatted exactly like SourceSink::PrintResult (lib/lab/src/sourcesink.cc:50-65) and the tables of simulator.cc:48-66.
#include <algorithm>
#include <condition_variable>
#include <array>
#include <atomic>
#include <cctype>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include "kml_internal.h"

namespace {

// ---- the counter reduction over NCCL.  libnccl is opened at run time (dlopen), not linked: a host process that already
// carries an NCCL (torch bundles its own) keeps exactly one copy, and a single-GPU run never touches it.
class CounterReducer {
 public:
  // returns false (with `why`) when NCCL cannot be used
  bool init(int n_gpus, std::string &why) {
    G_ = n_gpus;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      lib_ = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (lib_) break;
    }
    if (!lib_) {
      why = std::string("dlopen(libnccl.so.2): ") + dlerror();
      return false;
    }
#define KML_SYM(f)                                         \
  f##_ = reinterpret_cast<decltype(&f)>(dlsym(lib_, #f)); \
  if (!f##_) {                                             \
    why = "libnccl lacks " #f;                             \
    return false;                                          \
  }
    KML_SYM(ncclCommInitAll) KML_SYM(ncclCommDestroy) KML_SYM(ncclAllReduce) KML_SYM(ncclGroupStart) KML_SYM(ncclGroupEnd)
    KML_SYM(ncclGetErrorString)
#undef KML_SYM
    comms_.assign(G_, nullptr);
    std::vector<int> devs(G_);
    for (int g = 0; g < G_; g++) devs[g] = g;
    ncclResult_t r = ncclCommInitAll_(comms_.data(), G_, devs.data());
    if (r != ncclSuccess) {
      why = std::string("ncclCommInitAll: ") + ncclGetErrorString_(r);
      comms_.clear();
      return false;
    }
    dbuf_.assign(G_, nullptr);
    st_.assign(G_, nullptr);
    for (int g = 0; g < G_; g++) {
      if (cudaSetDevice(g) != cudaSuccess || cudaMalloc(&dbuf_[g], kMaxWords * sizeof(unsigned long long)) != cudaSuccess ||
          cudaStreamCreateWithFlags(&st_[g], cudaStreamNonBlocking) != cudaSuccess) {
        why = "device buffers for the counter reduction";
        return false;
      }
    }
    ok_ = true;
    return true;
  }
  bool ok() const { return ok_; }
  // mine[g * count .. g * count + count) = counters GPU g accumulated; tot[count] = their sum (read back from GPU 0)
  bool reduce(const uint64_t *mine, int count, uint64_t *tot, std::string &why) {
    if (count < 1 || count > kMaxWords) {
      why = "counter block size";
      return false;
    }
    for (int g = 0; g < G_; g++) {
      cudaSetDevice(g);
      if (cudaMemcpyAsync(dbuf_[g], mine + (size_t)g * count, count * sizeof(uint64_t), cudaMemcpyHostToDevice, st_[g]) != cudaSuccess) {
        why = "cudaMemcpyAsync (counters)";
        return false;
      }
    }
    ncclGroupStart_();
    for (int g = 0; g < G_; g++) {
      ncclResult_t r = ncclAllReduce_(dbuf_[g], dbuf_[g], count, ncclUint64, ncclSum, comms_[g], st_[g]);
      if (r != ncclSuccess) {
        ncclGroupEnd_();
        why = std::string("ncclAllReduce: ") + ncclGetErrorString_(r);
        return false;
      }
    }
    ncclResult_t r = ncclGroupEnd_();
    if (r != ncclSuccess) {
      why = std::string("ncclGroupEnd: ") + ncclGetErrorString_(r);
      return false;
    }
    for (int g = 0; g < G_; g++) {
      cudaSetDevice(g);
      if (g == 0) cudaMemcpyAsync(tot, dbuf_[0], count * sizeof(uint64_t), cudaMemcpyDeviceToHost, st_[0]);
      if (cudaStreamSynchronize(st_[g]) != cudaSuccess) {
        why = "stream sync after the counter reduction";
        return false;
      }
    }
    return true;
  }
  ~CounterReducer() {
    for (int g = 0; g < (int)st_.size(); g++) {
      cudaSetDevice(g);
      if (st_[g]) cudaStreamDestroy(st_[g]);
      if (dbuf_[g]) cudaFree(dbuf_[g]);
    }
    if (ncclCommDestroy_)
      for (ncclComm_t c : comms_)
        if (c) ncclCommDestroy_(c);
    // the library stays open: other users of the process (torch) may share it
  }

  int gpus() const { return G_; }
  static constexpr int kMaxWords = 1024;

 private:
  int G_ = 0;
  bool ok_ = false;
  void *lib_ = nullptr;
  decltype(&ncclCommInitAll) ncclCommInitAll_ = nullptr;
  decltype(&ncclCommDestroy) ncclCommDestroy_ = nullptr;
  decltype(&ncclAllReduce) ncclAllReduce_ = nullptr;
  decltype(&ncclGroupStart) ncclGroupStart_ = nullptr;
  decltype(&ncclGroupEnd) ncclGroupEnd_ = nullptr;
  decltype(&ncclGetErrorString) ncclGetErrorString_ = nullptr;
  std::vector<ncclComm_t> comms_;
  std::vector<void *> dbuf_;
  std::vector<cudaStream_t> st_;
};

using kml::set_global_error;

// ---- the TOML subset the reference's config.toml uses: [table], key = number | true | false | "string", # comments
struct Toml {
  std::map<std::string, std::string> kv;  // "table.key" → raw value (strings unquoted)
  bool has(const std::string &k) const { return kv.count(k) != 0; }
};

std::string trim(const std::string &s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) a++;
  while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}

bool parse_toml(const char *path, Toml &t, std::string &err) {
  std::ifstream in(path, std::ios::binary);
  if (!in.is_open()) {
    err = std::string("cannot open ") + path;
    return false;
  }
  std::string line, table;
  int ln = 0;
  while (std::getline(in, line)) {
    ln++;
    if (ln == 1 && line.size() >= 3 && (unsigned char)line[0] == 0xEF) line = line.substr(3);  // BOM
    // strip comments outside strings
    bool in_str = false;
    for (size_t i = 0; i < line.size(); i++) {
      if (line[i] == '"') in_str = !in_str;
      else if (line[i] == '#' && !in_str) {
        line.resize(i);
        break;
      }
    }
    line = trim(line);
    if (line.empty()) continue;
    if (line.front() == '[') {
      if (line.back() != ']') {
        err = "line " + std::to_string(ln) + ": bad table header";
        return false;
      }
      table = trim(line.substr(1, line.size() - 2));
      continue;
    }
    const size_t eq = line.find('=');
    if (eq == std::string::npos) {
      err = "line " + std::to_string(ln) + ": expected key = value";
      return false;
    }
    std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
    if (key.size() >= 2 && key.front() == '"' && key.back() == '"') key = key.substr(1, key.size() - 2);
    if (val.size() >= 2 && val.front() == '"' && val.back() == '"') val = val.substr(1, val.size() - 2);
    t.kv[table.empty() ? key : table + "." + key] = val;
  }
  return true;
}

bool get_num(const Toml &t, const std::string &k, double &out, std::string &err, bool required = true) {
  auto it = t.kv.find(k);
  if (it == t.kv.end()) {
    if (required) err = "missing key " + k;  // toml11 throws here (SURVEY §5 "Config")
    return !required;
  }
  char *end = nullptr;
  std::string v = it->second;
  v.erase(std::remove(v.begin(), v.end(), '_'), v.end());
  out = std::strtod(v.c_str(), &end);
  if (end == v.c_str() || *end) {
    err = "key " + k + ": not a number";
    return false;
  }
  return true;
}

bool get_bool(const Toml &t, const std::string &k, int &out, std::string &err, bool required = true) {
  auto it = t.kv.find(k);
  if (it == t.kv.end()) {
    if (required) err = "missing key " + k;
    return !required;
  }
  if (it->second == "true") out = 1;
  else if (it->second == "false") out = 0;
  else {
    err = "key " + k + ": not a boolean";
    return false;
  }
  return true;
}

std::string fmt_point_line(double snr, uint64_t tot_blk, uint64_t err_blk, uint64_t err_bit, double ber, double fer) {
  std::stringstream s;  // sourcesink.cc:50-65
  s << std::fixed << std::setprecision(3) << std::setfill('0') << "SNR = " << std::setw(3) << std::right << snr << ' '
    << "Total blk = " << std::setw(7) << std::right << std::setprecision(0) << tot_blk << ' '
    << "Error blk = " << std::setw(7) << std::right << err_blk << ' '
    << "Error bit = " << std::setw(7) << std::right << err_bit << ' ' << std::fixed << std::setprecision(14)
    << "BER = " << ber << ' ' << "FER = " << fer;
  return s.str();
}

std::string fmt_table_row(double snr, double v) {
  std::stringstream s;  // simulator.cc:52-56
  s << std::fixed << std::setprecision(3) << std::setfill('0') << std::setw(3) << std::right << snr << ' '
    << std::setprecision(14) << v;
  return s.str();
}

}  // namespace

extern "C" int kml_sweep_cfg_load(const char *config_toml, kml_sweep_cfg *cfg) {
  if (!config_toml || !cfg) return KML_ERR_ARG;
  Toml t;
  std::string err;
  if (!parse_toml(config_toml, t, err)) {
    set_global_error(err);
    return KML_ERR_IO;
  }
  std::memset(cfg, 0, sizeof *cfg);
  double d = 0;
  int b = 0;
  bool ok = true;
  ok = ok && get_num(t, "range.minimum_snr", cfg->min_snr, err);
  ok = ok && get_num(t, "range.maximum_snr", cfg->max_snr, err);
  ok = ok && get_num(t, "range.step_snr", cfg->step_snr, err);
  if (ok && (ok = get_num(t, "range.maximum_error_number", d, err))) cfg->max_err_blk = (uint64_t)d;
  if (ok && (ok = get_num(t, "range.maximum_block_number", d, err))) cfg->max_num_blk = (uint64_t)d;
  ok = ok && get_num(t, "range.thread_block_number", d, err);  // CPU fan-out granularity: read, unused
  if (ok && (ok = get_bool(t, "decoder.true_h_arg", b, err))) cfg->known_h = b;
  if (ok && (ok = get_bool(t, "xcodec.5gldpc", b, err))) cfg->is_5g = b;
  if (ok && (ok = get_bool(t, "xcodec.metric_type", b, err))) cfg->metric_type = b;
  if (ok && (ok = get_num(t, "xcodec.metric_iter", d, err))) cfg->metric_iter = (int)d;
  if (ok && (ok = get_bool(t, "histogram.enable", b, err))) cfg->histogram_enable = b;
  if (ok && (ok = get_num(t, "ldpc.max_iter", d, err))) cfg->max_iter = (int)d;
  if (ok && (ok = get_bool(t, "ldpc.active", b, err))) cfg->encoder_active = b;
  if (ok && !t.has("ldpc.matrix_file")) { ok = false; err = "missing key ldpc.matrix_file"; }
  if (ok && !t.has("modem.modem_file")) { ok = false; err = "missing key modem.modem_file"; }
  if (!ok) {
    set_global_error(std::string(config_toml) + ": " + err);
    return KML_ERR_IO;
  }
  std::snprintf(cfg->matrix_file, sizeof cfg->matrix_file, "%s", t.kv["ldpc.matrix_file"].c_str());
  std::snprintf(cfg->modem_file, sizeof cfg->modem_file, "%s", t.kv["modem.modem_file"].c_str());
  // optional [gpu] table — unknown tables are ignored by the reference binary, so one file serves both
  cfg->seed = 17;  // echo of the reference's fixed LCG state
  cfg->n_gpus = 1;
  cfg->max_batch = 0;
  cfg->early_exit = 1;
  if (get_num(t, "gpu.seed", d, err, false) && t.has("gpu.seed")) cfg->seed = (uint64_t)d;
  if (get_num(t, "gpu.gpus", d, err, false) && t.has("gpu.gpus")) cfg->n_gpus = (int)d;
  if (get_num(t, "gpu.batch", d, err, false) && t.has("gpu.batch")) cfg->max_batch = (int)d;
  if (get_bool(t, "gpu.early_exit", b, err, false) && t.has("gpu.early_exit")) cfg->early_exit = b;
  if (get_num(t, "gpu.algorithm", d, err, false) && t.has("gpu.algorithm")) cfg->algorithm = (int)d;
  if (t.has("gpu.reduce")) cfg->reduce_on_host = t.kv["gpu.reduce"] == "host";
  if (get_bool(t, "gpu.debug", b, err, false) && t.has("gpu.debug")) cfg->debug_frames = b;
  return KML_OK;
}

extern "C" int kml_sweep_points(const kml_sweep_cfg *cfg) {
  if (!cfg || !(cfg->step_snr > 0)) return 0;
  return (int)(unsigned long)((cfg->max_snr - cfg->min_snr) / cfg->step_snr + 1);  // simulator.cc:27
}

namespace {
std::mutex g_timing_mu;
double g_timing[2] = {0.0, 0.0};
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

extern "C" void kml_sweep_last_timing(double seconds[2]) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  seconds[0] = g_timing[0];
  seconds[1] = g_timing[1];
}

extern "C" int kml_sweep_run(const kml_sweep_cfg *cfg, const char *data_dir, double *ber, double *fer, uint64_t *counters,
                             void (*log_cb)(const char *, void *), void *user) {
  if (!cfg) return KML_ERR_ARG;
  const double t_begin = now_s();
  double t_points = t_begin;
  if (cfg->histogram_enable && cfg->known_h) {
    set_global_error("[histogram] enable = true needs blind detection (true_h_arg = false): one candidate has no histogram");
    return KML_ERR_ARG;
  }
  const int n_pts = kml_sweep_points(cfg);
  if (n_pts < 1) {
    set_global_error("empty SNR range");
    return KML_ERR_ARG;
  }
  auto log = [&](const std::string &s) {
    if (log_cb) log_cb(s.c_str(), user);
  };
  auto path_of = [&](const char *f) {
    std::string p = f;
    if (data_dir && *data_dir && !p.empty() && p[0] != '/') p = std::string(data_dir) + "/" + p;
    return p;
  };
  log(cfg->is_5g ? "Using 5G LDPC." : "Using traditional LDPC.");  // kmcodec.cc:27,32 (member init precedes the ctor body)
  {
    std::stringstream s;  // simulator.cc:16-21
    s << '[' << std::fixed << std::setprecision(3) << cfg->min_snr << ',' << cfg->step_snr << ',' << cfg->max_snr << ']';
    log(s.str());
    s.str("");
    s << '[' << "MAX_ERROR_BLK = " << cfg->max_err_blk << ',' << "MAX_BLK = " << cfg->max_num_blk << ']';
    log(s.str());
  }
  kml_code *code = nullptr;
  kml_modem *modem = nullptr;
  int rc = kml_code_load(path_of(cfg->matrix_file).c_str(), cfg->is_5g, cfg->encoder_active, &code);
  if (rc != KML_OK) return rc;
  rc = kml_modem_load(path_of(cfg->modem_file).c_str(), &modem);
  if (rc != KML_OK) {
    kml_code_free(code);
    return rc;
  }
  const int G = cfg->n_gpus > 0 ? cfg->n_gpus : 1;
  std::vector<double> ber_v(n_pts, 0.0), fer_v(n_pts, 0.0);
  kml_opts o{};
  o.max_iter = cfg->max_iter; o.known_h = cfg->known_h; o.metric_type = cfg->metric_type; o.metric_iter = cfg->metric_iter;
  o.kmeans_iter = 20; o.early_exit = cfg->early_exit; o.max_batch = cfg->max_batch; o.algorithm = cfg->algorithm;
  std::vector<kml_ctx *> ctx(G, nullptr);
  {  // device 0 first (it runs the layout annealing, which the others then take from the cache), the rest in parallel
    rc = kml_create(&ctx[0], 0, code, modem, &o);
    std::vector<int> rcs(G, KML_OK);
    std::vector<std::string> errs(G);
    std::vector<std::thread> th;
    for (int g = 1; g < G && rc == KML_OK; g++)
      th.emplace_back([&, g] {
        rcs[g] = kml_create(&ctx[g], g, code, modem, &o);
        if (rcs[g] != KML_OK) errs[g] = kml_last_error(nullptr);  // thread-local message: carry it over
      });
    for (auto &t : th) t.join();
    for (int g = 1; g < G && rc == KML_OK; g++)
      if (rcs[g] != KML_OK) {
        rc = rcs[g];
        set_global_error(errs[g]);
      }
  }
  // The counters of the G GPUs are summed by ONE ncclAllReduce per SNR point.  A host-side sum is used
  // only when the configuration asks for it ([gpu] reduce = "host"); NCCL failing is an error, not a silent fallback.
  CounterReducer reducer;
  if (rc == KML_OK && G > 1 && !cfg->reduce_on_host) {
    std::string why;
    if (!reducer.init(G, why)) {
      set_global_error("counter reduction over NCCL unavailable (" + why + "); set [gpu] reduce = \"host\" to sum on the host");
      rc = KML_ERR_NCCL;
    }
  }
  if (rc == KML_OK) {
    int32_t info[8];
    kml_info(ctx[0], info);
    // Frames are handed out in chunks of several batches, so that kml_simulate's two lanes overlap inside a call and its
    // own (lagged, per-batch) stop rule applies.  The GPUs take chunks from ONE queue that runs through the SNR points in
    // order: a GPU that finds a point's frames handed out moves on to the next point instead of waiting for the others, so
    // the only tail is at the end of the sweep.  (With a barrier per point, C5 on 8 GPUs spent a quarter of its time in 31
    // tails.)  Guided self-scheduling: a chunk is a quarter of an even share of ALL frames left, between 1 and 8 batches.
    // The caller's thread waits for the points in order, reduces each one's counters and logs its line as it completes.
    const uint64_t batch = (uint64_t)info[7];
    for (int g = 0; g < G; g++) {  // setup ends when every device is idle
      cudaSetDevice(g);
      cudaDeviceSynchronize();
    }
    t_points = now_s();
    struct Point {
      uint64_t cursor = 0, err_blk = 0;
      int inflight = 0;
      std::vector<uint64_t> mine;  // [G][4]
    };
    std::vector<Point> pts((size_t)n_pts);
    for (auto &pt : pts) pt.mine.assign((size_t)G * 4, 0);
    std::mutex mu;
    std::condition_variable cv;
    int cur = 0, failed = KML_OK;
    std::string err_msg;  // kml_last_error() of a failing worker (its thread-local copy dies with the thread)
    // simulator.cc:117 leaves before the first frame when err_blk >= maximum_error_number — with 0 that is at once
    const bool no_frames = cfg->max_err_blk == 0;
    const bool debug = cfg->debug_frames && !cfg->histogram_enable && !no_frames;
    const bool serial_mode = cfg->histogram_enable || debug || no_frames;  // frames in index order on GPU 0 (or none)
    auto point_seed_of = [&](int i) { return cfg->seed + (uint64_t)i * 0x9E3779B97F4A7C15ull; };  // disjoint Philox streams per point
    auto point_open = [&](const Point &pt) { return pt.cursor < cfg->max_num_blk && pt.err_blk < cfg->max_err_blk; };  // simulator.cc:117
    auto worker = [&](int g) {
      for (;;) {
        int i;
        uint64_t begin, count, seen;
        {
          std::lock_guard<std::mutex> lk(mu);
          while (cur < n_pts && !point_open(pts[cur])) cur++;
          if (cur >= n_pts || failed != KML_OK) break;
          Point &pt = pts[cur];
          const uint64_t left_here = cfg->max_num_blk - pt.cursor;
          const uint64_t left_all = left_here + (uint64_t)(n_pts - 1 - cur) * cfg->max_num_blk;
          uint64_t want = (left_all / (4 * (uint64_t)G) + batch - 1) / batch * batch;
          want = std::min<uint64_t>(std::max<uint64_t>(want, batch), 8 * batch);
          i = cur;
          begin = pt.cursor;
          count = std::min(want, left_here);
          seen = pt.err_blk;
          pt.cursor += count;
          pt.inflight++;
        }
        uint64_t cnt[4] = {0, 0, 0, 0};
        // what is left of the error budget goes down with the call: kml_simulate stops between its batches
        const int r = kml_simulate(ctx[g], cfg->min_snr + cfg->step_snr * i, point_seed_of(i), begin, count, cfg->max_err_blk - seen, cnt,
                                   nullptr);
        {
          std::lock_guard<std::mutex> lk(mu);
          Point &pt = pts[i];
          if (r != KML_OK && failed == KML_OK) {
            failed = r;
            err_msg = kml_last_error(ctx[g]);
          }
          pt.err_blk += cnt[1];  // (the stop rule sees this host-side total while other chunks are in flight)
          for (int k = 0; k < 4; k++) pt.mine[(size_t)g * 4 + k] += cnt[k];
          pt.inflight--;
        }
        cv.notify_all();
        if (r != KML_OK) break;
      }
      cv.notify_all();
    };
    std::vector<std::thread> th;
    if (!serial_mode)
      for (int g = 0; g < G; g++) th.emplace_back(worker, g);
    for (int i = 0; i < n_pts && rc == KML_OK; i++) {
      const double snr = cfg->min_snr + cfg->step_snr * i;
      const uint64_t point_seed = point_seed_of(i);
      uint64_t tot[4] = {0, 0, 0, 0};
      std::vector<uint64_t> &mine = pts[i].mine;  // per GPU, this point
      // histogram mode (simulator.cc:81-84,154-162): "histogram_<snr>.txt" in the working directory, one line per frame
      // with the four metrics rotated to start at the (first) minimum; frames in index order on GPU 0.
      if (cfg->histogram_enable && !no_frames) {
        const std::string fname = "histogram_" + std::to_string(snr) + ".txt";
        std::ofstream hout(fname);
        const uint64_t hchunk = (uint64_t)info[7];
        std::vector<float> met((size_t)hchunk * 4);
        for (uint64_t begin = 0; begin < cfg->max_num_blk && rc == KML_OK; begin += hchunk) {
          if (tot[1] >= cfg->max_err_blk) break;
          const uint64_t count = std::min<uint64_t>(hchunk, cfg->max_num_blk - begin);
          rc = kml_histogram(ctx[0], snr, point_seed, begin, count, met.data(), tot);
          if (rc != KML_OK) set_global_error(kml_last_error(ctx[0]));
          for (uint64_t f = 0; rc == KML_OK && f < count; f++) {
            const float *m = &met[f * 4];
            int lo = 0;
            for (int k = 1; k < 4; k++)
              if (m[k] < m[lo]) lo = k;
            for (int k = lo; k < lo + 4; k++) hout << m[k % 4] << ' ';
            hout << std::endl;
          }
        }
      }
      // [gpu] debug = true: the reference's per-frame log lines, frames in index order on GPU 0 (simulator.cc:124-126,
      // 149-152; kmcodec.cc:64,132-136).  The same frames and counters as the normal path, one batch at a time.
      if (debug) {
        const int nb_max = info[7];
        std::vector<float> dh(2 * (size_t)nb_max), dhh(2 * (size_t)nb_max), dm(4 * (size_t)nb_max);
        std::vector<int32_t> dk(nb_max), dr(nb_max);
        for (uint64_t begin = 0; begin < cfg->max_num_blk && rc == KML_OK; begin += (uint64_t)nb_max) {
          if (tot[1] >= cfg->max_err_blk) break;
          const int count = (int)std::min<uint64_t>((uint64_t)nb_max, cfg->max_num_blk - begin);
          const uint64_t blk0 = tot[0];
          rc = kml_simulate_frames(ctx[0], snr, point_seed, begin, count, tot, dh.data(), dhh.data(), dm.data(), dk.data(), dr.data());
          if (rc != KML_OK) { set_global_error(kml_last_error(ctx[0])); break; }
          for (int f = 0; f < count; f++) {
            std::stringstream st;
            st << "Generated H = " << std::complex<double>(dh[2 * f], dh[2 * f + 1]);
            log(st.str());
            if (cfg->known_h) continue;
            st.str("");
            st << std::fixed << std::setprecision(0) << std::setfill('0') << "Current Block Number = " << std::setw(7) << std::right
               << (blk0 + (uint64_t)f + 1);
            log(st.str());
            const std::complex<double> hh(dhh[2 * f], dhh[2 * f + 1]);
            for (int k = 0; k < 4; k++) {
              st.str("");
              st.clear();
              const double m = cfg->metric_type ? -(double)dm[4 * f + k] : (double)dm[4 * f + k];  // Metric() before std::abs
              st << std::fixed << std::setprecision(14) << "Hhat = " << hh * std::exp(std::complex<double>(0, (kml::kRefPi / 2) * k))
                 << " Metric = " << std::setw(5) << std::right << m;
              log(st.str());
            }
            log("hatIndex = " + std::to_string(dk[f]));
          }
        }
      }
      if (!serial_mode) {  // this point is complete when its frames are handed out (or its budget is spent) and have come back
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return failed != KML_OK || (!point_open(pts[i]) && pts[i].inflight == 0); });
        if (failed != KML_OK) {
          rc = failed;
          set_global_error(err_msg);  // on the CALLER's thread
        }
      }
      if (!cfg->histogram_enable) {
        std::string why;
        if (reducer.ok() && rc == KML_OK) {  // the path's one collective: 4 x uint64 per point over NCCL / NVLink
          if (!reducer.reduce(mine.data(), 4, tot, why)) {
            set_global_error("counter reduction: " + why);
            rc = KML_ERR_NCCL;
          }
        } else {
          for (int g = 0; g < G; g++)
            for (int k = 0; k < 4; k++) tot[k] += mine[(size_t)g * 4 + k];
        }
      }
      // (0 / 0 like the reference's SourceSink::ber()/fer() when nothing ran)
      const double b = (double)tot[3] / (double)tot[2], f = (double)tot[1] / (double)tot[0];
      ber_v[i] = b;
      fer_v[i] = f;
      if (ber) ber[i] = b;
      if (fer) fer[i] = f;
      if (counters)
        for (int k = 0; k < 4; k++) counters[(size_t)i * 4 + k] = tot[k];
      log(fmt_point_line(snr, tot[0], tot[1], tot[3], b, f));
    }
    if (rc != KML_OK) {  // (a failure outside the workers: stop them)
      std::lock_guard<std::mutex> lk(mu);
      if (failed == KML_OK) failed = rc;
    }
    for (auto &t : th) t.join();
    if (rc == KML_OK) {
      log("BER Result");
      for (int i = 0; i < n_pts; i++) log(fmt_table_row(cfg->min_snr + cfg->step_snr * i, ber_v[i]));
      log("FER Result");
      for (int i = 0; i < n_pts; i++) log(fmt_table_row(cfg->min_snr + cfg->step_snr * i, fer_v[i]));
    }
  }
  {
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_timing[0] = t_points - t_begin;
    g_timing[1] = now_s() - t_points;
  }
  for (auto *c : ctx) kml_destroy(c);
  kml_modem_free(modem);
  kml_code_free(code);
  return rc;
}

// ---- comm_init / reduce_counters as entry points of their own (kml_sweep_run uses the same class)
struct kml_comm {
  CounterReducer r;
  std::string err;
};

extern "C" int kml_comm_init(int n_gpus, kml_comm **out) {
  if (!out || n_gpus < 1) return KML_ERR_ARG;
  *out = nullptr;
  auto *c = new kml_comm();
  std::string why;
  if (!c->r.init(n_gpus, why)) {
    set_global_error("kml_comm_init: " + why);
    delete c;
    return KML_ERR_NCCL;
  }
  *out = c;
  return KML_OK;
}

extern "C" int kml_reduce_counters(kml_comm *comm, const uint64_t *per_gpu, int count, uint64_t *total) {
  if (!comm || !per_gpu || !total) return KML_ERR_ARG;
  std::string why;
  if (!comm->r.reduce(per_gpu, count, total, why)) {
    comm->err = why;
    set_global_error("kml_reduce_counters: " + why);
    return KML_ERR_NCCL;
  }
  return KML_OK;
}

extern "C" void kml_comm_destroy(kml_comm *comm) { delete comm; }
