// Drop-in replacement of the reference's kmldpc/include/simulator.h for the GPU path.
//
// Same public surface as the reference (`explicit Simulator(toml::value)`, `void Simulate()`,
// kmldpc/include/simulator.h:56-60), so the reference's kmldpc.cpp compiles UNCHANGED against it
// (kmldpc.cpp:31-33).  Construction reads the same config keys with toml11 exactly like
// src/simulator.cc:3-22, src/kmcodec.cc:20-40, lib/lab/src/binaryldpccodec.cc:62-73 and lib/lab/src/modem.cc:4-9
// (missing key → toml11 throws, as in the reference); Simulate() hands the sweep to libkmldpc_b200.so
// (kml_sweep_run) and forwards its lines to the reference's logger, so logs and BER/FER tables look the same.
// Optional GPU-only knobs live in a [gpu] table the CPU binary ignores: seed, gpus, batch, early_exit, algorithm, reduce, debug.
#ifndef KMLDPC_B200_SIMULATOR_FACADE_H
#define KMLDPC_B200_SIMULATOR_FACADE_H

#include <cstdio>
#include <cstring>
#include <string>
#include <utility>

#include "kmldpc_b200.h"
#include "log.h"
#include "toml.hpp"

class Simulator {
 public:
  explicit Simulator(toml::value arguments) : arguments_(std::move(arguments)) {
    std::memset(&cfg_, 0, sizeof cfg_);
    const auto range = toml::find(arguments_, "range");
    cfg_.min_snr = toml::find<double>(range, "minimum_snr");
    cfg_.max_snr = toml::find<double>(range, "maximum_snr");
    cfg_.step_snr = toml::find<double>(range, "step_snr");
    cfg_.max_err_blk = (uint64_t)toml::find<int>(range, "maximum_error_number");
    cfg_.max_num_blk = (uint64_t)toml::find<int>(range, "maximum_block_number");
    (void)toml::find<int>(range, "thread_block_number");  // CPU fan-out granularity; meaningless on the GPU
    cfg_.known_h = toml::find<bool>(toml::find(arguments_, "decoder"), "true_h_arg");
    const auto xcodec = toml::find(arguments_, "xcodec");
    cfg_.is_5g = toml::find<bool>(xcodec, "5gldpc");
    cfg_.metric_type = toml::find<bool>(xcodec, "metric_type");
    cfg_.metric_iter = toml::find<int>(xcodec, "metric_iter");
    cfg_.histogram_enable = toml::find<bool>(toml::find(arguments_, "histogram"), "enable");
    const auto ldpc = toml::find(arguments_, "ldpc");
    cfg_.max_iter = toml::find<int>(ldpc, "max_iter");
    cfg_.encoder_active = toml::find<bool>(ldpc, "active");
    std::snprintf(cfg_.matrix_file, sizeof cfg_.matrix_file, "%s", toml::find<std::string>(ldpc, "matrix_file").c_str());
    std::snprintf(cfg_.modem_file, sizeof cfg_.modem_file, "%s",
                  toml::find<std::string>(toml::find(arguments_, "modem"), "modem_file").c_str());
    cfg_.seed = 17;
    cfg_.n_gpus = 1;
    cfg_.early_exit = 1;
    if (arguments_.contains("gpu")) {
      const auto gpu = toml::find(arguments_, "gpu");
      cfg_.seed = (uint64_t)toml::find_or<std::int64_t>(gpu, "seed", 17);
      cfg_.n_gpus = (int)toml::find_or<std::int64_t>(gpu, "gpus", 1);
      cfg_.max_batch = (int)toml::find_or<std::int64_t>(gpu, "batch", 0);
      cfg_.early_exit = toml::find_or<bool>(gpu, "early_exit", true) ? 1 : 0;
      cfg_.algorithm = (int)toml::find_or<std::int64_t>(gpu, "algorithm", 0);  // 1-3 = min-sum variants (throughput mode)
      cfg_.reduce_on_host = toml::find_or<std::string>(gpu, "reduce", "nccl") == "host" ? 1 : 0;
      cfg_.debug_frames = toml::find_or<bool>(gpu, "debug", false) ? 1 : 0;  // the per-frame lines of the reference's log file
    }
  }
  virtual ~Simulator() = default;

  void Simulate() {
    // files are opened by bare relative name in the working directory, like the reference (kmldpc.cpp:29)
    const int rc = kml_sweep_run(&cfg_, ".", nullptr, nullptr, nullptr, &Simulator::on_line, nullptr);
    if (rc != KML_OK)
      lab::logger::ERROR(std::string("kmldpc_b200: ") + kml_last_error(nullptr) + " (rc=" + std::to_string(rc) + ")", true);
  }

 private:
  // per-frame lines go to the log file only, like the reference's INFO(…, false) (simulator.cc:126,152; kmcodec.cc:64,136)
  static void on_line(const char *line, void *) {
    static const char *const quiet[] = {"Generated H = ", "Current Block Number = ", "Hhat = ", "hatIndex = "};
    bool both = true;
    for (const char *q : quiet) both = both && std::strncmp(line, q, std::strlen(q)) != 0;
    lab::logger::INFO(line, both);
  }
  const toml::value arguments_;
  kml_sweep_cfg cfg_;
};
#endif
