// Internal declarations shared by the host-side C++ and the CUDA translation units of libkmldpc_b200.so.
#ifndef KML_INTERNAL_H
#define KML_INTERNAL_H
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/kmldpc_b200.h"

namespace kml {

void set_global_error(const std::string &msg);
const char *global_error();

// Truncated pi of the reference (lib/lab/include/utility.h:10) — the candidate rotations use it.
constexpr double kRefPi = 3.14159265358979;
// lib/lab/include/utility.h:12
constexpr float kSmallProbF = 1.0e-12f;
constexpr float kLrMin = 1.0e-12f;  // P0/P1 after the reference's clip of P0 to [1e-12, 1-1e-12]
constexpr float kLrMax = 1.0e12f;

// Philox stream ids (word 3 of the counter)
enum : uint32_t { STREAM_BITS = 0x5eed0001u, STREAM_FADE = 0x5eed0002u, STREAM_NOISE = 0x5eed0003u };

// host_code.cpp
std::vector<int> voronoi_neighbours_of_first(const double *pts, int q);
// Environment knobs.  knob() = a run-time switch between kernels that are all SHIPPED and all reproduce the reference
// (the fallback layouts / tilings the tests exercise): when one is set, the library says so on stderr — a stray
// variable must not change the kernel silently.  tuning_knob() = A/B and timing-ablation variants (some with knowingly
// WRONG results): compiled in only with -DKML_TUNING, otherwise the variable is ignored.
const char *knob(const char *name);
inline const char *tuning_knob(const char *name) {
#ifdef KML_TUNING
  return knob(name);
#else
  (void)name;
  return nullptr;
#endif
}

// layout_opt.cpp
int optimize_decoder_layout(int M, int N, int n_slots, int plane, int slot_stride, const int32_t *row_ptr,
                            const int32_t *col_idx,
                            const std::vector<int> &group_of_var, int n_groups, int slots_per_var,
                            std::vector<int> &slot_of_row, std::vector<int> &pos_of_edge,
                            std::vector<std::vector<int>> &edge_order, int *excess_wavefronts);

}  // namespace kml
#endif
