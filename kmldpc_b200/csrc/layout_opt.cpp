// Shared-memory layout optimiser for the BP decoder (host side, runs once per code; results are cached in-process).
//
// The decoder stores the message of edge e = (row r, position k inside the row) at word
// k * plane + slot(r) * slot_stride:  "planar" (slot_stride 1, plane = row slots + 1; any row degree) or "row-major"
// (slot_stride = row degree, plane 1: a check node's words are contiguous and move with 64-bit LDS/STS).
// Check-node threads own one slot each, so their accesses are conflict free for ANY assignment.  A warp of
// variable-node threads gathers 32 arbitrary edges per load: with slot(r) = r and plane % 32 == 0 the PEG graphs give
// 2.3–3.8-way bank conflicts on every variable-node access (ncu, profiles/r1a_bp_decoder_full.txt).
//
// With plane % 32 == 1 the bank of an edge is (slot(r) + k) mod 32, so both the slot of a row and the ORDER of the
// edges inside a row (which the check node is indifferent to) move edges between banks.  The optimiser
//   1. anneals  sum_{group g, bank b} max(0, load[g][b] - cap_g)  over row-slot swaps and in-row position swaps, where a
//      group is the 32 consecutive variables one warp handles in one unrolled step and cap_g is the largest variable
//      degree in the group (= the number of gather instructions the warp issues for it);
//   2. colours, per group, the bipartite multigraph (variables x banks) with cap_g colours (König: possible whenever
//      no bank is overloaded; overloaded banks are split into virtual banks first, which leaves exactly the residual
//      2-way conflicts the annealer could not remove).  Colour = which of the variable's gather instructions fetches
//      the edge; two edges of one colour never share a (virtual) bank.
// Deterministic (fixed seed).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "kml_internal.h"

namespace kml {

namespace {
struct Rng {
  uint64_t s;
  uint32_t next() {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (uint32_t)(s >> 33);
  }
  uint32_t below(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 31); }
  float unit() { return (next() & 0xffffff) * (1.0f / 16777216.0f); }
};

struct Cached {
  // the full key: the hash only finds the bucket, a hit needs these to compare equal
  std::vector<int32_t> row_ptr, col_idx;
  std::vector<int> group_of_var;
  int par[5];
  std::vector<int> slot, pos;
  std::vector<std::vector<int>> order;
  int residual, excess;
};
std::mutex g_mu;
std::multimap<uint64_t, Cached> g_cache;

uint64_t fnv(uint64_t h, const void *p, size_t n) {
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}
}  // namespace

// Inputs : CSR of the graph; group_of_var[v] (variables of one group are fetched by the same warp instructions);
//          n_slots >= M row slots; plane = words between consecutive k planes; slots_per_var = gather instructions per
//          variable (>= max variable degree).
// Outputs: slot_of_row[M]; pos_of_edge[E] (position k of each CSR edge inside its row, a permutation per row);
//          edge_order[v][i] = CSR edge fetched by variable v's i-th gather instruction (-1 = none);
//          *excess_wavefronts = shared-memory wavefronts above one per warp gather that remain (0 = conflict free).
// Returns the residual annealing cost.
int optimize_decoder_layout(int M, int N, int n_slots, int plane, int slot_stride, const int32_t *row_ptr,
                            const int32_t *col_idx,
                            const std::vector<int> &group_of_var, int n_groups, int slots_per_var,
                            std::vector<int> &slot_of_row, std::vector<int> &pos_of_edge,
                            std::vector<std::vector<int>> &edge_order, int *excess_wavefronts) {
  const int B = 32, E = row_ptr[M];
  uint64_t key = fnv(1469598103934665603ull, row_ptr, sizeof(int32_t) * (M + 1));
  key = fnv(key, col_idx, sizeof(int32_t) * E);
  key = fnv(key, group_of_var.data(), sizeof(int) * N);
  const int par[5] = {n_slots, plane, slots_per_var, n_groups, slot_stride};
  key = fnv(key, par, sizeof par);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto range = g_cache.equal_range(key);
    for (auto it = range.first; it != range.second; ++it) {
      const Cached &cd = it->second;
      if (!std::equal(par, par + 5, cd.par) || cd.row_ptr.size() != (size_t)M + 1 || cd.col_idx.size() != (size_t)E ||
          !std::equal(cd.row_ptr.begin(), cd.row_ptr.end(), row_ptr) || !std::equal(cd.col_idx.begin(), cd.col_idx.end(), col_idx) ||
          cd.group_of_var != group_of_var)
        continue;  // a different graph behind the same 64-bit hash
      slot_of_row = it->second.slot;
      pos_of_edge = it->second.pos;
      edge_order = it->second.order;
      if (excess_wavefronts) *excess_wavefronts = it->second.excess;
      return it->second.residual;
    }
  }
  std::vector<int> cap(n_groups, 0), vdeg(N, 0), erow(E), egrp(E);
  for (int r = 0; r < M; r++)
    for (int e = row_ptr[r]; e < row_ptr[r + 1]; e++) {
      vdeg[col_idx[e]]++;
      erow[e] = r;
      egrp[e] = group_of_var[col_idx[e]];
    }
  for (int v = 0; v < N; v++) cap[group_of_var[v]] = std::max(cap[group_of_var[v]], vdeg[v]);
  std::vector<int> slot(n_slots), pos(E);  // slot[r] for r >= M are empty place holders (so rows can move to free slots)
  for (int r = 0; r < n_slots; r++) slot[r] = r;
  for (int r = 0; r < M; r++)
    for (int e = row_ptr[r]; e < row_ptr[r + 1]; e++) pos[e] = e - row_ptr[r];
  auto ebank = [&](int e) { return (slot[erow[e]] * slot_stride + pos[e] * plane) & (B - 1); };
  std::vector<int16_t> load((size_t)n_groups * B, 0);
  auto cell = [&](int g, int b) -> int16_t & { return load[(size_t)g * B + b]; };
  for (int e = 0; e < E; e++) cell(egrp[e], ebank(e))++;
  long cost = 0;
  for (int g = 0; g < n_groups; g++)
    for (int b = 0; b < B; b++) cost += std::max(0, cell(g, b) - cap[g]);
  auto rem = [&](int e) {
    int16_t &a = cell(egrp[e], ebank(e));
    const int d = a > cap[egrp[e]] ? -1 : 0;
    a--;
    return d;
  };
  auto add = [&](int e) {
    int16_t &a = cell(egrp[e], ebank(e));
    a++;
    return a > cap[egrp[e]] ? 1 : 0;
  };
  auto rem_row = [&](int r) { int d = 0; if (r < M) for (int e = row_ptr[r]; e < row_ptr[r + 1]; e++) d += rem(e); return d; };
  auto add_row = [&](int r) { int d = 0; if (r < M) for (int e = row_ptr[r]; e < row_ptr[r + 1]; e++) d += add(e); return d; };
  Rng rng{0x9E3779B97F4A7C15ull};
  const long moves = std::min<long>(12000000, 600L * E);
  const double t0 = 0.4, t1 = 0.12;
  float temp = (float)t0;
  std::vector<int> hot;
  for (long it = 0; it < moves && cost > 0; it++) {
    if ((it & 0x3fff) == 0) temp = (float)(t0 * std::pow(t1 / t0, (double)it / moves));
    // pick a row that currently sits in an overloaded cell: the list of edges in overloaded cells is rebuilt every 4096
    // moves (late in the run only a handful of cells are over capacity, and blind sampling almost never finds them)
    if ((it & 0xfff) == 0) {
      hot.clear();
      for (int e = 0; e < E; e++)
        if (cell(egrp[e], ebank(e)) > cap[egrp[e]]) hot.push_back(e);
    }
    int r1 = (int)rng.below(M);
    if (!hot.empty() && (rng.next() & 3) != 0) r1 = erow[hot[rng.below((uint32_t)hot.size())]];
    int d = 0;
    if (it & 1) {  // swap the positions of two edges inside row r1
      const int deg = row_ptr[r1 + 1] - row_ptr[r1];
      if (deg < 2 || (plane & (B - 1)) == 0) continue;
      const int a = row_ptr[r1] + (int)rng.below(deg), b = row_ptr[r1] + (int)rng.below(deg);
      if (a == b) continue;
      d = rem(a) + rem(b);
      std::swap(pos[a], pos[b]);
      d += add(a) + add(b);
      if (d <= 0 || rng.unit() < std::exp(-(float)d / temp)) cost += d;
      else {
        rem(a); rem(b);
        std::swap(pos[a], pos[b]);
        add(a); add(b);
      }
    } else {  // swap the slots of two rows
      const int r2 = (int)rng.below(n_slots);
      if ((((slot[r1] - slot[r2]) * slot_stride) & (B - 1)) == 0) continue;
      d = rem_row(r1) + rem_row(r2);
      std::swap(slot[r1], slot[r2]);
      d += add_row(r1) + add_row(r2);
      if (d <= 0 || rng.unit() < std::exp(-(float)d / temp)) cost += d;
      else {
        rem_row(r1); rem_row(r2);
        std::swap(slot[r1], slot[r2]);
        add_row(r1); add_row(r2);
      }
    }
  }
  slot_of_row.assign(slot.begin(), slot.begin() + M);
  pos_of_edge = pos;

  // ---- per group: colour the (variable, virtual bank) multigraph with slots_per_var colours
  edge_order.assign(N, std::vector<int>(slots_per_var, -1));
  std::vector<std::vector<int>> var_edges(N), members(n_groups);
  for (int e = 0; e < E; e++) var_edges[col_idx[e]].push_back(e);
  for (int v = 0; v < N; v++) members[group_of_var[v]].push_back(v);
  int excess = 0;
  for (int g = 0; g < n_groups; g++) {
    const int C = slots_per_var;
    const auto &vs = members[g];
    const int nv = (int)vs.size();
    struct Ed { int vi, vb, id, col; };
    std::vector<Ed> es;
    std::vector<int> seen(B, 0);
    const int per_vbank = std::max(1, cap[g]);
    for (int i = 0; i < nv; i++)
      for (int e : var_edges[vs[i]]) {
        const int b = ebank(e);
        es.push_back({i, b * 64 + seen[b] / per_vbank, e, -1});  // overloaded banks spill into virtual banks
        seen[b]++;
      }
    std::map<int, int> vb_index;
    for (auto &ed : es) {
      auto it = vb_index.find(ed.vb);
      if (it == vb_index.end()) it = vb_index.emplace(ed.vb, (int)vb_index.size()).first;
      ed.vb = it->second;
    }
    const int nb = (int)vb_index.size();
    std::vector<std::vector<int>> at_var(nv, std::vector<int>(C, -1)), at_bank(nb, std::vector<int>(C, -1));
    for (size_t x = 0; x < es.size(); x++) {
      Ed &ed = es[x];
      int cv = -1, cb = -1, both = -1;
      for (int c = 0; c < C; c++) {
        const bool fv = at_var[ed.vi][c] < 0, fb = at_bank[ed.vb][c] < 0;
        if (fv && fb && both < 0) both = c;
        if (fv && cv < 0) cv = c;
        if (fb && cb < 0) cb = c;
      }
      if (both < 0) {
        // cv is free at the variable but taken at the bank, cb the other way round: flip the cv/cb alternating path that
        // starts at the bank; the graph is bipartite, so the path cannot end at the variable and cv becomes free at both.
        int at_b = 1, node = ed.vb, want = cv, other = cb;
        std::vector<int> path;
        while (path.size() <= es.size()) {
          const int nxt = at_b ? at_bank[node][want] : at_var[node][want];
          if (nxt < 0) break;
          path.push_back(nxt);
          node = at_b ? es[nxt].vi : es[nxt].vb;
          at_b ^= 1;
          std::swap(want, other);
        }
        for (int idx : path) {
          at_var[es[idx].vi][es[idx].col] = -1;
          at_bank[es[idx].vb][es[idx].col] = -1;
        }
        for (int idx : path) {
          es[idx].col = es[idx].col == cv ? cb : cv;
          at_var[es[idx].vi][es[idx].col] = idx;
          at_bank[es[idx].vb][es[idx].col] = idx;
        }
        both = cv;
      }
      ed.col = both;
      at_var[ed.vi][both] = (int)x;
      at_bank[ed.vb][both] = (int)x;
    }
    for (const Ed &ed : es) edge_order[vs[ed.vi]][ed.col] = ed.id;
    for (int c = 0; c < C; c++) {  // what the hardware will see: wavefronts of gather instruction c of this group
      int cnt[32] = {0}, mx = 0, any = 0;
      for (int i = 0; i < nv; i++) {
        const int e = edge_order[vs[i]][c];
        if (e >= 0) {
          any = 1;
          mx = std::max(mx, ++cnt[ebank(e)]);
        }
      }
      if (any) excess += mx - 1;
    }
  }
  if (excess_wavefronts) *excess_wavefronts = excess;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    Cached cd;
    cd.row_ptr.assign(row_ptr, row_ptr + M + 1);
    cd.col_idx.assign(col_idx, col_idx + E);
    cd.group_of_var = group_of_var;
    std::copy(par, par + 5, cd.par);
    cd.slot = slot_of_row; cd.pos = pos_of_edge; cd.order = edge_order; cd.residual = (int)cost; cd.excess = excess;
    g_cache.emplace(key, std::move(cd));
  }
  return (int)cost;
}

}  // namespace kml
