// Host-side construction of the code and constellation descriptions (include/kmldpc_b200.h).
//
// kml_code_load reproduces, with 64-bit packed rows, the outcome of the reference's byte-matrix Gaussian elimination:
//   PEG / generic flavour : lib/lab/src/binaryldpccodec.cc:346-493  (pivots walk down the diagonal from the top-left,
//                            first column >= i that has a one in rows >= i, identity ends up on the LEFT)
//   5G flavour            : lib/lab/src/binary5gldpccodec.cc:240-391 (pivots walk up from the bottom-right, identity RIGHT)
// i.e. the same pivot choices → the same column permutation tempP, the same reduced matrix enc_h_ and the same
// permuted Tanner graph.  PEG8064 takes ~1 s here versus 25.7 s in the reference.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "kml_internal.h"

namespace kml {

static thread_local std::string g_last_error;
void set_global_error(const std::string &msg) { g_last_error = msg; }
const char *global_error() { return g_last_error.c_str(); }

namespace {

struct BitMatrix {
  int rows = 0, cols = 0, stride = 0;
  std::vector<uint64_t> w;
  BitMatrix(int r, int c) : rows(r), cols(c), stride((c + 63) / 64), w((size_t)r * ((c + 63) / 64), 0) {}
  uint64_t *row(int r) { return w.data() + (size_t)r * stride; }
  const uint64_t *row(int r) const { return w.data() + (size_t)r * stride; }
  bool get(int r, int c) const { return (row(r)[c >> 6] >> (c & 63)) & 1u; }
  void set(int r, int c) { row(r)[c >> 6] |= (uint64_t)1 << (c & 63); }
  void swap_rows(int a, int b) {
    if (a != b) std::swap_ranges(row(a), row(a) + stride, row(b));
  }
  void swap_cols(int a, int b) {
    const int wa = a >> 6, wb = b >> 6, sa = a & 63, sb = b & 63;
    for (int r = 0; r < rows; r++) {
      uint64_t *p = row(r);
      const uint64_t d = ((p[wa] >> sa) ^ (p[wb] >> sb)) & 1u;
      p[wa] ^= d << sa;
      p[wb] ^= d << sb;
    }
  }
  void xor_row(int dst, int src) {
    uint64_t *d = row(dst);
    const uint64_t *s = row(src);
    for (int t = 0; t < stride; t++) d[t] ^= s[t];
  }
};

// Returns the rank; `a` becomes the reduced matrix, `perm` the column permutation.
int eliminate(BitMatrix &a, std::vector<int32_t> &perm, bool from_bottom_right) {
  const int m = a.rows, n = a.cols;
  int rank = 0;
  auto step = [&](int i, int target_col, int col_begin, int col_end, int col_step, int row_begin, int row_end,
                  int row_step) -> bool {
    int pr = -1, pc = -1;
    for (int jj = col_begin; jj != col_end && pr < 0; jj += col_step)
      for (int ii = row_begin; ii != row_end; ii += row_step)
        if (a.get(ii, jj)) {
          pr = ii;
          pc = jj;
          break;
        }
    if (pr < 0) return false;
    rank++;
    a.swap_rows(i, pr);
    if (pc != target_col) {
      std::swap(perm[target_col], perm[pc]);
      a.swap_cols(target_col, pc);
    }
    const int wq = target_col >> 6, sq = target_col & 63;
    for (int r = 0; r < m; r++)
      if (r != i && ((a.row(r)[wq] >> sq) & 1u)) a.xor_row(r, i);
    return true;
  };
  if (!from_bottom_right) {
    for (int i = 0; i < m; i++)
      if (!step(i, i, i, n, +1, i, m, +1)) break;
  } else {
    for (int i = m - 1; i >= 0; --i)
      if (!step(i, i + n - m, i + n - m, -1, -1, i, -1, -1)) break;
  }
  return rank;
}

struct CodeOwner {
  kml_code pub{};
  std::vector<int32_t> row_ptr, col_idx, perm;
  std::vector<uint32_t> enc_rows;
};

struct ModemOwner {
  kml_modem pub{};
  std::vector<double> pts;
};

}  // namespace
}  // namespace kml

using namespace kml;

extern "C" int kml_code_load(const char *h_file, int is_5g, int encoder_active, kml_code **out) {
  if (!h_file || !out) return KML_ERR_ARG;
  *out = nullptr;
  FILE *fp = fopen(h_file, "r");
  if (!fp) {
    set_global_error(std::string("cannot open parity-check file ") + h_file);
    return KML_ERR_IO;
  }
  char tok[256];
  int m = 0, n = 0, chk = 0, z = 0;
  bool ok = fscanf(fp, "%255s", tok) == 1;
  if (ok) ok = is_5g ? fscanf(fp, "%d %d %d %d", &m, &n, &chk, &z) == 4 : fscanf(fp, "%d %d %d", &m, &n, &chk) == 3;
  ok = ok && fscanf(fp, "%255s", tok) == 1 && m > 0 && n > m && n < (1 << 20);
  std::vector<std::vector<int32_t>> rows(ok ? m : 0);
  for (int i = 0; ok && i < m; i++) {
    int row_no, deg;
    ok = fscanf(fp, "%d %d", &row_no, &deg) == 2 && deg >= 0 && deg <= n;
    for (int j = 0; ok && j < deg; j++) {
      int c;
      ok = fscanf(fp, "%d", &c) == 1 && c >= 0 && c < n;
      if (ok) rows[i].push_back(c);
    }
  }
  fclose(fp);
  if (!ok) {
    set_global_error(std::string("malformed parity-check file ") + h_file);
    return KML_ERR_IO;
  }
  auto *o = new CodeOwner();
  o->perm.resize(n);
  for (int j = 0; j < n; j++) o->perm[j] = j;
  int rank = chk;
  if (encoder_active) {
    BitMatrix a(m, n);
    for (int i = 0; i < m; i++)
      for (int c : rows[i]) a.set(i, c);
    rank = eliminate(a, o->perm, is_5g != 0);
    std::vector<int32_t> inv(n);
    for (int j = 0; j < n; j++) inv[o->perm[j]] = j;
    // permuted graph: column c of the file becomes column inv[c]; the dense rebuild de-duplicates entries
    for (int i = 0; i < m; i++) {
      for (auto &c : rows[i]) c = inv[c];
      std::sort(rows[i].begin(), rows[i].end());
      rows[i].erase(std::unique(rows[i].begin(), rows[i].end()), rows[i].end());
    }
    // parity part of the reduced matrix, re-indexed by information bit
    const int k = n - rank;
    const int words = (k + 31) / 32;
    o->enc_rows.assign((size_t)rank * words, 0u);
    const int info0 = is_5g ? 0 : rank;  // PEG: info bits sit right of the identity; 5G: left of it
    for (int t = 0; t < rank; t++)
      for (int j = 0; j < k; j++)
        if (a.get(t, info0 + j)) o->enc_rows[(size_t)t * words + (j >> 5)] |= 1u << (j & 31);
    o->pub.enc_words = words;
  }
  // NOTE (!encoder_active): the reference keeps the file's graph as is, duplicates included
  // (binaryldpccodec.cc:103-124); the row order inside a row is irrelevant to the fp32 kernels.
  o->row_ptr.assign(m + 1, 0);
  for (int i = 0; i < m; i++) {
    if (!encoder_active) std::sort(rows[i].begin(), rows[i].end());
    o->row_ptr[i + 1] = o->row_ptr[i] + (int32_t)rows[i].size();
    o->col_idx.insert(o->col_idx.end(), rows[i].begin(), rows[i].end());
  }
  kml_code &p = o->pub;
  p.n_rows = m;
  p.n_graph = n;
  p.puncture = 2 * z;
  p.n_tx = n - 2 * z;
  p.n_chk = rank;
  p.k = n - rank;
  p.info_offset = is_5g ? 0 : rank;
  p.n_edges = (int32_t)o->col_idx.size();
  p.is_5g = is_5g ? 1 : 0;
  p.encoder_active = encoder_active ? 1 : 0;
  p.row_ptr = o->row_ptr.data();
  p.col_idx = o->col_idx.data();
  p.perm = o->perm.data();
  p.enc_rows = encoder_active ? o->enc_rows.data() : nullptr;
  *out = &o->pub;  // pub is the first member: the owner is recovered by a cast in kml_code_free
  return KML_OK;
}

namespace kml {
const char *knob(const char *name) {
  const char *v = getenv(name);
  if (!v || !*v) return nullptr;
  static std::mutex mu;
  static std::set<std::string> said;
  std::lock_guard<std::mutex> lk(mu);
  if (said.insert(std::string(name) + "=" + v).second)
    fprintf(stderr, "kmldpc_b200: environment knob %s=%s is active (selects a non-default kernel variant)\n", name, v);
  return v;
}
}  // namespace kml

extern "C" void kml_code_free(kml_code *code) {
  if (code) delete reinterpret_cast<CodeOwner *>(code);
}

extern "C" int kml_modem_load(const char *modem_file, kml_modem **out) {
  if (!modem_file || !out) return KML_ERR_ARG;
  *out = nullptr;
  FILE *fp = fopen(modem_file, "r");
  if (!fp) {
    set_global_error(std::string("cannot open constellation file ") + modem_file);
    return KML_ERR_IO;
  }
  char tok[512];
  int bits = 0, dims = 0;
  bool ok = fscanf(fp, "%511s %d %511s %d %511s", tok, &bits, tok, &dims, tok) == 5 && bits >= 1 && bits <= 10;
  auto *o = new ModemOwner();
  const int q = ok ? (1 << bits) : 0;
  o->pts.resize(2 * (size_t)q);
  double energy = 0.0;
  for (int i = 0; ok && i < q; i++) {
    int dec, label = 0, b;
    ok = fscanf(fp, "%d", &dec) == 1;
    for (int j = 0; ok && j < bits; j++) {
      ok = fscanf(fp, "%d", &b) == 1 && (b == 0 || b == 1);
      label = (label << 1) | b;
    }
    ok = ok && dec == label && dec == i;  // modem.cc:113
    double re, im;
    ok = ok && fscanf(fp, "%lf %lf", &re, &im) == 2;
    if (ok) {
      o->pts[2 * i] = re;
      o->pts[2 * i + 1] = im;
      const double a = std::hypot(re, im);
      energy += a * a;
    }
  }
  fclose(fp);
  if (!ok) {
    delete o;
    set_global_error(std::string("malformed constellation file ") + modem_file);
    return KML_ERR_IO;
  }
  energy /= q;
  const double s = std::sqrt(energy);
  for (auto &v : o->pts) v /= s;
  o->pub.bits_per_symbol = bits;
  o->pub.n_points = q;
  o->pub.points = o->pts.data();
  *out = &o->pub;
  return KML_OK;
}

namespace kml {
// Voronoi neighbours of constellation point 0: the only points whose bisector bounds the cell of s_0.  The k-means
// kernel needs just "is the nearest centroid cluster 0?" (kmeans.cc:36-46 feeds only cluster 0 back), and the centroids
// are always the constellation scaled and rotated by one complex number, so the neighbour set never changes.
// Half-plane clipping of a large square; a point is kept when its half-plane cuts area off the cell built from all the
// others (a bisector that only touches the cell at a vertex never decides a sample: there d_0 = d_k exactly).
std::vector<int> voronoi_neighbours_of_first(const double *pts, int q) {
  struct P { double x, y; };
  const P s0{pts[0], pts[1]};
  double scale = 1.0;
  for (int k = 0; k < q; k++) scale = std::max(scale, std::hypot(pts[2 * k], pts[2 * k + 1]));
  const double R = 1.0e3 * scale, eps = 1.0e-9 * scale * scale;
  auto clip = [&](std::vector<P> poly, int j) {  // keep { p : (p - mid) . (s_j - s_0) <= 0 }
    const double dx = pts[2 * j] - s0.x, dy = pts[2 * j + 1] - s0.y;
    const double mx = 0.5 * (pts[2 * j] + s0.x), my = 0.5 * (pts[2 * j + 1] + s0.y);
    auto side = [&](const P &p) { return (p.x - mx) * dx + (p.y - my) * dy; };
    std::vector<P> out;
    for (size_t i = 0; i < poly.size(); i++) {
      const P a = poly[i], b = poly[(i + 1) % poly.size()];
      const double sa = side(a), sb = side(b);
      if (sa <= 0) out.push_back(a);
      if ((sa < 0 && sb > 0) || (sa > 0 && sb < 0)) {
        const double t = sa / (sa - sb);
        out.push_back({a.x + t * (b.x - a.x), a.y + t * (b.y - a.y)});
      }
    }
    return out;
  };
  std::vector<int> nb;
  for (int k = 1; k < q; k++) {
    const double dx = pts[2 * k] - s0.x, dy = pts[2 * k + 1] - s0.y;
    if (dx * dx + dy * dy < 1e-24) continue;  // coincides with s_0: never strictly closer
    std::vector<P> poly = {{s0.x - R, s0.y - R}, {s0.x + R, s0.y - R}, {s0.x + R, s0.y + R}, {s0.x - R, s0.y + R}};
    for (int j = 1; j < q && !poly.empty(); j++)
      if (j != k) {
        const double ex = pts[2 * j] - s0.x, ey = pts[2 * j + 1] - s0.y;
        if (ex * ex + ey * ey >= 1e-24) poly = clip(poly, j);
      }
    const double mx = 0.5 * (pts[2 * k] + s0.x), my = 0.5 * (pts[2 * k + 1] + s0.y);
    bool touches = false;
    for (const P &v : poly) touches = touches || ((v.x - mx) * dx + (v.y - my) * dy > eps);
    if (touches) nb.push_back(k);
  }
  return nb;
}
}  // namespace kml

extern "C" void kml_modem_free(kml_modem *modem) {
  if (modem) delete reinterpret_cast<ModemOwner *>(modem);
}
