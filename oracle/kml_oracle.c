/* TEST INFRASTRUCTURE ONLY — see kml_oracle.h.  CPU (fp64) restatement of the reference's link path.
 * Array/CSR based; arithmetic is kept in the reference's order so that results can be compared bit for bit
 * with dumps of the unmodified reference classes (oracle/ref/ref_harness.cc).  Compile with
 * -ffp-contract=off (oracle/Makefile) so no FMA contraction changes the rounding.
 * Citations are relative to /root/reference/kmldpc.
 */
#define _GNU_SOURCE
#include "kml_oracle.h"

#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define KMO_SMALLEST_PROB 1.0e-12       /* lib/lab/include/utility.h:12 */
#define KMO_PI 3.14159265358979         /* lib/lab/include/utility.h:10 (truncated on purpose) */
#define KMO_SQRT2 1.4142135623730950488016 /* utility.h:16 */

/* ------------------------------------------------------------------------------------------------ RNG */
/* Park–Miller LCG with Schrage's trick: lib/lab/src/randnum.cc:4-6,36-45 */
enum { LCG_A = 48271, LCG_Q = 2147483647 / 48271, LCG_R = 2147483647 % 48271 };
#define LCG_M 2147483647L

void kmo_lcg_seed(kmo_lcg *g, long state) { g->state = state; }

double kmo_lcg_uniform(kmo_lcg *g) {
  int t = (int)(LCG_A * (g->state % LCG_Q) - LCG_R * (g->state / LCG_Q));
  g->state = (t >= 0) ? t : t + LCG_M;
  return g->state / (double)LCG_M;
}

/* Marsaglia polar method, pairs: randnum.cc:48-72 */
void kmo_lcg_normal(kmo_lcg *g, double *nn, int len) {
  double x1 = 0, x2 = 0, w;
  for (int t = 0; 2 * t + 1 < len; t++) {
    w = 2.0;
    while (w > 1.0) {
      x1 = 2.0 * kmo_lcg_uniform(g) - 1.0;
      x2 = 2.0 * kmo_lcg_uniform(g) - 1.0;
      w = x1 * x1 + x2 * x2;
    }
    w = sqrt(-2.0 * log(w) / w);
    nn[2 * t] = x1 * w;
    nn[2 * t + 1] = x2 * w;
  }
  if (len % 2 == 1) {
    w = 2.0;
    while (w > 1.0) {
      x1 = 2.0 * kmo_lcg_uniform(g) - 1.0;
      x2 = 2.0 * kmo_lcg_uniform(g) - 1.0;
      w = x1 * x1 + x2 * x2;
    }
    w = sqrt(-2.0 * log(w) / w);
    nn[len - 1] = x1 * w;
  }
}

/* ------------------------------------------------------------------------------------------------ code */
struct kmo_code {
  int m, n, chk, k, n_tx, two_z, e, is_5g, active;
  /* adjacency in the reference object's traversal order; edge id = position in the row CSR */
  int32_t *row_ptr, *row_col;  /* rows: `right` walk from row_head (descending column after GE) */
  int32_t *col_ptr, *col_edge; /* columns: `down` walk from col_head (descending row); values are edge ids */
  int32_t *edge_row;
  int32_t *perm;               /* tempP: new column j holds original column perm[j] */
  uint64_t *enc;               /* reduced matrix, bit-packed rows, stride words (NULL when !active) */
  int stride;
};

static inline int bit_get(const uint64_t *row, int j) { return (int)((row[j >> 6] >> (j & 63)) & 1u); }
static inline void bit_flip(uint64_t *row, int j) { row[j >> 6] ^= (uint64_t)1 << (j & 63); }

/* Head insertion into the doubly linked row/column lists (binaryldpccodec.cc:106-122, :467-481) means both
 * walks visit edges in REVERSE insertion order.  `ins` lists (row, col) in insertion order. */
static void build_graph(kmo_code *c, const int32_t *ins_row, const int32_t *ins_col, int e) {
  c->e = e;
  c->row_ptr = calloc(c->m + 1, sizeof(int32_t));
  c->col_ptr = calloc(c->n + 1, sizeof(int32_t));
  c->row_col = malloc(sizeof(int32_t) * e);
  c->col_edge = malloc(sizeof(int32_t) * e);
  c->edge_row = malloc(sizeof(int32_t) * e);
  for (int i = 0; i < e; i++) {
    c->row_ptr[ins_row[i] + 1]++;
    c->col_ptr[ins_col[i] + 1]++;
  }
  for (int r = 0; r < c->m; r++) c->row_ptr[r + 1] += c->row_ptr[r];
  for (int v = 0; v < c->n; v++) c->col_ptr[v + 1] += c->col_ptr[v];
  int32_t *rfill = malloc(sizeof(int32_t) * c->m), *cfill = malloc(sizeof(int32_t) * c->n);
  int32_t *eid = malloc(sizeof(int32_t) * e);
  for (int r = 0; r < c->m; r++) rfill[r] = c->row_ptr[r + 1];
  for (int v = 0; v < c->n; v++) cfill[v] = c->col_ptr[v + 1];
  for (int i = 0; i < e; i++) { /* fill from the back: last inserted is visited first */
    int pos = --rfill[ins_row[i]];
    c->row_col[pos] = ins_col[i];
    c->edge_row[pos] = ins_row[i];
    eid[i] = pos;
  }
  for (int i = 0; i < e; i++) c->col_edge[--cfill[ins_col[i]]] = eid[i];
  free(rfill);
  free(cfill);
  free(eid);
}

/* Gaussian elimination with column swaps.  PEG flavour: binaryldpccodec.cc:383-428 (pivots top-left, identity on
 * the LEFT).  5G flavour: binary5gldpccodec.cc:277-322 (pivots bottom-right, identity on the RIGHT). */
static void reduce_h(kmo_code *c, uint64_t *h) {
  const int m = c->m, n = c->n, w = c->stride;
  c->chk = 0;
  if (!c->is_5g) {
    for (int i = 0; i < m; i++) {
      int ii = 0, jj, found = 0;
      for (jj = i; jj < n; jj++) {
        for (ii = i; ii < m; ii++)
          if (bit_get(h + (size_t)ii * w, jj)) { found = 1; break; }
        if (found) { c->chk++; break; }
      }
      if (!found) break;
      if (ii != i)
        for (int t = 0; t < w; t++) { uint64_t x = h[(size_t)i * w + t]; h[(size_t)i * w + t] = h[(size_t)ii * w + t]; h[(size_t)ii * w + t] = x; }
      if (jj != i) {
        int32_t x = c->perm[i]; c->perm[i] = c->perm[jj]; c->perm[jj] = x;
        for (int r = 0; r < m; r++) {
          uint64_t *row = h + (size_t)r * w;
          if (bit_get(row, i) != bit_get(row, jj)) { bit_flip(row, i); bit_flip(row, jj); }
        }
      }
      const uint64_t *piv = h + (size_t)i * w;
      for (int r = 0; r < m; r++)
        if (r != i && bit_get(h + (size_t)r * w, i))
          for (int t = 0; t < w; t++) h[(size_t)r * w + t] ^= piv[t];
    }
  } else {
    for (int i = m - 1; i >= 0; --i) {
      const int pc = i + n - m;
      int ii = 0, jj, found = 0;
      for (jj = pc; jj >= 0; --jj) {
        for (ii = i; ii >= 0; --ii)
          if (bit_get(h + (size_t)ii * w, jj)) { found = 1; break; }
        if (found) { c->chk++; break; }
      }
      if (!found) break;
      if (ii != i)
        for (int t = 0; t < w; t++) { uint64_t x = h[(size_t)i * w + t]; h[(size_t)i * w + t] = h[(size_t)ii * w + t]; h[(size_t)ii * w + t] = x; }
      if (jj != pc) {
        int32_t x = c->perm[pc]; c->perm[pc] = c->perm[jj]; c->perm[jj] = x;
        for (int r = 0; r < m; r++) {
          uint64_t *row = h + (size_t)r * w;
          if (bit_get(row, pc) != bit_get(row, jj)) { bit_flip(row, pc); bit_flip(row, jj); }
        }
      }
      const uint64_t *piv = h + (size_t)i * w;
      for (int r = m - 1; r >= 0; --r)
        if (r != i && bit_get(h + (size_t)r * w, pc))
          for (int t = 0; t < w; t++) h[(size_t)r * w + t] ^= piv[t];
    }
  }
}

/* File format + construction: binaryldpccodec.cc:62-129, binary5gldpccodec.cc:12-80; rebuild after the column
 * permutation: binaryldpccodec.cc:440-483 */
kmo_code *kmo_code_load(const char *h_file, int is_5g, int encoder_active) {
  FILE *fp = fopen(h_file, "r");
  if (!fp) return NULL;
  kmo_code *c = calloc(1, sizeof *c);
  char tok[256];
  int z = 0;
  c->is_5g = is_5g;
  c->active = encoder_active;
  if (fscanf(fp, "%255s", tok) != 1) goto bad;
  if (is_5g) {
    if (fscanf(fp, "%d %d %d %d", &c->m, &c->n, &c->chk, &z) != 4) goto bad;
  } else {
    if (fscanf(fp, "%d %d %d", &c->m, &c->n, &c->chk) != 3) goto bad;
  }
  c->two_z = 2 * z;
  c->n_tx = c->n - c->two_z;
  c->k = c->n - c->chk;
  if (fscanf(fp, "%255s", tok) != 1) goto bad;
  int cap = 16 * c->m, e = 0;
  int32_t *ir = malloc(sizeof(int32_t) * cap), *ic = malloc(sizeof(int32_t) * cap);
  for (int i = 0; i < c->m; i++) {
    int row_no, deg;
    if (fscanf(fp, "%d %d", &row_no, &deg) != 2) goto bad;
    for (int j = 0; j < deg; j++) {
      int col;
      if (fscanf(fp, "%d", &col) != 1) goto bad;
      if (e == cap) { cap *= 2; ir = realloc(ir, sizeof(int32_t) * cap); ic = realloc(ic, sizeof(int32_t) * cap); }
      ir[e] = i;
      ic[e] = col;
      e++;
    }
  }
  fclose(fp);
  c->perm = malloc(sizeof(int32_t) * c->n);
  for (int j = 0; j < c->n; j++) c->perm[j] = j;
  c->stride = (c->n + 63) / 64;
  if (!encoder_active) {
    build_graph(c, ir, ic, e);
  } else {
    uint64_t *orig = calloc((size_t)c->m * c->stride, sizeof(uint64_t));
    for (int i = 0; i < e; i++) orig[(size_t)ir[i] * c->stride + (ic[i] >> 6)] |= (uint64_t)1 << (ic[i] & 63);
    c->enc = malloc((size_t)c->m * c->stride * sizeof(uint64_t));
    memcpy(c->enc, orig, (size_t)c->m * c->stride * sizeof(uint64_t));
    reduce_h(c, c->enc);
    /* dec_h[i][j] = H[i][perm[j]], rows NOT permuted; edges inserted row-major ascending */
    int e2 = 0;
    for (int i = 0; i < c->m; i++)
      for (int j = 0; j < c->n; j++)
        if (bit_get(orig + (size_t)i * c->stride, c->perm[j])) {
          if (e2 == cap) { cap *= 2; ir = realloc(ir, sizeof(int32_t) * cap); ic = realloc(ic, sizeof(int32_t) * cap); }
          ir[e2] = i;
          ic[e2] = j;
          e2++;
        }
    build_graph(c, ir, ic, e2);
    free(orig);
    c->k = c->n - c->chk; /* binaryldpccodec.cc:484-485 */
  }
  free(ir);
  free(ic);
  return c;
bad:
  fclose(fp);
  free(c);
  return NULL;
}

void kmo_code_free(kmo_code *c) {
  if (!c) return;
  free(c->row_ptr); free(c->row_col); free(c->col_ptr); free(c->col_edge); free(c->edge_row); free(c->perm); free(c->enc);
  free(c);
}

void kmo_code_info(const kmo_code *c, int32_t info[8]) {
  info[0] = c->m; info[1] = c->n; info[2] = c->n_tx; info[3] = c->k;
  info[4] = c->chk; info[5] = c->two_z; info[6] = c->e; info[7] = c->active;
}

void kmo_code_export(const kmo_code *c, int32_t *row_ptr, int32_t *col_idx, int32_t *col_ptr, int32_t *row_idx,
                     int32_t *perm, uint8_t *enc_h) {
  if (row_ptr) memcpy(row_ptr, c->row_ptr, sizeof(int32_t) * (c->m + 1));
  if (col_idx) memcpy(col_idx, c->row_col, sizeof(int32_t) * c->e);
  if (col_ptr) memcpy(col_ptr, c->col_ptr, sizeof(int32_t) * (c->n + 1));
  if (row_idx)
    for (int i = 0; i < c->e; i++) row_idx[i] = c->edge_row[c->col_edge[i]];
  if (perm) memcpy(perm, c->perm, sizeof(int32_t) * c->n);
  if (enc_h && c->enc)
    for (int r = 0; r < c->m; r++)
      for (int j = 0; j < c->n; j++) enc_h[(size_t)r * c->n + j] = (uint8_t)bit_get(c->enc + (size_t)r * c->stride, j);
}

/* binaryldpccodec.cc:144-162 (info on the right of the permuted word); binary5gldpccodec.cc:86-109 */
void kmo_encode(const kmo_code *c, int *uu, int *cc) {
  if (!c->active) {
    for (int i = 0; i < c->k; i++) uu[i] = 0;
    for (int i = 0; i < c->n_tx; i++) cc[i] = 0;
    return;
  }
  if (!c->is_5g) {
    for (int t = c->chk; t < c->n; t++) cc[t] = uu[t - c->chk];
    for (int t = 0; t < c->chk; t++) {
      const uint64_t *row = c->enc + (size_t)t * c->stride;
      int p = 0;
      for (int j = c->chk; j < c->n; j++) p ^= (cc[j] & bit_get(row, j));
      cc[t] = p;
    }
  } else {
    int *np = malloc(sizeof(int) * c->n);
    for (int t = 0; t < c->k; t++) np[t] = uu[t];
    for (int t = 0; t < c->chk; t++) {
      const uint64_t *row = c->enc + (size_t)t * c->stride;
      int p = 0;
      for (int j = 0; j < c->k; j++) p ^= (np[j] & bit_get(row, j));
      np[c->k + t] = p;
    }
    for (int t = 0; t < c->n_tx; t++) cc[t] = np[t + c->two_z];
    free(np);
  }
}

/* binaryldpccodec.cc:281-299 */
int kmo_parity_check(const kmo_code *c, const int *rr) {
  int count = 0;
  for (int r = 0; r < c->m; r++) {
    int p = 0;
    for (int t = c->row_ptr[r]; t < c->row_ptr[r + 1]; t++) p ^= rr[c->row_col[t]];
    count += (p != 0);
  }
  return count;
}

/* Flooding sum-product in the probability domain: binaryldpccodec.cc:165-278, binary5gldpccodec.cc:112-232.
 * Messages are pairs (P(bit=0), P(bit=1)); every product is renormalised like the reference does. */
#define KMO_MAX_DEG 64
int kmo_decode(const kmo_code *c, const double *p0, int iter_count, int max_iter, int *uu_hat, int *cc_hat,
               double *syndrom_soft) {
  const int e = c->e;
  double *c2v0 = malloc(sizeof(double) * e), *c2v1 = malloc(sizeof(double) * e);
  double *v2c0 = malloc(sizeof(double) * e), *v2c1 = malloc(sizeof(double) * e);
  double a0[KMO_MAX_DEG + 1], a1[KMO_MAX_DEG + 1];
  for (int i = 0; i < e; i++) c2v0[i] = c2v1[i] = 0.5; /* InitMsg, :302-313 */
  for (int i = 0; i < e; i++) v2c0[i] = v2c1[i] = 0.5;
  int iter;
  for (iter = 0; iter < iter_count; iter++) {
    /* variable nodes: forward chain over the `down` walk, hard decision, backward chain (:177-212) */
    for (int v = 0; v < c->n; v++) {
      const int32_t *ed = c->col_edge + c->col_ptr[v];
      const int d = c->col_ptr[v + 1] - c->col_ptr[v];
      if (v < c->two_z) { a0[0] = 0.5; a1[0] = 1.0 - 0.5; }
      else { a0[0] = p0[v - c->two_z]; a1[0] = 1.0 - p0[v - c->two_z]; }
      for (int t = 0; t < d; t++) {
        double x0 = a0[t] * c2v0[ed[t]], x1 = a1[t] * c2v1[ed[t]];
        double s = x0 + x1;
        a0[t + 1] = x0 / s;
        a1[t + 1] = x1 / s;
      }
      cc_hat[v] = (a0[d] > a1[d]) ? 0 : 1;
      double b0 = 1.0, b1 = 1.0;
      for (int t = d - 1; t >= 0; t--) {
        double t0 = a0[t] * b0, t1 = a1[t] * b1;
        double s = t0 + t1;
        v2c0[ed[t]] = t0 / s;
        v2c1[ed[t]] = t1 / s;
        double n0 = b0 * c2v0[ed[t]], n1 = b1 * c2v1[ed[t]];
        s = n0 + n1;
        b0 = n0 / s;
        b1 = n1 / s;
      }
    }
    if (c->is_5g) for (int i = 0; i < c->k; i++) uu_hat[i] = cc_hat[i];           /* 5G :167-170 */
    else          for (int i = 0; i < c->k; i++) uu_hat[i] = cc_hat[i + c->chk];  /* :214-216 */
    /* syndrome, stop before the check-node phase when it is zero (:217-232) */
    int success = 1;
    for (int r = 0; r < c->m && success; r++) {
      int p = 0;
      for (int t = c->row_ptr[r]; t < c->row_ptr[r + 1]; t++) p ^= cc_hat[c->row_col[t]];
      if (p) success = 0;
    }
    if (success) break;
    /* check nodes: 2-state trellis forward/backward over the `right` walk, clip c2v (:235-275) */
    for (int r = 0; r < c->m; r++) {
      const int base = c->row_ptr[r];
      const int d = c->row_ptr[r + 1] - base;
      a0[0] = 1.0;
      a1[0] = 0.0;
      for (int t = 0; t < d; t++) {
        double x0 = a0[t] * v2c0[base + t] + a1[t] * v2c1[base + t];
        double x1 = a0[t] * v2c1[base + t] + a1[t] * v2c0[base + t];
        double s = x0 + x1;
        a0[t + 1] = x0 / s;
        a1[t + 1] = x1 / s;
      }
      double b0 = 1.0, b1 = 0.0;
      for (int t = d - 1; t >= 0; t--) {
        double t0 = a0[t] * b0 + a1[t] * b1;
        double t1 = a0[t] * b1 + a1[t] * b0;
        double s = t0 + t1;
        double q0 = t0 / s;
        if (q0 > 1.0 - KMO_SMALLEST_PROB) q0 = 1.0 - KMO_SMALLEST_PROB;
        if (q0 < KMO_SMALLEST_PROB) q0 = KMO_SMALLEST_PROB;
        c2v0[base + t] = q0;
        c2v1[base + t] = 1.0 - q0;
        double n0 = b0 * v2c0[base + t] + b1 * v2c1[base + t];
        double n1 = b0 * v2c1[base + t] + b1 * v2c0[base + t];
        s = n0 + n1;
        b0 = n0 / s;
        b1 = n1 / s;
      }
      if (syndrom_soft) syndrom_soft[r] = a0[d];
    }
  }
  free(c2v0); free(c2v1); free(v2c0); free(v2c1);
  return iter + (iter < max_iter);
}

/* ------------------------------------------------------------------------------------------------ modem */
struct kmo_modem {
  int bits, q;
  double complex *pts;
};

/* lib/lab/src/modem.cc:87-129 */
kmo_modem *kmo_modem_load(const char *modem_file) {
  FILE *fp = fopen(modem_file, "r");
  if (!fp) return NULL;
  kmo_modem *m = calloc(1, sizeof *m);
  char tok[512];
  int out_len;
  if (fscanf(fp, "%511s %d %511s %d %511s", tok, &m->bits, tok, &out_len, tok) != 5) { fclose(fp); free(m); return NULL; }
  m->q = 1 << m->bits;
  m->pts = malloc(sizeof(double complex) * m->q);
  double energy = 0;
  for (int i = 0; i < m->q; i++) {
    int dec, lab = 0, b;
    if (fscanf(fp, "%d", &dec) != 1) goto bad;
    for (int j = 0; j < m->bits; j++) {
      if (fscanf(fp, "%d", &b) != 1) goto bad;
      lab = (lab << 1) + b;
    }
    if (dec != lab || dec != i) goto bad;
    double re, im;
    if (fscanf(fp, "%lf %lf", &re, &im) != 2) goto bad;
    m->pts[i] = re + im * I;
    energy += pow(cabs(m->pts[i]), 2);
  }
  fclose(fp);
  energy /= m->q;
  for (int i = 0; i < m->q; i++) m->pts[i] /= sqrt(energy);
  return m;
bad:
  fclose(fp);
  free(m->pts);
  free(m);
  return NULL;
}

void kmo_modem_free(kmo_modem *m) { if (m) { free(m->pts); free(m); } }
void kmo_modem_info(const kmo_modem *m, int32_t info[2]) { info[0] = m->bits; info[1] = m->q; }
void kmo_modem_points(const kmo_modem *m, double *re_im) { memcpy(re_im, m->pts, sizeof(double) * 2 * m->q); }

/* modem.cc:12-20 */
void kmo_map(const kmo_modem *m, const int *cc, int n_sym, double *xx) {
  for (int i = 0; i < n_sym; i++) {
    int idx = 0;
    for (int j = 0; j < m->bits; j++) idx = (idx << 1) + cc[j + i * m->bits];
    xx[2 * i] = creal(m->pts[idx]);
    xx[2 * i + 1] = cimag(m->pts[idx]);
  }
}

/* modemlinearsystem.cc:37-48 with a single h; noise drawn symbol by symbol (randnum.cc:75-87) */
void kmo_channel(kmo_lcg *g, const double *xx, int n_sym, double h_re, double h_im, double sigma, double *yy) {
  double *noise = malloc(sizeof(double) * 2 * n_sym);
  for (int j = 0; j < n_sym; j++) kmo_lcg_normal(g, noise + 2 * j, 2);
  const double s = sigma / KMO_SQRT2;
  for (int j = 0; j < n_sym; j++) {
    double tr = xx[2 * j] * h_re - xx[2 * j + 1] * h_im;
    double ti = xx[2 * j] * h_im + xx[2 * j + 1] * h_re;
    yy[2 * j] = tr + (noise[2 * j] * s - noise[2 * j + 1] * 0.0);
    yy[2 * j + 1] = ti + (noise[2 * j] * 0.0 + noise[2 * j + 1] * s);
  }
  free(noise);
}

static inline double prob_clip(double x) { /* lib/lab/src/utility.cc:19-27 */
  if (x < KMO_SMALLEST_PROB) return KMO_SMALLEST_PROB;
  if (x > 1.0 - KMO_SMALLEST_PROB) return 1.0 - KMO_SMALLEST_PROB;
  return x;
}

/* modemlinearsystem.cc:51-79 (per-symbol softmax + clip) then modem.cc:23-79 with all bit priors = 0.5
 * (kmcodec.cc:92-103).  Output P(bit = 0), MSB first within a symbol. */
void kmo_demap(const kmo_modem *m, const double *yy, int n_sym, double h_re, double h_im, double var, double *p0) {
  const int q = m->q, bits = m->bits;
  double *sp = malloc(sizeof(double) * q);
  for (int i = 0; i < n_sym; i++) {
    for (int k = 0; k < q; k++) {
      double sr = creal(m->pts[k]), si = cimag(m->pts[k]);
      double pr = sr * h_re - si * h_im, pi = sr * h_im + si * h_re;
      pr -= yy[2 * i];
      pi -= yy[2 * i + 1];
      sp[k] = -((pr * pr + pi * pi) / var);
    }
    double mx = sp[0];
    for (int k = 1; k < q; k++) if (sp[k] > mx) mx = sp[k];
    double sum = 0.0;
    for (int k = 0; k < q; k++) sp[k] = exp(sp[k] - mx);
    for (int k = 0; k < q; k++) sum += sp[k];
    for (int k = 0; k < q; k++) sp[k] = prob_clip(sp[k] / sum);
    /* bit-level: prior product 0.5^bits (exact), times symbol likelihood, renormalise (modem.cc:30-57) */
    double prior = 1.0;
    for (int j = 0; j < bits; j++) prior *= 0.5;
    sum = 0.0;
    for (int k = 0; k < q; k++) { sp[k] = prior * sp[k]; sum += sp[k]; }
    for (int k = 0; k < q; k++) sp[k] /= sum;
    for (int j = 0; j < bits; j++) {
      double z0 = 0.0, z1 = 0.0;
      for (int k = 0; k < q; k++) {
        if (((k >> (bits - 1 - j)) & 1) == 0) z0 += sp[k];
        else z1 += sp[k];
      }
      z0 /= 0.5;
      z1 /= (1.0 - 0.5);
      p0[i * bits + j] = prob_clip(z0 / (z0 + z1));
    }
  }
  free(sp);
}

/* ------------------------------------------------------------------------------------------------ k-means */
/* src/kmeans.cc:15-84 as COMPILED: clear()+operator[] leaves the accumulators cumulative and makes the
 * "largest cluster" search a no-op, so cluster 0 is always the anchor.  Returns the number of assignment passes. */
int kmo_kmeans(const double *yy, int n, const double *cons, int q, int max_iter, double *clusters) {
  const double complex *y = (const double complex *)yy;
  const double complex *s = (const double complex *)cons;
  double complex *c = (double complex *)clusters;
  double complex *prev = calloc(q, sizeof(double complex)), *sum = calloc(q, sizeof(double complex));
  int *cnt = calloc(q, sizeof(int));
  int a = 0;
  double best = cabs(y[0]);
  for (int i = 1; i < n; i++) { double v = cabs(y[i]); if (v > best) { best = v; a = i; } }
  double complex hat = y[a] / s[0];
  for (int k = 0; k < q; k++) c[k] = s[k] * hat;
  int passes = 0;
  for (int it = 0; it < max_iter; it++) {
    passes++;
    for (int j = 0; j < n; j++) {
      int mi = 0;
      double md = cabs(c[0] - y[j]);
      for (int k = 1; k < q; k++) { double d = cabs(c[k] - y[j]); if (d < md) { md = d; mi = k; } }
      cnt[mi]++;
      sum[mi] += y[j];
    }
    int same = 1;
    for (int k = 0; k < q; k++)
      if (!(creal(c[k]) == creal(prev[k]) && cimag(c[k]) == cimag(prev[k]))) { same = 0; break; }
    if (same) break;
    memcpy(prev, c, sizeof(double complex) * q);
    double complex c0 = sum[0] / ((double)cnt[0] + 0.0 * I);
    hat = c0 / s[0];
    for (int k = 0; k < q; k++) c[k] = s[k] * hat;
  }
  free(prev); free(sum); free(cnt);
  return passes;
}

/* ------------------------------------------------------------------------------------------------ frame */
void kmo_receive(const kmo_code *c, const kmo_modem *m, const kmo_opts *o, const double *yy, const double *true_h,
                 double var, kmo_frame_out *out, double *clusters, double *p0, int *cc_hat, int *uu_hat,
                 double *soft_state) {
  const int n_sym = c->n_tx / m->bits;
  double complex hh[4];
  int n_hat;
  double *cl = clusters ? clusters : malloc(sizeof(double) * 2 * m->q);
  double *pp = p0 ? p0 : malloc(sizeof(double) * c->n_tx);
  int *cch = cc_hat ? cc_hat : malloc(sizeof(int) * c->n);
  int *rr = malloc(sizeof(int) * c->n_tx);
  /* syndrom_soft_ is a MEMBER of the codec (binaryldpccodec.h:50): it survives from one Decoder call to the next and is
   * only overwritten by a check-node phase (binaryldpccodec.cc:274).  soft_state is that member; without one the frame
   * starts from all ones (the reference's array starts uninitialised, binaryldpccodec.cc:88). */
  double *soft = soft_state ? soft_state : malloc(sizeof(double) * c->m);
  if (!soft_state) for (int r = 0; r < c->m; r++) soft[r] = 1.0;
  out->hhat[0] = out->hhat[1] = 0;
  for (int k = 0; k < 4; k++) out->metric[k] = 0;
  if (o->known_h) { /* simulator.cc:132-133 */
    hh[0] = true_h[0] + true_h[1] * I;
    n_hat = 1;
    for (int k = 0; k < 2 * m->q; k++) cl[k] = 0;
  } else { /* simulator.cc:134-148 */
    kmo_kmeans(yy, n_sym, (const double *)m->pts, m->q, o->kmeans_iter, cl);
    double complex hat = (cl[0] + cl[1] * I) / m->pts[0];
    out->hhat[0] = creal(hat);
    out->hhat[1] = cimag(hat);
    for (int k = 0; k < 4; k++) hh[k] = hat * cexp(((KMO_PI / 2) * k) * I);
    n_hat = 4;
  }
  int kstar = 0;
  if (n_hat > 1) { /* kmcodec.cc:58-66,122-163 */
    for (int k = 0; k < 4; k++) {
      kmo_demap(m, yy, n_sym, creal(hh[k]), cimag(hh[k]), var, pp);
      double metric;
      if (o->metric_type) {
        kmo_decode(c, pp, o->metric_iter, o->max_iter, uu_hat, cch, soft);
        metric = 0.0;
        for (int r = 0; r < c->m; r++) metric += log(soft[r]);
      } else if (o->is_5g) {
        kmo_decode(c, pp, o->metric_iter, o->max_iter, uu_hat, cch, soft);
        metric = kmo_parity_check(c, cch);
      } else {
        for (int i = 0; i < c->n_tx; i++) rr[i] = pp[i] > 0.5 ? 1 : 0; /* inverted on purpose, kmcodec.cc:110-115 */
        metric = kmo_parity_check(c, rr);
      }
      out->metric[k] = fabs(metric);
    }
    for (int k = 1; k < 4; k++) if (out->metric[k] < out->metric[kstar]) kstar = k; /* first argmin */
  }
  out->kstar = kstar;
  if (o->histogram) { /* simulator.cc:154-162: the metrics are the product; nothing is decoded after them */
    out->ret = 0;
    if (!(n_hat > 1 && (o->metric_type || o->is_5g))) for (int t = 0; t < c->k; t++) uu_hat[t] = 0;
  } else {
    kmo_demap(m, yy, n_sym, creal(hh[kstar]), cimag(hh[kstar]), var, pp);
    out->ret = kmo_decode(c, pp, o->max_iter, o->max_iter, uu_hat, cch, soft);
  }
  if (!clusters) free(cl);
  if (!p0) free(pp);
  if (!cc_hat) free(cch);
  free(rr);
  if (!soft_state) free(soft);
}

void kmo_frame(const kmo_code *c, const kmo_modem *m, const kmo_opts *o, kmo_lcg *g, double snr_db, kmo_frame_out *out,
               int *uu, int *cc, double *yy, double *clusters, double *p0, int *cc_hat, int *uu_hat, double *soft_state) {
  const int n_sym = c->n_tx / m->bits;
  const double var = pow(10.0, -0.1 * snr_db), sigma = sqrt(var); /* simulator.cc:74-77 */
  int *u = uu ? uu : malloc(sizeof(int) * c->k);
  int *cw = cc ? cc : malloc(sizeof(int) * c->n_tx);
  double *y = yy ? yy : malloc(sizeof(double) * 2 * n_sym);
  int *uh = uu_hat ? uu_hat : malloc(sizeof(int) * c->k);
  double *xx = malloc(sizeof(double) * 2 * n_sym);
  for (int t = 0; t < c->k; t++) u[t] = (kmo_lcg_uniform(g) < 0.5 ? 0 : 1); /* sourcesink.cc:5-9 */
  kmo_encode(c, u, cw);
  double hn[2];
  kmo_lcg_normal(g, hn, 2);                                                 /* simulator.cc:121-123 */
  out->h[0] = hn[0] * sqrt(0.5);
  out->h[1] = hn[1] * sqrt(0.5);
  kmo_map(m, cw, n_sym, xx);
  kmo_channel(g, xx, n_sym, out->h[0], out->h[1], sigma, y);
  kmo_receive(c, m, o, y, out->h, var, out, clusters, p0, cc_hat, uh, soft_state);
  int ne = 0;
  for (int t = 0; t < c->k; t++) ne += (u[t] != uh[t]);                     /* sourcesink.cc:29-47 */
  out->nerr = ne;
  free(xx);
  if (!uu) free(u);
  if (!cc) free(cw);
  if (!yy) free(y);
  if (!uu_hat) free(uh);
}

typedef struct {
  const kmo_code *c; const kmo_modem *m; const kmo_opts *o;
  double snr; long seed, frames;
  uint64_t cnt[4]; int64_t iters;
} run_arg;

static void *run_thread(void *p) {
  run_arg *a = p;
  kmo_lcg g;
  kmo_lcg_seed(&g, a->seed);
  double *soft = malloc(sizeof(double) * a->c->m); /* this worker's codec copy (simulator.cc:93-97) */
  for (int r = 0; r < a->c->m; r++) soft[r] = 1.0;
  for (long f = 0; f < a->frames; f++) {
    kmo_frame_out fo;
    kmo_frame(a->c, a->m, a->o, &g, a->snr, &fo, 0, 0, 0, 0, 0, 0, 0, soft);
    a->cnt[0] += 1;
    a->cnt[1] += (fo.nerr > 0);
    a->cnt[2] += (uint64_t)a->c->k;
    a->cnt[3] += (uint64_t)fo.nerr;
    a->iters += fo.ret > a->o->max_iter ? a->o->max_iter : fo.ret;
  }
  free(soft);
  return NULL;
}

int64_t kmo_run(const kmo_code *c, const kmo_modem *m, const kmo_opts *o, double snr_db, long seed0, long frames,
                int threads, uint64_t counters[4]) {
  if (threads < 1) threads = 1;
  pthread_t *th = malloc(sizeof(pthread_t) * threads);
  run_arg *ar = calloc(threads, sizeof(run_arg));
  for (int t = 0; t < threads; t++) {
    ar[t].c = c; ar[t].m = m; ar[t].o = o; ar[t].snr = snr_db; ar[t].seed = seed0 + t;
    ar[t].frames = frames / threads + (t < frames % threads ? 1 : 0);
    pthread_create(&th[t], NULL, run_thread, &ar[t]);
  }
  int64_t iters = 0;
  for (int k = 0; k < 4; k++) counters[k] = 0;
  for (int t = 0; t < threads; t++) {
    pthread_join(th[t], NULL);
    for (int k = 0; k < 4; k++) counters[k] += ar[t].cnt[k];
    iters += ar[t].iters;
  }
  free(th);
  free(ar);
  return iters;
}

typedef struct {
  const kmo_code *c; const kmo_modem *m; const kmo_opts *o;
  double snr; long frame0, frames; int tid, threads; long chain_block;
  double *yy, *h, *hhat; int32_t *kstar, *ret, *nerr; uint8_t *converged, *uu, *uu_hat; double *metric;
} bulk_arg;

static void *bulk_thread(void *p) {
  bulk_arg *a = p;
  const int n_sym = a->c->n_tx / a->m->bits, k = a->c->k;
  int *u = malloc(sizeof(int) * k), *uh = malloc(sizeof(int) * k), *cch = malloc(sizeof(int) * a->c->n);
  double *y = malloc(sizeof(double) * 2 * n_sym);
  double *soft = malloc(sizeof(double) * a->c->m);
  const long blk = a->chain_block > 0 ? a->chain_block : 1;
  /* blocks of `blk` consecutive frames go through ONE codec state in frame order (a block = one reference worker task,
   * simulator.cc:86-97); blocks are dealt out to the threads, so the result does not depend on the thread count */
  for (long b0 = (long)a->tid * blk; b0 < a->frames; b0 += (long)a->threads * blk)
  for (long f = b0; f < a->frames && f < b0 + blk; f++) {
    if (f == b0) for (int r = 0; r < a->c->m; r++) soft[r] = 1.0;
    kmo_lcg g;
    kmo_lcg_seed(&g, (17 + 1000003L * (f + a->frame0)) % (LCG_M - 1) + 1);
    kmo_frame_out fo;
    kmo_frame(a->c, a->m, a->o, &g, a->snr, &fo, u, 0, y, 0, 0, cch, uh, soft);
    if (a->yy) memcpy(a->yy + (size_t)f * 2 * n_sym, y, sizeof(double) * 2 * n_sym);
    if (a->h) { a->h[2 * f] = fo.h[0]; a->h[2 * f + 1] = fo.h[1]; }
    if (a->hhat) { a->hhat[2 * f] = fo.hhat[0]; a->hhat[2 * f + 1] = fo.hhat[1]; }
    if (a->kstar) a->kstar[f] = fo.kstar;
    if (a->metric) for (int k = 0; k < 4; k++) a->metric[4 * f + k] = fo.metric[k];
    if (a->ret) a->ret[f] = fo.ret;
    if (a->nerr) a->nerr[f] = fo.nerr;
    if (a->converged) a->converged[f] = (uint8_t)(kmo_parity_check(a->c, cch) == 0);
    if (a->uu) for (int t = 0; t < k; t++) a->uu[(size_t)f * k + t] = (uint8_t)u[t];
    if (a->uu_hat) for (int t = 0; t < k; t++) a->uu_hat[(size_t)f * k + t] = (uint8_t)uh[t];
  }
  free(u); free(uh); free(cch); free(y); free(soft);
  return NULL;
}

void kmo_bulk(const kmo_code *c, const kmo_modem *m, const kmo_opts *o, double snr_db, long frame0, long frames, int threads,
              long chain_block, double *yy, double *h, double *hhat, int32_t *kstar, int32_t *ret, int32_t *nerr,
              uint8_t *converged, uint8_t *uu, uint8_t *uu_hat, double *metric) {
  if (threads < 1) threads = 1;
  pthread_t *th = malloc(sizeof(pthread_t) * threads);
  bulk_arg *ar = calloc(threads, sizeof(bulk_arg));
  for (int t = 0; t < threads; t++) {
    ar[t] = (bulk_arg){c, m, o, snr_db, frame0, frames, t, threads, chain_block, yy, h, hhat, kstar, ret, nerr, converged, uu, uu_hat, metric};
    pthread_create(&th[t], NULL, bulk_thread, &ar[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  free(th);
  free(ar);
}
