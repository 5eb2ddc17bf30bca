/*
 * nadavca_oracle.c -- CPU restatement of nadavca's native DP core.   *** TEST INFRASTRUCTURE ***
 *
 * This file is the parity ORACLE for the B200 kernels in nadavca_b200/csrc. It is a plain-C restatement
 * (written from the algorithm, not copied) of the reference's C++ under /root/reference/nadavca/dtw.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED. oracle/check_against_ref.py compares every entry point with the unmodified
 * reference compiled by oracle/Makefile (oracle/_ref) on randomised cases and requires bit-identical
 * doubles / ints; tests/golden/ holds vectors generated from oracle/_ref by oracle/make_golden.py.
 *
 * The arithmetic ORDER of the reference is kept on purpose (descending accumulation of the first cell,
 * recomputed m-term window, a + log(1 + exp(b - a)) with the larger operand first) so that results are
 * bit-identical to the reference built with the same compiler flags (-O2, no FMA contraction).
 *
 * All log-probabilities are doubles; "-inf" is log(0).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define NVO_NEG_INF (-INFINITY)

/* ---- log-space scalar: reference nadavca/dtw/probability.cpp:29-44 ---------------------------------- */

/* product of probabilities = sum of logs (probability.cpp:29-31) */
static inline double lp_mul(double a, double b) { return a + b; }

/* sum of probabilities (probability.cpp:33-40): order operands, short-circuit on log(0), log(1+exp()) */
static inline double lp_add(double a, double b) {
  if (a < b) { double t = a; a = b; b = t; }
  if (b == NVO_NEG_INF) return a;
  return a + log(1 + exp(b - a));
}

/* ---- k-mer model: reference nadavca/dtw/kmer_model.cpp ----------------------------------------------- */

typedef struct {
  int k, central, alphabet;
  int64_t size;          /* alphabet^k */
  double *mean, *ac, *mc; /* per k-mer: mean, additive const, multiplicative const (kmer_model.cpp:6-14) */
} nvo_model;

nvo_model *nvo_model_create(int k, int central, int alphabet, const double *mean, const double *sigma,
                            int64_t size) {
  nvo_model *m = (nvo_model *)calloc(1, sizeof(nvo_model));
  m->k = k; m->central = central; m->alphabet = alphabet; m->size = size;
  m->mean = (double *)malloc(sizeof(double) * size);
  m->ac = (double *)malloc(sizeof(double) * size);
  m->mc = (double *)malloc(sizeof(double) * size);
  for (int64_t i = 0; i < size; i++) {
    double s = sigma[i];
    m->mean[i] = mean[i];
    m->ac[i] = log(1 / sqrt(2 * M_PI * s * s)); /* kmer_model.cpp:11 */
    m->mc[i] = 1 / (2 * s * s);                 /* kmer_model.cpp:12 */
  }
  return m;
}

void nvo_model_destroy(nvo_model *m) {
  if (!m) return;
  free(m->mean); free(m->ac); free(m->mc); free(m);
}

/* extended sequence = context_before ++ reference ++ context_after, reads outside give base 0
 * (sequence.cpp:6-28); an optional single-base override models ModifiedSequence (sequence.cpp:30-38). */
typedef struct {
  const int *ref, *before, *after;
  int n, nb, na;
  int mod_pos, mod_val; /* mod_pos = INT32_MIN when unmodified */
} nvo_seq;

static inline int seq_at(const nvo_seq *s, int idx) {
  if (idx == s->mod_pos) return s->mod_val;
  int j = idx + s->nb;
  if (j < 0 || j >= s->nb + s->n + s->na) return 0;
  if (j < s->nb) return s->before[j];
  if (j < s->nb + s->n) return s->ref[j - s->nb];
  return s->after[j - s->nb - s->n];
}

/* kmer_model.cpp:22-30: base-alphabet number of seq[i-central .. i-central+k-1], first base most significant */
static inline int kmer_id(const nvo_model *m, const nvo_seq *s, int i) {
  int id = 0;
  for (int p = i - m->central; p < i - m->central + m->k; p++) id = id * m->alphabet + seq_at(s, p);
  return id;
}

/* One row's emission: Gaussian (kmer_model.cpp:44-52), wobble mixture (:54-62, "/ 2" subtracts 2.0 in
 * log space through the implicit Probability(double) ctor), or the constant transition (:64-94). */
enum { EM_GAUSS = 0, EM_MIX = 1, EM_CONST = 2 };
typedef struct {
  int kind;
  double mu1, ac1, mc1, mu2, ac2, mc2, cst;
} nvo_emis;

static inline double gauss_ld(double x, double mu, double ac, double mc) {
  double diff = x - mu;
  return ac - diff * diff * mc;
}

static inline double emis_ld(const nvo_emis *e, double x) {
  switch (e->kind) {
  case EM_GAUSS: return gauss_ld(x, e->mu1, e->ac1, e->mc1);
  case EM_MIX: return lp_add(gauss_ld(x, e->mu1, e->ac1, e->mc1), gauss_ld(x, e->mu2, e->ac2, e->mc2)) - 2.0;
  default: return e->cst;
  }
}

static nvo_emis emis_gauss(const nvo_model *m, const nvo_seq *s, int i) {
  nvo_emis e; memset(&e, 0, sizeof e);
  int id = kmer_id(m, s, i);
  e.kind = EM_GAUSS; e.mu1 = m->mean[id]; e.ac1 = m->ac[id]; e.mc1 = m->mc[id];
  return e;
}
static nvo_emis emis_mix(const nvo_model *m, const nvo_seq *s, int i1, int i2) {
  nvo_emis e; memset(&e, 0, sizeof e);
  int a = kmer_id(m, s, i1), b = kmer_id(m, s, i2);
  e.kind = EM_MIX;
  e.mu1 = m->mean[a]; e.ac1 = m->ac[a]; e.mc1 = m->mc[a];
  e.mu2 = m->mean[b]; e.ac2 = m->ac[b]; e.mc2 = m->mc[b];
  return e;
}
static nvo_emis emis_transition(const nvo_model *m, const nvo_seq *s, int i1, int i2) {
  nvo_emis e; memset(&e, 0, sizeof e);
  e.kind = EM_CONST;
  /* kmer_model.cpp:72-75 equal means -> log(0); otherwise the lambda returns p_in = log(0.01) first (:86) */
  e.cst = (m->mean[kmer_id(m, s, i1)] == m->mean[kmer_id(m, s, i2)]) ? NVO_NEG_INF : log(0.01);
  return e;
}

void nvo_expected_signal(const nvo_model *m, const int *ref, int n, const int *before, int nb,
                         const int *after, int na, double *out) {
  nvo_seq s = {ref, before, after, n, nb, na, INT32_MIN, 0};
  for (int i = 0; i < n; i++) out[i] = m->mean[kmer_id(m, &s, i)]; /* kmer_model.cpp:32-42 */
}

/* ---- bands: reference nadavca/dtw/dtw.cpp:7-35 -------------------------------------------------------- */

void nvo_band_bounds(const int *anchors, int n_anchors, int n_signal, int n_ref, int bandwidth, int *starts,
                     int *ends) {
  for (int i = 0; i <= n_ref; i++) { starts[i] = 0; ends[i] = n_signal; }
  for (int a = 0; a < n_anchors; a++) {
    int sig = anchors[2 * a], ref = anchors[2 * a + 1];
    int lo = sig - bandwidth, hi = sig + bandwidth;
    starts[ref] = lo > 0 ? lo : 0;
    ends[ref] = hi < n_signal ? hi : n_signal;
  }
  for (int i = 1; i <= n_ref; i++)
    if (starts[i - 1] > starts[i]) starts[i] = starts[i - 1];
  for (int i = n_ref - 1; i >= 0; i--)
    if (ends[i + 1] < ends[i]) ends[i] = ends[i + 1];
}

/* ---- one banded DP row: reference nadavca/dtw/node.cpp:5-37, node_next_row.h:6-61 ----------------------- */

typedef struct { int s, e; double *v; } nvo_row; /* inclusive [s,e]; v[i-s] */

static inline double row_at(const nvo_row *r, int i) { return (i < r->s || i > r->e) ? NVO_NEG_INF : r->v[i - r->s]; }

static void row_init(nvo_row *r, int s, int e, double fill, double *storage) {
  r->s = s; r->e = e; r->v = storage;
  for (int i = 0; i <= e - s; i++) storage[i] = fill;
}

static void next_row(nvo_row *res, int s, int e, const nvo_row *pred, const nvo_emis *em, const double *signal,
                     int n_signal, int m, int reverse, double *storage) {
  row_init(res, s, e, NVO_NEG_INF, storage);
  if (e < s) return;
  if (!reverse) {
    if (s >= m) { /* node_next_row.h:37-48 */
      double p = 0.0;
      for (int i = s; i >= pred->s; i--) {
        if (i < s) p = lp_mul(p, emis_ld(em, signal[i]));
        if (s - i >= m) res->v[0] = lp_add(res->v[0], lp_mul(p, row_at(pred, i)));
      }
    }
    for (int i = (s + 1 > m ? s + 1 : m); i <= e; i++) { /* node_next_row.h:49-58 */
      double p = 0.0;
      for (int j = i - 1; j >= i - m; j--) p = lp_mul(p, emis_ld(em, signal[j]));
      res->v[i - s] = lp_add(lp_mul(p, row_at(pred, i - m)), lp_mul(emis_ld(em, signal[i - 1]), row_at(res, i - 1)));
    }
  } else {
    if (e + m <= n_signal) { /* node_next_row.h:13-24 */
      double p = 0.0;
      for (int i = e; i <= pred->e; i++) {
        if (i > e) p = lp_mul(p, emis_ld(em, signal[i - 1]));
        if (i - e >= m) res->v[e - s] = lp_add(res->v[e - s], lp_mul(p, row_at(pred, i)));
      }
    }
    int top = e - 1 < n_signal - m ? e - 1 : n_signal - m;
    for (int i = top; i >= s; i--) { /* node_next_row.h:25-35 */
      double p = 0.0;
      for (int j = i; j < i + m; j++) p = lp_mul(p, emis_ld(em, signal[j]));
      res->v[i - s] = lp_add(lp_mul(p, row_at(pred, i + m)), lp_mul(emis_ld(em, signal[i]), row_at(res, i + 1)));
    }
  }
}

/* node.cpp:31-37 */
static double total_likelihood(const nvo_row *prefix, const nvo_row *suffix) {
  double r = NVO_NEG_INF;
  for (int i = prefix->s; i <= prefix->e; i++) r = lp_add(r, lp_mul(prefix->v[i - prefix->s], row_at(suffix, i)));
  return r;
}

/* Storage helper: rows of a banded matrix packed back to back. */
typedef struct { double *data; int64_t *off; nvo_row *rows; int nrows; } nvo_mat;

static int mat_alloc(nvo_mat *M, int nrows, const int *bs, const int *be) {
  M->nrows = nrows;
  M->off = (int64_t *)malloc(sizeof(int64_t) * (nrows + 1));
  M->rows = (nvo_row *)calloc(nrows, sizeof(nvo_row));
  int64_t t = 0;
  for (int r = 0; r < nrows; r++) { M->off[r] = t; int w = be[r] - bs[r] + 1; t += w > 0 ? w : 0; }
  M->off[nrows] = t;
  M->data = (double *)malloc(sizeof(double) * (t > 0 ? t : 1));
  return M->data && M->off && M->rows;
}
static void mat_free(nvo_mat *M) { free(M->data); free(M->off); free(M->rows); }

/* ---- RefineAlignment: reference nadavca/dtw/dtw.cpp:133-228 -------------------------------------------- */

/* Returns 1 and fills events[n][2] when a path exists, 0 for "no valid path" (reference returns an empty list).
 * Optional debug outputs (may be NULL): dbg_rows[0] = row count, dbg_bs/dbg_be per row, dbg_prefix/dbg_suffix
 * packed rows (caller sizes them with nvo_refine_cells). */
int64_t nvo_refine_cells(const int *anchors, int n_anchors, int n_signal, int n_ref, int bandwidth,
                         int transitions) {
  int *bs = (int *)malloc(sizeof(int) * (n_ref + 1)), *be = (int *)malloc(sizeof(int) * (n_ref + 1));
  nvo_band_bounds(anchors, n_anchors, n_signal, n_ref, bandwidth, bs, be);
  int64_t t = 0;
  if (transitions) {
    for (int i = 0; i < n_ref; i++) t += (be[i] - bs[i] + 1) + (be[i + 1] - bs[i + 1] + 1);
  } else {
    for (int i = 0; i <= n_ref; i++) t += be[i] - bs[i] + 1;
  }
  free(bs); free(be);
  return t;
}

int nvo_refine_alignment(const nvo_model *model, const double *signal, int n_signal, const int *ref, int n_ref,
                         const int *before, int nb, const int *after, int na, const int *anchors, int n_anchors,
                         int bandwidth, int m_len, int transitions, int *events, double *dbg_prefix,
                         double *dbg_suffix, int *dbg_bs, int *dbg_be) {
  if (n_ref <= 0) return 0;
  nvo_seq seq = {ref, before, after, n_ref, nb, na, INT32_MIN, 0};
  int *b0s = (int *)malloc(sizeof(int) * (n_ref + 1)), *b0e = (int *)malloc(sizeof(int) * (n_ref + 1));
  nvo_band_bounds(anchors, n_anchors, n_signal, n_ref, bandwidth, b0s, b0e);

  int R = transitions ? 2 * n_ref : n_ref + 1; /* dtw.cpp:144-159 */
  int *bs = (int *)malloc(sizeof(int) * R), *be = (int *)malloc(sizeof(int) * R);
  nvo_emis *em = (nvo_emis *)calloc(R, sizeof(nvo_emis));
  int *mel = (int *)calloc(R, sizeof(int));
  if (transitions) {
    for (int i = 0; i < n_ref; i++) {
      bs[2 * i] = b0s[i]; be[2 * i] = b0e[i];
      bs[2 * i + 1] = b0s[i + 1]; be[2 * i + 1] = b0e[i + 1];
      em[2 * i] = emis_gauss(model, &seq, i); mel[2 * i] = m_len; /* dtw.cpp:165-174 */
      if (i + 1 < n_ref) { em[2 * i + 1] = emis_transition(model, &seq, i, i + 1); mel[2 * i + 1] = 0; }
    }
  } else {
    for (int i = 0; i <= n_ref; i++) { bs[i] = b0s[i]; be[i] = b0e[i]; }
    for (int i = 0; i < n_ref; i++) { em[i] = emis_gauss(model, &seq, i); mel[i] = m_len; }
  }

  nvo_mat P, S;
  mat_alloc(&P, R, bs, be); mat_alloc(&S, R, bs, be);
  row_init(&P.rows[0], bs[0], be[0], 0.0, P.data + P.off[0]); /* dtw.cpp:182 */
  for (int r = 0; r + 1 < R; r++)
    next_row(&P.rows[r + 1], bs[r + 1], be[r + 1], &P.rows[r], &em[r], signal, n_signal, mel[r], 0,
             P.data + P.off[r + 1]);
  row_init(&S.rows[R - 1], bs[R - 1], be[R - 1], 0.0, S.data + S.off[R - 1]); /* dtw.cpp:190 */
  for (int r = R - 1; r > 0; r--)
    next_row(&S.rows[r - 1], bs[r - 1], be[r - 1], &S.rows[r], &em[r - 1], signal, n_signal, mel[r - 1], 1,
             S.data + S.off[r - 1]);

  if (dbg_prefix) memcpy(dbg_prefix, P.data, sizeof(double) * P.off[R]);
  if (dbg_suffix) memcpy(dbg_suffix, S.data, sizeof(double) * S.off[R]);
  if (dbg_bs) memcpy(dbg_bs, bs, sizeof(int) * R);
  if (dbg_be) memcpy(dbg_be, be, sizeof(int) * R);

  /* posterior rows (node.cpp:23-29) then the max-product path search (node.cpp:60-91) */
  int *prev = (int *)malloc(sizeof(int) * (P.off[R] > 0 ? P.off[R] : 1));
  double *dp_prev = NULL, *dp_cur = NULL;
  int wmax = 1;
  for (int r = 0; r < R; r++) if (be[r] - bs[r] + 1 > wmax) wmax = be[r] - bs[r] + 1;
  dp_prev = (double *)malloc(sizeof(double) * wmax); dp_cur = (double *)malloc(sizeof(double) * wmax);
  for (int i = bs[0]; i <= be[0]; i++) {
    dp_prev[i - bs[0]] = lp_mul(P.rows[0].v[i - bs[0]], S.rows[0].v[i - bs[0]]);
    prev[P.off[0] + i - bs[0]] = -1;
  }
  for (int r = 1; r < R; r++) {
    int s = bs[r], e = be[r], ps = bs[r - 1], pe = be[r - 1], m = mel[r - 1];
    int best_i = -1; double best = NVO_NEG_INF;
    for (int i = ps; i <= pe && i < s - m; i++)
      if (dp_prev[i - ps] > best) { best = dp_prev[i - ps]; best_i = i; }
    for (int i = s; i <= e; i++) {
      int from = i - m;
      if (from >= ps && from <= pe && dp_prev[from - ps] > best) { best = dp_prev[from - ps]; best_i = from; }
      double score = lp_mul(P.rows[r].v[i - s], S.rows[r].v[i - s]);
      dp_cur[i - s] = lp_mul(best, score);
      prev[P.off[r] + i - s] = best_i;
    }
    double *t = dp_prev; dp_prev = dp_cur; dp_cur = t;
  }
  /* node.cpp:48-58 final argmax, strict '>' from log(0) so the lowest index wins; -1 = no path */
  int best_i = -1; double best = NVO_NEG_INF;
  for (int i = bs[R - 1]; i <= be[R - 1]; i++)
    if (dp_prev[i - bs[R - 1]] > best) { best = dp_prev[i - bs[R - 1]]; best_i = i; }
  int ok = best_i != -1;
  if (ok) { /* dtw.cpp:215-227 */
    for (int r = R - 1; r >= 0; r--) {
      if (transitions) events[(r / 2) * 2 + (r % 2)] = best_i;
      else {
        if (r > 0) events[(r - 1) * 2 + 1] = best_i;
        if (r + 1 < R) events[r * 2] = best_i;
      }
      best_i = prev[P.off[r] + best_i - bs[r]];
    }
  }
  free(prev); free(dp_prev); free(dp_cur); mat_free(&P); mat_free(&S);
  free(bs); free(be); free(em); free(mel); free(b0s); free(b0e);
  return ok;
}

/* ---- EstimateLogLikelihoods: reference nadavca/dtw/dtw.cpp:37-131 --------------------------------------- */

void nvo_estimate_log_likelihoods(const nvo_model *model, const double *signal, int n_signal, const int *ref,
                                  int n_ref, const int *before, int nb, const int *after, int na,
                                  const int *anchors, int n_anchors, int bandwidth, int m_len, int wobbling,
                                  double *out /* n_ref x alphabet */, double *dbg_prefix, double *dbg_suffix) {
  if (n_ref <= 0) return;
  nvo_seq seq = {ref, before, after, n_ref, nb, na, INT32_MIN, 0};
  int n = n_ref, A = model->alphabet;
  int *bs = (int *)malloc(sizeof(int) * (n + 1)), *be = (int *)malloc(sizeof(int) * (n + 1));
  nvo_band_bounds(anchors, n_anchors, n_signal, n, bandwidth, bs, be);
  int wmax = 1;
  for (int r = 0; r <= n; r++) if (be[r] - bs[r] + 1 > wmax) wmax = be[r] - bs[r] + 1;
  double *tmp_a = (double *)malloc(sizeof(double) * wmax), *tmp_b = (double *)malloc(sizeof(double) * wmax);

  nvo_mat P, S;
  mat_alloc(&P, n + 1, bs, be); mat_alloc(&S, n + 1, bs, be);
  row_init(&P.rows[0], bs[0], be[0], 0.0, P.data + P.off[0]); /* dtw.cpp:50 */
  for (int i = 0; i < n; i++) {                                /* dtw.cpp:51-64 */
    nvo_row pred = P.rows[i], wob;
    if (i > 0 && wobbling) {
      nvo_emis mix = emis_mix(model, &seq, i - 1, i);
      next_row(&wob, bs[i], be[i], &pred, &mix, signal, n_signal, 0, 0, tmp_a);
      pred = wob;
    }
    nvo_emis g = emis_gauss(model, &seq, i);
    next_row(&P.rows[i + 1], bs[i + 1], be[i + 1], &pred, &g, signal, n_signal, m_len, 0, P.data + P.off[i + 1]);
  }
  row_init(&S.rows[n], bs[n], be[n], 0.0, S.data + S.off[n]); /* dtw.cpp:66-67 */
  for (int i = n; i > 0; i--) {                                /* dtw.cpp:68-81 */
    nvo_row pred = S.rows[i], wob;
    if (i < n && wobbling) {
      nvo_emis mix = emis_mix(model, &seq, i, i - 1);
      next_row(&wob, bs[i], be[i], &pred, &mix, signal, n_signal, 0, 1, tmp_a);
      pred = wob;
    }
    nvo_emis g = emis_gauss(model, &seq, i - 1);
    next_row(&S.rows[i - 1], bs[i - 1], be[i - 1], &pred, &g, signal, n_signal, m_len, 1, S.data + S.off[i - 1]);
  }
  if (dbg_prefix) memcpy(dbg_prefix, P.data, sizeof(double) * P.off[n + 1]);
  if (dbg_suffix) memcpy(dbg_suffix, S.data, sizeof(double) * S.off[n + 1]);

  double no_snp = total_likelihood(&P.rows[n], &S.rows[n]); /* dtw.cpp:83-85 */
  int back = model->k - model->central - 1, fwd = model->central; /* dtw.cpp:88-89 */

  for (int i = 0; i < n; i++) { /* dtw.cpp:93-129 */
    int first = i - back > 0 ? i - back : 0;
    int last = i + fwd < n - 1 ? i + fwd : n - 1;
    for (int base = 0; base < A; base++) {
      if (base == seq_at(&seq, i)) { out[i * A + base] = no_snp; continue; }
      nvo_seq mod = seq; mod.mod_pos = i; mod.mod_val = base;
      nvo_row cur = P.rows[first], nxt;
      double *buf = tmp_a, *other = tmp_b;
      for (int j = first; j <= last; j++) {
        if (j > 0 && wobbling) {
          nvo_emis mix = emis_mix(model, &mod, j - 1, j);
          next_row(&nxt, bs[j], be[j], &cur, &mix, signal, n_signal, 0, 0, buf);
          cur = nxt; { double *t = buf; buf = other; other = t; }
        }
        nvo_emis g = emis_gauss(model, &mod, j);
        next_row(&nxt, bs[j + 1], be[j + 1], &cur, &g, signal, n_signal, m_len, 0, buf);
        cur = nxt; { double *t = buf; buf = other; other = t; }
      }
      if (last + 1 < n && wobbling) { /* dtw.cpp:116-123: band row `last`, not last+1 */
        nvo_emis mix = emis_mix(model, &mod, last, last + 1);
        next_row(&nxt, bs[last], be[last], &cur, &mix, signal, n_signal, 0, 0, buf);
        cur = nxt;
      }
      out[i * A + base] = total_likelihood(&cur, &S.rows[last + 1]);
    }
  }
  mat_free(&P); mat_free(&S); free(bs); free(be); free(tmp_a); free(tmp_b);
}

/* Cell-count formulas of SURVEY.md section 8(d) (units of work for the throughput metric). */
void nvo_count_cells(const int *anchors, int n_anchors, int n_signal, int n_ref, int bandwidth, int k, int central,
                     int64_t *refine_trans, int64_t *refine_plain, int64_t *ell_fb, int64_t *ell_snp) {
  int n = n_ref;
  int *bs = (int *)malloc(sizeof(int) * (n + 1)), *be = (int *)malloc(sizeof(int) * (n + 1));
  nvo_band_bounds(anchors, n_anchors, n_signal, n, bandwidth, bs, be);
#define W(j) ((int64_t)(be[j] - bs[j] + 1))
  int64_t t = 0;
  for (int rho = 0; rho < 2 * n; rho++) {
    int64_t w = (rho % 2 == 0) ? W(rho / 2) : W(rho / 2 + 1);
    if (rho >= 1) t += w;
    if (rho <= 2 * n - 2) t += w;
  }
  *refine_trans = t;
  t = 0;
  for (int j = 1; j <= n; j++) t += W(j);
  for (int j = 0; j < n; j++) t += W(j);
  *refine_plain = t;
  t = 0;
  for (int i = 0; i < n; i++) t += W(i + 1);
  for (int i = 1; i < n; i++) t += W(i);
  for (int i = 1; i <= n; i++) t += W(i - 1);
  for (int i = 1; i < n; i++) t += W(i);
  *ell_fb = t;
  t = 0;
  int back = k - central - 1, fwd = central;
  for (int i = 0; i < n; i++) {
    int first = i - back > 0 ? i - back : 0, last = i + fwd < n - 1 ? i + fwd : n - 1;
    int64_t c = 0;
    for (int j = first; j <= last; j++) c += W(j + 1) + (j > 0 ? W(j) : 0);
    if (last + 1 < n) c += W(last);
    t += 3 * c;
  }
  *ell_snp = t;
#undef W
  free(bs); free(be);
}
