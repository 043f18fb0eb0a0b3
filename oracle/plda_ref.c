/* CPU fp64 ORACLE (C restatement) of the reference's PLDA path -- TEST / BASELINE
 * INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.
 *
 * PARITY UNPINNED (see oracle/kaldi_plda.py header): Kaldi's ivector/plda.cc is not in the
 * reference tree and cannot be built here; this file restates its published algorithm in
 * the order src/pldamodule.cpp drives it, single-threaded like the reference
 * (Kaldi + ATLAS, no threads anywhere in src/), so it can be TIMED as "the reference CPU
 * path" (BASELINE.md "CPU-A reference-faithful").  tests/test_oracle_c.py pins it against
 * oracle/kaldi_plda.py.  The all-pairs scoring loop can additionally use OpenMP threads
 * (the caller's Python double loop, README.md:108-113, is embarrassingly parallel).
 *
 * Build: make -C oracle   ->  oracle/_build/libplda_ref.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define M_LOG_2PI 1.8378770664093454835606594728112

typedef struct {
  int d, k;
  double *offset_scatter; /* d*d (full symmetric)            PldaStats::offset_scatter_ */
  double *sum;            /* d                                PldaStats::sum_            */
  double *means;          /* k*d, sorted by num_examples      ClassInfo::mean            */
  double *weights;        /* k                                ClassInfo::weight          */
  int *nex;               /* k                                ClassInfo::num_examples    */
  double class_weight, example_weight;
  double *within, *between;                 /* PldaEstimator::within_var_, between_var_ */
  double *within_stats, *between_stats;
  double within_count, between_count;
} ref_state;

/* ---- small dense helpers (row-major d x d) ---- */
static void chol_lower(const double *a, double *l, int d) { /* TpMatrix::Cholesky */
  memset(l, 0, sizeof(double) * d * d);
  for (int j = 0; j < d; ++j) {
    double s = a[j * d + j];
    for (int k = 0; k < j; ++k) s -= l[j * d + k] * l[j * d + k];
    double ljj = sqrt(s);
    l[j * d + j] = ljj;
    for (int i = j + 1; i < d; ++i) {
      double t = a[i * d + j];
      for (int k = 0; k < j; ++k) t -= l[i * d + k] * l[j * d + k];
      l[i * d + j] = t / ljj;
    }
  }
}
static void tri_inv_lower(const double *l, double *x, int d) { /* TpMatrix::Invert */
  memset(x, 0, sizeof(double) * d * d);
  for (int j = 0; j < d; ++j) {
    x[j * d + j] = 1.0 / l[j * d + j];
    for (int i = j + 1; i < d; ++i) {
      double s = 0.0;
      for (int k = j; k < i; ++k) s += l[i * d + k] * x[k * d + j];
      x[i * d + j] = -s / l[i * d + i];
    }
  }
}
static void spd_invert(const double *a, double *inv, double *w1, double *w2, int d) { /* SpMatrix::Invert */
  chol_lower(a, w1, d);
  tri_inv_lower(w1, w2, d);
  for (int i = 0; i < d; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = 0.0;
      for (int k = i; k < d; ++k) s += w2[k * d + i] * w2[k * d + j];
      inv[i * d + j] = s;
      inv[j * d + i] = s;
    }
}
static void symv(const double *a, const double *x, double alpha, double *y, int d) { /* AddSpVec(alpha, A, x, 0) */
  for (int i = 0; i < d; ++i) {
    double s = 0.0;
    const double *r = a + (size_t)i * d;
    for (int j = 0; j < d; ++j) s += r[j] * x[j];
    y[i] = alpha * s;
  }
}
static void add_sp(double *a, double alpha, const double *b, int d) { /* AddSp: lower triangle (packed) */
  for (int i = 0; i < d; ++i)
    for (int j = 0; j <= i; ++j) a[i * d + j] += alpha * b[i * d + j];
}
static void add_vec2(double *a, double alpha, const double *v, int d) { /* AddVec2: lower triangle */
  for (int i = 0; i < d; ++i) {
    double vi = alpha * v[i];
    for (int j = 0; j <= i; ++j) a[i * d + j] += vi * v[j];
  }
}
static void mirror_lower(double *a, int d) {
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < i; ++j) a[j * d + i] = a[i * d + j];
}
/* cyclic Jacobi symmetric eigensolver (stands in for SpMatrix::Eig); v columns = eigenvectors */
static void jacobi_eig(double *a, double *v, double *w, int d) {
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) v[i * d + j] = (i == j);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < d; ++i) {
      diag += a[i * d + i] * a[i * d + i];
      for (int j = 0; j < i; ++j) off += a[i * d + j] * a[i * d + j];
    }
    if (off <= 1e-30 * diag) break;
    for (int p = 0; p < d - 1; ++p)
      for (int q = p + 1; q < d; ++q) {
        double apq = a[p * d + q];
        if (fabs(apq) < 1e-300) continue;
        double theta = (a[q * d + q] - a[p * d + p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < d; ++k) {
          double akp = a[k * d + p], akq = a[k * d + q];
          a[k * d + p] = c * akp - s * akq;
          a[k * d + q] = s * akp + c * akq;
        }
        for (int k = 0; k < d; ++k) {
          double apk = a[p * d + k], aqk = a[q * d + k];
          a[p * d + k] = c * apk - s * aqk;
          a[q * d + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < d; ++k) {
          double vkp = v[k * d + p], vkq = v[k * d + q];
          v[k * d + p] = c * vkp - s * vkq;
          v[k * d + q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < d; ++i) w[i] = a[i * d + i];
}

/* ---- PldaStats: src/pldamodule.cpp:70-100 ---- */
ref_state *plda_ref_stats(const double *x, int64_t n, int d, const int64_t *labels, int k) {
  ref_state *s = (ref_state *)calloc(1, sizeof(ref_state));
  s->d = d; s->k = k;
  s->offset_scatter = (double *)calloc((size_t)d * d, sizeof(double));
  s->sum = (double *)calloc(d, sizeof(double));
  double *means = (double *)calloc((size_t)k * d, sizeof(double));
  int *cnt = (int *)calloc(k, sizeof(int));
  /* bucket rows per speaker (:88-92) */
  int64_t *start = (int64_t *)calloc(k + 1, sizeof(int64_t));
  for (int64_t i = 0; i < n; ++i) start[labels[i] + 1]++;
  for (int c = 0; c < k; ++c) start[c + 1] += start[c];
  int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * k);
  int64_t *rows = (int64_t *)malloc(sizeof(int64_t) * n);
  memcpy(fill, start, sizeof(int64_t) * k);
  for (int64_t i = 0; i < n; ++i) rows[fill[labels[i]]++] = i;
  double *tmp = NULL; size_t tmp_cap = 0;
  for (int c = 0; c < k; ++c) { /* :94-98 CopyRows + AddSamples(1/n_s, tmp) */
    int ns = (int)(start[c + 1] - start[c]);
    if ((size_t)ns * d > tmp_cap) { tmp_cap = (size_t)ns * d; tmp = (double *)realloc(tmp, tmp_cap * sizeof(double)); }
    for (int r = 0; r < ns; ++r) memcpy(tmp + (size_t)r * d, x + (size_t)rows[start[c] + r] * d, sizeof(double) * d);
    double w = 1.0 / ns;
    double *m = means + (size_t)c * d;
    for (int r = 0; r < ns; ++r)
      for (int j = 0; j < d; ++j) m[j] += tmp[(size_t)r * d + j];
    for (int j = 0; j < d; ++j) m[j] *= 1.0 / ns;
    for (int r = 0; r < ns; ++r) { /* AddMat2(weight, group, kTrans): lower triangle */
      const double *g = tmp + (size_t)r * d;
      for (int i = 0; i < d; ++i) {
        double gi = w * g[i];
        double *o = s->offset_scatter + (size_t)i * d;
        for (int j = 0; j <= i; ++j) o[j] += gi * g[j];
      }
    }
    add_vec2(s->offset_scatter, -ns * w, m, d);
    cnt[c] = ns;
    s->class_weight += w;
    s->example_weight += w * ns;
    for (int j = 0; j < d; ++j) s->sum[j] += w * m[j];
  }
  mirror_lower(s->offset_scatter, d);
  /* Sort(): ascending num_examples (:100) -- stable counting order */
  s->means = (double *)malloc(sizeof(double) * (size_t)k * d);
  s->weights = (double *)malloc(sizeof(double) * k);
  s->nex = (int *)malloc(sizeof(int) * k);
  int *order = (int *)malloc(sizeof(int) * k);
  for (int c = 0; c < k; ++c) order[c] = c;
  /* insertion-stable merge sort substitute: simple stable bucket by count */
  int maxn = 0;
  for (int c = 0; c < k; ++c) if (cnt[c] > maxn) maxn = cnt[c];
  int *bstart = (int *)calloc(maxn + 2, sizeof(int));
  for (int c = 0; c < k; ++c) bstart[cnt[c] + 1]++;
  for (int b = 0; b <= maxn; ++b) bstart[b + 1] += bstart[b];
  for (int c = 0; c < k; ++c) order[bstart[cnt[c]]++] = c;
  for (int i = 0; i < k; ++i) {
    int c = order[i];
    memcpy(s->means + (size_t)i * d, means + (size_t)c * d, sizeof(double) * d);
    s->weights[i] = 1.0 / cnt[c];
    s->nex[i] = cnt[c];
  }
  s->within = (double *)calloc((size_t)d * d, sizeof(double));
  s->between = (double *)calloc((size_t)d * d, sizeof(double));
  s->within_stats = (double *)calloc((size_t)d * d, sizeof(double));
  s->between_stats = (double *)calloc((size_t)d * d, sizeof(double));
  for (int i = 0; i < d; ++i) s->within[i * d + i] = s->between[i * d + i] = 1.0; /* InitParameters */
  free(means); free(cnt); free(start); free(fill); free(rows); free(tmp); free(order); free(bstart);
  return s;
}

/* ---- PldaEstimator::EstimateOneIter ---- */
void plda_ref_em_iter(ref_state *s) {
  const int d = s->d;
  const size_t dd = (size_t)d * d;
  double *binv = (double *)malloc(sizeof(double) * dd), *winv = (double *)malloc(sizeof(double) * dd);
  double *mixed = (double *)malloc(sizeof(double) * dd), *tmpm = (double *)malloc(sizeof(double) * dd);
  double *w1 = (double *)malloc(sizeof(double) * dd), *w2 = (double *)malloc(sizeof(double) * dd);
  double *m = (double *)malloc(sizeof(double) * d), *temp = (double *)malloc(sizeof(double) * d);
  double *w = (double *)malloc(sizeof(double) * d), *mw = (double *)malloc(sizeof(double) * d);
  memset(s->within_stats, 0, sizeof(double) * dd);
  memset(s->between_stats, 0, sizeof(double) * dd);
  s->within_count = s->between_count = 0.0;
  /* GetStatsFromIntraClass */
  add_sp(s->within_stats, 1.0, s->offset_scatter, d);
  s->within_count += s->example_weight - s->class_weight;
  /* GetStatsFromClassMeans */
  spd_invert(s->between, binv, w1, w2, d);
  spd_invert(s->within, winv, w1, w2, d);
  int n = -1;
  for (int c = 0; c < s->k; ++c) {
    const double weight = s->weights[c];
    if (s->nex[c] != n) {
      n = s->nex[c];
      for (size_t i = 0; i < dd; ++i) tmpm[i] = binv[i] + n * winv[i];
      spd_invert(tmpm, mixed, w1, w2, d);
    }
    for (int j = 0; j < d; ++j) m[j] = s->means[(size_t)c * d + j] - s->sum[j] / s->class_weight;
    symv(winv, m, (double)n, temp, d);
    symv(mixed, temp, 1.0, w, d);
    for (int j = 0; j < d; ++j) mw[j] = m[j] - w[j];
    add_sp(s->between_stats, weight, mixed, d);
    add_vec2(s->between_stats, weight, w, d);
    s->between_count += weight;
    add_sp(s->within_stats, weight * n, mixed, d);
    add_vec2(s->within_stats, weight * n, mw, d);
    s->within_count += weight;
  }
  /* EstimateFromStats */
  for (int i = 0; i < d; ++i)
    for (int j = 0; j <= i; ++j) {
      s->within[i * d + j] = s->within[j * d + i] = s->within_stats[i * d + j] / s->within_count;
      s->between[i * d + j] = s->between[j * d + i] = s->between_stats[i * d + j] / s->between_count;
    }
  free(binv); free(winv); free(mixed); free(tmpm); free(w1); free(w2); free(m); free(temp); free(w); free(mw);
}

/* ---- PldaEstimator::GetOutput ---- */
void plda_ref_get_output(ref_state *s, double *mean, double *transform, double *psi) {
  const int d = s->d;
  const size_t dd = (size_t)d * d;
  double *c = (double *)malloc(sizeof(double) * dd), *t1 = (double *)malloc(sizeof(double) * dd);
  double *bp = (double *)malloc(sizeof(double) * dd), *tmp = (double *)malloc(sizeof(double) * dd);
  double *u = (double *)malloc(sizeof(double) * dd), *ev = (double *)malloc(sizeof(double) * d);
  for (int j = 0; j < d; ++j) mean[j] = s->sum[j] / s->class_weight;
  chol_lower(s->within, c, d);
  tri_inv_lower(c, t1, d);
  for (int i = 0; i < d; ++i) /* tmp = T1 * B */
    for (int j = 0; j < d; ++j) {
      double acc = 0.0;
      for (int k = 0; k <= i; ++k) acc += t1[i * d + k] * s->between[k * d + j];
      tmp[i * d + j] = acc;
    }
  for (int i = 0; i < d; ++i) /* bp = tmp * T1^T */
    for (int j = 0; j <= i; ++j) {
      double acc = 0.0;
      for (int k = 0; k <= j; ++k) acc += tmp[i * d + k] * t1[j * d + k];
      bp[i * d + j] = bp[j * d + i] = acc;
    }
  jacobi_eig(bp, u, ev, d);
  int *ord = (int *)malloc(sizeof(int) * d);
  for (int i = 0; i < d; ++i) { ord[i] = i; if (ev[i] < 0.0) ev[i] = 0.0; }
  for (int i = 1; i < d; ++i) { /* SortSvd: descending */
    int o = ord[i], j = i - 1;
    while (j >= 0 && ev[ord[j]] < ev[o]) { ord[j + 1] = ord[j]; --j; }
    ord[j + 1] = o;
  }
  for (int i = 0; i < d; ++i) {
    psi[i] = ev[ord[i]];
    for (int j = 0; j < d; ++j) { /* transform = U^T T1 */
      double acc = 0.0;
      for (int k = j; k < d; ++k) acc += u[k * d + ord[i]] * t1[k * d + j];
      transform[i * d + j] = acc;
    }
  }
  free(c); free(t1); free(bp); free(tmp); free(u); free(ev); free(ord);
}

void plda_ref_get_covariances(ref_state *s, double *within, double *between) {
  memcpy(within, s->within, sizeof(double) * s->d * s->d);
  memcpy(between, s->between, sizeof(double) * s->d * s->d);
}

void plda_ref_free(ref_state *s) {
  if (!s) return;
  free(s->offset_scatter); free(s->sum); free(s->means); free(s->weights); free(s->nex);
  free(s->within); free(s->between); free(s->within_stats); free(s->between_stats); free(s);
}

/* ---- Plda::TransformIvector (default config), src/pldamodule.cpp:171,224 ---- */
void plda_ref_transform(const double *mean, const double *transform, const double *psi, int d, const double *x,
                        int num_examples, double *out) {
  double dot = 0.0;
  for (int i = 0; i < d; ++i) {
    double acc = 0.0, off = 0.0;
    const double *r = transform + (size_t)i * d;
    for (int j = 0; j < d; ++j) { acc += r[j] * x[j]; off += r[j] * mean[j]; }
    out[i] = acc - off;
    dot += out[i] * out[i] / (psi[i] + 1.0 / num_examples);
  }
  double f = sqrt(d / dot);
  for (int i = 0; i < d; ++i) out[i] *= f;
}

/* ---- Plda::LogLikelihoodRatio, src/pldamodule.cpp:235,266 ---- */
double plda_ref_llr(const double *psi, int d, const double *train, int n, const double *test) {
  double logdet = 0.0, quad = 0.0, logdet2 = 0.0, quad2 = 0.0;
  for (int i = 0; i < d; ++i) {
    double mean = n * psi[i] / (n * psi[i] + 1.0) * train[i];
    double var = 1.0 + psi[i] / (n * psi[i] + 1.0);
    double diff = test[i] - mean;
    logdet += log(var);
    quad += diff * diff / var;
    double var2 = psi[i] + 1.0;
    logdet2 += log(var2);
    quad2 += test[i] * test[i] / var2;
  }
  double given = -0.5 * (logdet + M_LOG_2PI * d + quad);
  double without = -0.5 * (logdet2 + M_LOG_2PI * d + quad2);
  return given - without;
}

/* The caller's all-pairs double loop (README.md:108-113, scoring/scorePLDA.py:302-318): one
 * LogLikelihoodRatio per trial, result through float32 like MPlda_score (:276). threads<=0: all cores. */
int plda_ref_score_grid(const double *psi, int d, const double *enrol, const int32_t *counts, int64_t ne,
                        const double *test, int64_t nt, float *out, int threads) {
  int used = 1;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
  used = threads;
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
  for (int64_t e = 0; e < ne; ++e)
    for (int64_t t = 0; t < nt; ++t)
      out[e * nt + t] = (float)plda_ref_llr(psi, d, enrol + e * d, counts[e], test + t * d);
  return used;
}

int plda_ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
