/*
 * TEST INFRASTRUCTURE — plain-C restatement of the two per-chunk primitives on the
 * dask_ml.cluster.KMeans hot path.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs
 * may load this; the product never does.
 *
 *  okm_centers_dense_*  : the M-step scatter-add, dask_ml/cluster/k_means.py:572-582
 *                         (float64 accumulators regardless of the dtype of X, :576).
 *  okm_argmin_min_*     : the E-step of sklearn.metrics.pairwise_distances_argmin_min as the
 *                         reference calls it (dask_ml/metrics/pairwise.py:35-38, squared=True):
 *                         d2 = ||x||^2 - 2 x.y + ||y||^2 evaluated in float64, max(d2, 0),
 *                         strict '<' scan so that ties go to the lowest index.
 */
#include <stdint.h>
#include <stdlib.h>
#include <math.h>

#define CENTERS_DENSE(NAME, T)                                                                   \
  void NAME(const T* X, const int32_t* labels, int64_t n, int d, int k, double* out) {           \
    (void)k;                                                                                     \
    for (int64_t i = 0; i < n; ++i) {                                                            \
      double* row = out + (int64_t)labels[i] * d;                                                \
      const T* x = X + i * d;                                                                    \
      for (int j = 0; j < d; ++j) row[j] += (double)x[j];                                        \
    }                                                                                            \
  }
CENTERS_DENSE(okm_centers_dense_f32, float)
CENTERS_DENSE(okm_centers_dense_f64, double)

#define ARGMIN_MIN(NAME, T)                                                                      \
  void NAME(const T* X, int64_t n, int d, const T* Y, int k, int64_t* lab, double* mn) {         \
    double* yn = (double*)malloc(sizeof(double) * (size_t)k);                                    \
    for (int j = 0; j < k; ++j) {                                                                \
      double s = 0.0;                                                                            \
      for (int t = 0; t < d; ++t) s += (double)Y[(int64_t)j * d + t] * (double)Y[(int64_t)j * d + t]; \
      yn[j] = s;                                                                                 \
    }                                                                                            \
    for (int64_t i = 0; i < n; ++i) {                                                            \
      const T* x = X + i * d;                                                                    \
      double xn = 0.0;                                                                           \
      for (int t = 0; t < d; ++t) xn += (double)x[t] * (double)x[t];                             \
      double best = INFINITY; int64_t bj = 0;                                                    \
      for (int j = 0; j < k; ++j) {                                                              \
        const T* y = Y + (int64_t)j * d;                                                         \
        double dot = 0.0;                                                                        \
        for (int t = 0; t < d; ++t) dot += (double)x[t] * (double)y[t];                          \
        double d2 = xn - 2.0 * dot + yn[j];                                                      \
        if (d2 < 0.0) d2 = 0.0;                                                                  \
        if (d2 < best) { best = d2; bj = j; }                                                    \
      }                                                                                          \
      lab[i] = bj; mn[i] = best;                                                                 \
    }                                                                                            \
    free(yn);                                                                                    \
  }
ARGMIN_MIN(okm_argmin_min_f32, float)
ARGMIN_MIN(okm_argmin_min_f64, double)
