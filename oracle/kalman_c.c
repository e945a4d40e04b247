/* kalman_c.c -- plain-C restatement of oracle/kalman_np.py (local level).
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY; never linked into the product.
 *
 * Follows the same reference lines as kalman_np.py: model + priors
 * causalimpact/causalimpact_lib.py:398-500, mask extension :548-562; the
 * Kalman recursion itself is TFP's LinearGaussianStateSpaceModel (not in the
 * reference tree), restated with its update-then-predict convention.
 * Chains are distributed over OpenMP threads; arithmetic is float64.
 * Parity status: unpinned against the reference (see oracle/__init__.py);
 * pinned to kalman_np.py by tests/test_oracle_c.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LOG2PI 1.8378770664093453

/* prior[8] = m0, P0, obs_conc, obs_scale, obs_ub, lvl_conc, lvl_scale, lvl_ub */
static double one_chain(const double* y, const double* X, const double* Om, int T, int p,
                        const double* prior, const double* th, double* g, int with_prior,
                        double* r, double* A, double* P) {
  const double* w = th;
  const double u = th[p], l = th[p + 1];
  const double s_e = exp(u), s_h = exp(l);
  int t, j;
  for (t = 0; t < T; ++t) {
    double acc = y[t];
    const double* x = X + (size_t)t * p;
    for (j = 0; j < p; ++j) acc -= x[j] * w[j];
    r[t] = acc;
  }
  double a = prior[0], Pv = prior[1], ll = 0.0;
  for (t = 0; t < T; ++t) {
    A[t] = a; P[t] = Pv;
    if (r[t] == r[t]) {
      const double v = r[t] - a, F = Pv + s_e, K = Pv / F;
      ll += -0.5 * (LOG2PI + log(F) + v * v / F);
      a += K * v;
      Pv *= (1.0 - K);
    }
    Pv += s_h;
  }
  double val = ll;
  if (g) {
    double abar = 0.0, Pbar = 0.0, ge = 0.0, gh = 0.0;
    for (j = 0; j < p + 2; ++j) g[j] = 0.0;
    for (t = T - 1; t >= 0; --t) {
      gh += Pbar;
      if (!(r[t] == r[t])) continue;
      const double v = r[t] - A[t], F = P[t] + s_e, K = P[t] / F;
      const double dF = -0.5 * (1.0 / F - v * v / (F * F));
      const double rbar = K * abar - v / F;
      const double* x = X + (size_t)t * p;
      for (j = 0; j < p; ++j) g[j] -= rbar * x[j];
      ge += K * K * Pbar - abar * v * P[t] / (F * F) + dF;
      const double Pn = (1.0 - K) * (1.0 - K) * Pbar + abar * v * s_e / (F * F) + dF;
      abar = (1.0 - K) * abar + v / F;
      Pbar = Pn;
    }
    g[p] = ge * s_e;
    g[p + 1] = gh * s_h;
  }
  if (with_prior) {
    const double oc = prior[2], os = prior[3], oub = prior[4];
    const double lc = prior[5], ls = prior[6], lub = prior[7];
    double lp = -(oc + 1.0) * u - os / s_e + u - (lc + 1.0) * l - ls / s_h + l;
    if (g) {
      g[p] += -(oc + 1.0) + os / s_e + 1.0;
      g[p + 1] += -(lc + 1.0) + ls / s_h + 1.0;
    }
    if (p > 0) {
      double q = 0.0;
      int i;
      for (j = 0; j < p; ++j) {
        double ow = 0.0;
        for (i = 0; i < p; ++i) ow += Om[(size_t)j * p + i] * w[i];
        q += w[j] * ow;
        if (g) g[j] -= ow / s_e;
      }
      lp += -0.5 * p * u - 0.5 * q / s_e;
      if (g) g[p] += -0.5 * p + 0.5 * q / s_e;
    }
    val += lp;
    if (!(sqrt(s_e) <= oub && sqrt(s_h) <= lub)) val = -INFINITY;
  }
  return val;
}

/* theta [C, p+2] -> val [C], grad [C, p+2] (grad may be NULL).  Returns threads used. */
int ci_oracle_logpost_grad(const double* y, const double* X, const double* Om, int T, int p,
                           const double* prior, const double* theta, int C, double* val,
                           double* grad, int with_prior, int nthreads) {
  int used = 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  used = omp_get_max_threads();
#pragma omp parallel
#endif
  {
    double* buf = (double*)malloc(sizeof(double) * 3 * (size_t)T);
    int c;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (c = 0; c < C; ++c)
      val[c] = one_chain(y, X, Om, T, p, prior, theta + (size_t)c * (p + 2),
                         grad ? grad + (size_t)c * (p + 2) : NULL, with_prior, buf, buf + T,
                         buf + 2 * (size_t)T);
    free(buf);
  }
  return used;
}
