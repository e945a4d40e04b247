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

/* ------------------------------------------------------------------------------------------
 * Local linear trend (d = 2): restates oracle/kalman_np.py gen_filter / gen_filter_grad for
 * A = [[1,1],[0,1]], h = (1,0), Q = diag(q1, q2) (BASELINE.json configs[2]; the reference has no
 * slope, causalimpact_lib.py:496).  prior[15] = m0, P0, obs_conc, obs_scale, obs_ub, lvl_conc,
 * lvl_scale, lvl_ub, slope_conc, slope_scale, slope_ub, m0_slope, P0_slope (ub on the SCALE).
 * Work arrays: r [T], S [5 T] (predicted a0, a1, P00, P01, P11 per step), rb [T].
 * ---------------------------------------------------------------------------------------- */
static double one_chain_llt(const double* y, const double* X, const double* Om, int T, int p,
                            const double* prior, const double* th, double* g, int with_prior,
                            double* r, double* S, double* rb) {
  const double* w = th;
  const double u = th[p], l = th[p + 1], z = th[p + 2];
  const double s_e = exp(u), q1 = exp(l), q2 = exp(z);
  int t, j;
  for (t = 0; t < T; ++t) {
    double acc = y[t];
    const double* x = X + (size_t)t * p;
    for (j = 0; j < p; ++j) acc -= x[j] * w[j];
    r[t] = acc;
  }
  double a0 = prior[0], a1 = prior[11], P00 = prior[1], P01 = 0.0, P11 = prior[12], ll = 0.0;
  for (t = 0; t < T; ++t) {
    double* s = S + 5 * (size_t)t;
    s[0] = a0; s[1] = a1; s[2] = P00; s[3] = P01; s[4] = P11;
    if (r[t] == r[t]) {
      const double v = r[t] - a0, F = P00 + s_e, K0 = P00 / F, K1 = P01 / F;
      ll += -0.5 * (LOG2PI + log(F) + v * v / F);
      a0 += K0 * v; a1 += K1 * v;
      const double n00 = P00 - K0 * P00, n01 = P01 - K0 * P01, n11 = P11 - K1 * P01;
      P00 = n00; P01 = n01; P11 = n11;
    }
    a0 += a1;
    const double m00 = P00 + 2.0 * P01 + P11 + q1, m01 = P01 + P11, m11 = P11 + q2;
    P00 = m00; P01 = m01; P11 = m11;
  }
  double val = ll;
  if (g) {
    double ab0 = 0.0, ab1 = 0.0, B00 = 0.0, B01 = 0.0, B10 = 0.0, B11 = 0.0;
    double ge = 0.0, gq1 = 0.0, gq2 = 0.0;
    for (j = 0; j < p + 3; ++j) g[j] = 0.0;
    for (t = T - 1; t >= 0; --t) {
      gq1 += B00; gq2 += B11;
      /* through the prediction: abar A, A' Pbar A */
      const double f0 = ab0, f1 = ab0 + ab1;
      const double M00 = B00, M01 = B00 + B01, M10 = B10, M11 = B10 + B11;
      const double C00 = M00, C01 = M01, C10 = M00 + M10, C11 = M01 + M11;
      if (!(r[t] == r[t])) {
        ab0 = f0; ab1 = f1; B00 = C00; B01 = C01; B10 = C10; B11 = C11;
        rb[t] = 0.0;
        continue;
      }
      const double* s = S + 5 * (size_t)t;
      const double v = r[t] - s[0], F = s[2] + s_e, K0 = s[2] / F, K1 = s[3] / F;
      double kb0 = f0 * v, kb1 = f1 * v;
      double vbar = K0 * f0 + K1 * f1;
      /* P+ = P - K (h'P) */
      double N00 = C00 - (K0 * C00 + K1 * C10), N01 = C01 - (K0 * C01 + K1 * C11);
      double N10 = C10, N11 = C11;
      kb0 -= C00 * s[2] + C01 * s[3];
      kb1 -= C10 * s[2] + C11 * s[3];
      /* K = P h / F */
      N00 += kb0 / F; N10 += kb1 / F;
      double Fbar = -(kb0 * K0 + kb1 * K1) / F;
      vbar -= v / F;
      Fbar -= 0.5 * (1.0 / F - v * v / (F * F));
      N00 += Fbar;
      ge += Fbar;
      rb[t] = vbar;
      ab0 = f0 - vbar; ab1 = f1;
      B00 = N00; B01 = N01; B10 = N10; B11 = N11;
    }
    for (t = 0; t < T; ++t) {
      if (rb[t] == 0.0) continue;
      const double* x = X + (size_t)t * p;
      for (j = 0; j < p; ++j) g[j] -= rb[t] * x[j];
    }
    g[p] = ge * s_e; g[p + 1] = gq1 * q1; g[p + 2] = gq2 * q2;
  }
  if (with_prior) {
    const double oc = prior[2], os = prior[3], oub = prior[4];
    const double lc = prior[5], ls = prior[6], lub = prior[7];
    const double zc = prior[8], zs = prior[9], zub = prior[10];
    double lp = -(oc + 1.0) * u - os / s_e + u - (lc + 1.0) * l - ls / q1 + l
                - (zc + 1.0) * z - zs / q2 + z;
    if (g) {
      g[p] += -(oc + 1.0) + os / s_e + 1.0;
      g[p + 1] += -(lc + 1.0) + ls / q1 + 1.0;
      g[p + 2] += -(zc + 1.0) + zs / q2 + 1.0;
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
    if (!(sqrt(s_e) <= oub && sqrt(q1) <= lub && sqrt(q2) <= zub)) val = -INFINITY;
  }
  return val;
}

/* theta [C, p+3] -> val [C], grad [C, p+3] (grad may be NULL).  Returns threads used. */
int ci_oracle_llt_logpost_grad(const double* y, const double* X, const double* Om, int T, int p,
                               const double* prior, const double* theta, int C, double* val,
                               double* grad, int with_prior, int nthreads) {
  int used = 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  used = omp_get_max_threads();
#pragma omp parallel
#endif
  {
    double* buf = (double*)malloc(sizeof(double) * 7 * (size_t)T);
    int c;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (c = 0; c < C; ++c)
      val[c] = one_chain_llt(y, X, Om, T, p, prior, theta + (size_t)c * (p + 3),
                             grad ? grad + (size_t)c * (p + 3) : NULL, with_prior, buf, buf + T,
                             buf + 6 * (size_t)T);
    free(buf);
  }
  return used;
}

/* ------------------------------------------------------------------------------------------
 * Posterior predictive draws (local level): restates oracle/smoother_np.py posterior_predict --
 * forward filter + backward sampling (the law of TFP's LGSSM posterior_sample inside
 * gibbs_sampler._resample_latents, call site causalimpact_lib.py:365-388) and the one-step
 * predictive draw of _get_posterior_means_and_trajectories (causalimpact_lib.py:609-632), with
 * the engine's Philox4x32-10 streams (oracle/philox_np.py predict_normals).
 * ---------------------------------------------------------------------------------------- */
#include <stdint.h>
static void philox4x32(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                       uint32_t out[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  int i;
  for (i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static void box_muller(uint32_t x0, uint32_t x1, double* z0, double* z1) {
  const double u = ((double)(x0 >> 8) + 0.5) * 5.9604644775390625e-08;
  const double w = ((double)(x1 >> 8) + 0.5) * 5.9604644775390625e-08;
  const double rad = sqrt(-2.0 * log(u));
  *z0 = rad * cos(6.283185307179586 * w); *z1 = rad * sin(6.283185307179586 * w);
}

/* theta [S, p+2] -> level [S,T] (may be NULL), traj [S,T], loc_sum [T] (sum over draws of
 * level + x.w; the caller divides by S).  prior: m0, P0 as above.  Returns threads used. */
int ci_oracle_predict(const double* y, const double* X, int T, int p, const double* prior,
                      const double* theta, int S, uint64_t seed, uint64_t draw_id0, double* level,
                      double* traj, double* loc_sum, int nthreads) {
  int used = 1, t;
  for (t = 0; t < T; ++t) loc_sum[t] = 0.0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  used = omp_get_max_threads();
#pragma omp parallel
#endif
  {
    double* buf = (double*)malloc(sizeof(double) * 5 * (size_t)T);
    double *xw = buf, *m = buf + T, *Cv = buf + 2 * (size_t)T, *lv = buf + 3 * (size_t)T,
           *acc = buf + 4 * (size_t)T;
    int s, tt, j;
    for (tt = 0; tt < T; ++tt) acc[tt] = 0.0;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (s = 0; s < S; ++s) {
      const double* th = theta + (size_t)s * (p + 2);
      const double s_e = exp(th[p]), s_h = exp(th[p + 1]), sig_e = sqrt(s_e);
      double a = prior[0], Pv = prior[1];
      for (tt = 0; tt < T; ++tt) {
        const double* x = X + (size_t)tt * p;
        double d = 0.0;
        for (j = 0; j < p; ++j) d += x[j] * th[j];
        xw[tt] = d;
        const double r = y[tt] - d;
        if (r == r) {
          const double F = Pv + s_e, K = Pv / F;
          a += K * (r - a); Pv *= (1.0 - K);
        }
        m[tt] = a; Cv[tt] = Pv;
        Pv += s_h;
      }
      const uint64_t gid = draw_id0 + (uint64_t)s;
      const uint32_t c0 = (uint32_t)gid, c1 = 4u | ((uint32_t)(gid >> 32) << 8);   /* RNG_SMOOTH */
      double xn = 0.0;
      double* out_t = traj + (size_t)s * T;
      for (tt = T - 1; tt >= 0; --tt) {
        uint32_t o[4];
        double z0, z1, z2, z3;
        philox4x32(seed, c0, c1, (uint32_t)(tt >> 1), 0u, o);
        box_muller(o[0], o[1], &z0, &z1);
        box_muller(o[2], o[3], &z2, &z3);
        const double zs = (tt & 1) ? z2 : z0, zp = (tt & 1) ? z3 : z1;
        double xv;
        if (tt == T - 1) xv = m[tt] + sqrt(Cv[tt]) * zs;
        else {
          const double J = Cv[tt] / (Cv[tt] + s_h);
          xv = m[tt] + J * (xn - m[tt]) + sqrt(Cv[tt] * (1.0 - J)) * zs;
        }
        xn = xv; lv[tt] = xv;
        out_t[tt] = xv + xw[tt] + sig_e * zp;
        acc[tt] += xv + xw[tt];
      }
      if (level) memcpy(level + (size_t)s * T, lv, sizeof(double) * (size_t)T);
    }
#ifdef _OPENMP
#pragma omp critical
#endif
    for (tt = 0; tt < T; ++tt) loc_sum[tt] += acc[tt];
    free(buf);
  }
  return used;
}
