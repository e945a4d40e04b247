// ci_llt.cuh -- local linear trend (state = (level, slope), d = 2): Kalman
// log-likelihood + adjoint gradient, time-parallel inside a warp.
//
// The reference model has NO slope (causalimpact/causalimpact_lib.py:496,
// slope_scale = 0 at :373-374); BASELINE.json config 3 asks for it as a
// capability extension.  Conventions are TFP's LocalLinearTrend: level_t =
// level_{t-1} + slope_{t-1} + N(0, q1), slope_t = slope_{t-1} + N(0, q2),
// y_t = level_t + x_t.w + N(0, s_e); prior N((m0, m0s), diag(P0, P0s)) on the
// state AT t = 0; update-then-predict.  Oracle: oracle/kalman_np.py gen_filter*.
//
// Forward: the associative filtering elements of Sarkka & Garcia-Fernandez
// (IEEE TAC 2021), (A, b, C, eta, J) with 2x2 blocks, composed per lane over its
// 8 steps, scanned across lanes with shuffles, applied to the carry-in filtered
// state; then each lane replays its 8 steps sequentially for (K, 1/F, v).
// Backward: with G_t = A (I - K_t h') the adjoints obey
//   abar_t = G_t' abar_{t+1} + h v_t/F_t
//   Pbar_t = G_t' Pbar_{t+1} G_t + sym(G_t' abar_{t+1} h') v_t/F_t + dF_t h h'
// (Pbar kept symmetric), i.e. affine / congruence maps sharing the multiplier
// G, scanned in reverse.  Validated in float64 against the generic oracle
// (oracle/scan_np.py:llt_*, tests/test_oracle_kalman.py).
#pragma once
#include "ci_device.cuh"

namespace ci {

template <typename R> struct El2 {
  R a00, a01, a10, a11, b0, b1, c00, c01, c11, e0, e1, j00, j01, j11;
};

template <typename R> __device__ __forceinline__ El2<R> el2_identity() {
  El2<R> e;
  e.a00 = 1; e.a01 = 0; e.a10 = 0; e.a11 = 1; e.b0 = 0; e.b1 = 0; e.c00 = 0; e.c01 = 0; e.c11 = 0;
  e.e0 = 0; e.e1 = 0; e.j00 = 0; e.j01 = 0; e.j11 = 0;
  return e;
}

// i = earlier segment, j = later segment
template <typename R>
__device__ __forceinline__ El2<R> el2_combine(const El2<R>& i, const El2<R>& j) {
  // D = I + C_i J_j
  const R d00 = (R)1 + i.c00 * j.j00 + i.c01 * j.j01, d01 = i.c00 * j.j01 + i.c01 * j.j11;
  const R d10 = i.c01 * j.j00 + i.c11 * j.j01, d11 = (R)1 + i.c01 * j.j01 + i.c11 * j.j11;
  const R rdet = Num<R>::rcp(d00 * d11 - d01 * d10);
  const R n00 = d11 * rdet, n01 = -d01 * rdet, n10 = -d10 * rdet, n11 = d00 * rdet;   // D^-1
  // X = A_j D^-1
  const R x00 = j.a00 * n00 + j.a01 * n10, x01 = j.a00 * n01 + j.a01 * n11;
  const R x10 = j.a10 * n00 + j.a11 * n10, x11 = j.a10 * n01 + j.a11 * n11;
  El2<R> o;
  o.a00 = x00 * i.a00 + x01 * i.a10; o.a01 = x00 * i.a01 + x01 * i.a11;
  o.a10 = x10 * i.a00 + x11 * i.a10; o.a11 = x10 * i.a01 + x11 * i.a11;
  const R t0 = i.b0 + i.c00 * j.e0 + i.c01 * j.e1, t1 = i.b1 + i.c01 * j.e0 + i.c11 * j.e1;
  o.b0 = x00 * t0 + x01 * t1 + j.b0; o.b1 = x10 * t0 + x11 * t1 + j.b1;
  // X C_i A_j' + C_j
  const R y00 = x00 * i.c00 + x01 * i.c01, y01 = x00 * i.c01 + x01 * i.c11;
  const R y10 = x10 * i.c00 + x11 * i.c01, y11 = x10 * i.c01 + x11 * i.c11;
  o.c00 = y00 * j.a00 + y01 * j.a01 + j.c00;
  o.c01 = y00 * j.a10 + y01 * j.a11 + j.c01;
  o.c11 = y10 * j.a10 + y11 * j.a11 + j.c11;
  // Y = A_i' D^-T
  const R z00 = i.a00 * n00 + i.a10 * n01, z01 = i.a00 * n10 + i.a10 * n11;
  const R z10 = i.a01 * n00 + i.a11 * n01, z11 = i.a01 * n10 + i.a11 * n11;
  const R s0 = j.e0 - (j.j00 * i.b0 + j.j01 * i.b1), s1 = j.e1 - (j.j01 * i.b0 + j.j11 * i.b1);
  o.e0 = z00 * s0 + z01 * s1 + i.e0; o.e1 = z10 * s0 + z11 * s1 + i.e1;
  // Y J_j A_i + J_i
  const R w00 = z00 * j.j00 + z01 * j.j01, w01 = z00 * j.j01 + z01 * j.j11;
  const R w10 = z10 * j.j00 + z11 * j.j01, w11 = z10 * j.j01 + z11 * j.j11;
  o.j00 = w00 * i.a00 + w01 * i.a10 + i.j00;
  o.j01 = w00 * i.a01 + w01 * i.a11 + i.j01;
  o.j11 = w10 * i.a01 + w11 * i.a11 + i.j11;
  return o;
}

template <typename R> __device__ __forceinline__ El2<R> el2_shfl_up(const El2<R>& e, int off) {
  El2<R> o;
  o.a00 = __shfl_up_sync(FULL, e.a00, off); o.a01 = __shfl_up_sync(FULL, e.a01, off);
  o.a10 = __shfl_up_sync(FULL, e.a10, off); o.a11 = __shfl_up_sync(FULL, e.a11, off);
  o.b0 = __shfl_up_sync(FULL, e.b0, off); o.b1 = __shfl_up_sync(FULL, e.b1, off);
  o.c00 = __shfl_up_sync(FULL, e.c00, off); o.c01 = __shfl_up_sync(FULL, e.c01, off);
  o.c11 = __shfl_up_sync(FULL, e.c11, off);
  o.e0 = __shfl_up_sync(FULL, e.e0, off); o.e1 = __shfl_up_sync(FULL, e.e1, off);
  o.j00 = __shfl_up_sync(FULL, e.j00, off); o.j01 = __shfl_up_sync(FULL, e.j01, off);
  o.j11 = __shfl_up_sync(FULL, e.j11, off);
  return o;
}

// filtered state carried across tiles: mean (b0,b1), covariance (c00,c01,c11)
template <typename R> struct St2 { R b0, b1, c00, c01, c11; };

// apply element e (a later segment) to a filtered state s
template <typename R>
__device__ __forceinline__ St2<R> el2_apply(const El2<R>& e, const St2<R>& s) {
  const R d00 = (R)1 + s.c00 * e.j00 + s.c01 * e.j01, d01 = s.c00 * e.j01 + s.c01 * e.j11;
  const R d10 = s.c01 * e.j00 + s.c11 * e.j01, d11 = (R)1 + s.c01 * e.j01 + s.c11 * e.j11;
  const R rdet = Num<R>::rcp(d00 * d11 - d01 * d10);
  const R n00 = d11 * rdet, n01 = -d01 * rdet, n10 = -d10 * rdet, n11 = d00 * rdet;
  const R x00 = e.a00 * n00 + e.a01 * n10, x01 = e.a00 * n01 + e.a01 * n11;
  const R x10 = e.a10 * n00 + e.a11 * n10, x11 = e.a10 * n01 + e.a11 * n11;
  const R t0 = s.b0 + s.c00 * e.e0 + s.c01 * e.e1, t1 = s.b1 + s.c01 * e.e0 + s.c11 * e.e1;
  St2<R> o;
  o.b0 = x00 * t0 + x01 * t1 + e.b0; o.b1 = x10 * t0 + x11 * t1 + e.b1;
  const R y00 = x00 * s.c00 + x01 * s.c01, y01 = x00 * s.c01 + x01 * s.c11;
  const R y10 = x10 * s.c00 + x11 * s.c01, y11 = x10 * s.c01 + x11 * s.c11;
  o.c00 = y00 * e.a00 + y01 * e.a01 + e.c00;
  o.c01 = y00 * e.a10 + y01 * e.a11 + e.c01;
  o.c11 = y10 * e.a10 + y11 * e.a11 + e.c11;
  return o;
}

// slope-related problem constants (ProbDev carries the level ones)
template <typename R> struct LltDev {
  R q_conc, q_scale, q_ub;   // InvGamma on sigma_slope^2, bound on the slope VARIANCE
  R m0s, P0s;                // initial slope ~ N(m0s, P0s)
};

// per-tile, per-lane filter path of the trend model
template <typename R> struct Blk2 {
  R r[KS];              // residual (0 where masked)
  R K1[KS], K2[KS];     // gain
  R rF[KS], v[KS];      // 1/F, innovation  (0 where masked)
  uint32_t obs;
};

// Forward over one tile.  st = filtered state after the previous tile's last
// step (ignored for tile 0, whose first step carries the prior).
template <typename R>
__device__ __forceinline__ void llt_forward(Blk2<R>& B, R s_e, R q1, R q2, const ProbDev<R>& pr,
                                            const LltDev<R>& ld2, St2<R>& st, bool first_tile,
                                            int lane, R& ll_terms) {
  const R rS = Num<R>::rcp(q1 + s_e), g = s_e * rS, kap = q1 * rS;
  // ---- lane aggregate: combine the 8 step elements ----
  El2<R> E;
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const bool o = (B.obs >> k) & 1u;
    El2<R> s;
    if (first_tile && lane == 0 && k == 0) {
      // prior N((m0,m0s), diag(P0,P0s)) updated with y_0
      const R rS0 = o ? Num<R>::rcp(pr.P0 + s_e) : (R)0;
      const R k0 = pr.P0 * rS0;
      s.a00 = 0; s.a01 = 0; s.a10 = 0; s.a11 = 0;
      s.b0 = pr.m0 + k0 * (B.r[0] - (o ? pr.m0 : (R)0)); s.b1 = ld2.m0s;
      s.c00 = pr.P0 - k0 * pr.P0; s.c01 = 0; s.c11 = ld2.P0s;
      s.e0 = 0; s.e1 = 0; s.j00 = 0; s.j01 = 0; s.j11 = 0;
    } else if (o) {
      s.a00 = g; s.a01 = g; s.a10 = 0; s.a11 = 1;
      s.b0 = kap * B.r[k]; s.b1 = 0;
      s.c00 = q1 * g; s.c01 = 0; s.c11 = q2;
      s.e0 = B.r[k] * rS; s.e1 = s.e0;
      s.j00 = rS; s.j01 = rS; s.j11 = rS;
    } else {
      s.a00 = 1; s.a01 = 1; s.a10 = 0; s.a11 = 1; s.b0 = 0; s.b1 = 0;
      s.c00 = q1; s.c01 = 0; s.c11 = q2; s.e0 = 0; s.e1 = 0; s.j00 = 0; s.j01 = 0; s.j11 = 0;
    }
    E = (k == 0) ? s : el2_combine(E, s);
  }
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const El2<R> O = el2_shfl_up(E, off);
    if (lane >= off) E = el2_combine(O, E);
  }
  El2<R> X = el2_shfl_up(E, 1);
  if (lane == 0) X = el2_identity<R>();
  // filtered state just before this lane's first step
  St2<R> f = el2_apply(X, st);
  if (!first_tile && lane == 0) f = st;
  const bool at_prior = first_tile && lane == 0;
  R lt = 0;
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    R a0, a1, P00, P01, P11;
    if (at_prior && k == 0) {
      a0 = pr.m0; a1 = ld2.m0s; P00 = pr.P0; P01 = 0; P11 = ld2.P0s;
    } else {
      a0 = f.b0 + f.b1; a1 = f.b1;
      P00 = f.c00 + (R)2 * f.c01 + f.c11 + q1; P01 = f.c01 + f.c11; P11 = f.c11 + q2;
    }
    const bool o = (B.obs >> k) & 1u;
    const R F = P00 + s_e;
    const R rF = o ? Num<R>::rcp(F) : (R)0;
    const R v = o ? (B.r[k] - a0) : (R)0;
    const R K1 = P00 * rF, K2 = P01 * rF;
    B.K1[k] = K1; B.K2[k] = K2; B.rF[k] = rF; B.v[k] = v;
    if (o) lt += Num<R>::log(F) + v * v * rF;
    f.b0 = fma(K1, v, a0); f.b1 = fma(K2, v, a1);
    f.c00 = fma(-K1, P00, P00); f.c01 = fma(-K1, P01, P01); f.c11 = fma(-K2, P01, P11);
  }
  ll_terms = lt;
  st.b0 = __shfl_sync(FULL, f.b0, 31); st.b1 = __shfl_sync(FULL, f.b1, 31);
  st.c00 = __shfl_sync(FULL, f.c00, 31); st.c01 = __shfl_sync(FULL, f.c01, 31);
  st.c11 = __shfl_sync(FULL, f.c11, 31);
}

// reverse-scan element shared by abar (G, c) and Pbar (G, E)
template <typename R> struct Rv2 { R g00, g01, g10, g11, c0, c1, c2; };

template <typename R> __device__ __forceinline__ Rv2<R> rv2_shfl_down(const Rv2<R>& e, int off) {
  Rv2<R> o;
  o.g00 = __shfl_down_sync(FULL, e.g00, off); o.g01 = __shfl_down_sync(FULL, e.g01, off);
  o.g10 = __shfl_down_sync(FULL, e.g10, off); o.g11 = __shfl_down_sync(FULL, e.g11, off);
  o.c0 = __shfl_down_sync(FULL, e.c0, off); o.c1 = __shfl_down_sync(FULL, e.c1, off);
  o.c2 = __shfl_down_sync(FULL, e.c2, off);
  return o;
}
// composite of an EARLY segment e followed (in time) by a LATE segment l:
//   G = G_l G_e ;  vector form  c = G_e' c_l + c_e ;  sym form  E = G_e' E_l G_e + E_e
template <typename R, bool SYM>
__device__ __forceinline__ Rv2<R> rv2_combine(const Rv2<R>& e, const Rv2<R>& l) {
  Rv2<R> o;
  o.g00 = l.g00 * e.g00 + l.g01 * e.g10; o.g01 = l.g00 * e.g01 + l.g01 * e.g11;
  o.g10 = l.g10 * e.g00 + l.g11 * e.g10; o.g11 = l.g10 * e.g01 + l.g11 * e.g11;
  if (!SYM) {
    o.c0 = e.g00 * l.c0 + e.g10 * l.c1 + e.c0;
    o.c1 = e.g01 * l.c0 + e.g11 * l.c1 + e.c1;
    o.c2 = 0;
  } else {   // (c0, c1, c2) = (E00, E01, E11)
    const R t00 = l.c0 * e.g00 + l.c1 * e.g10, t01 = l.c0 * e.g01 + l.c1 * e.g11;
    const R t10 = l.c1 * e.g00 + l.c2 * e.g10, t11 = l.c1 * e.g01 + l.c2 * e.g11;
    o.c0 = e.g00 * t00 + e.g10 * t10 + e.c0;
    o.c1 = e.g00 * t01 + e.g10 * t11 + e.c1;
    o.c2 = e.g01 * t01 + e.g11 * t11 + e.c2;
  }
  return o;
}

// Adjoint sweep over one tile.  ab = (abar0, abar1), pb = (Pbar00, Pbar01, Pbar11)
// of the step just after the tile in / of the tile's first step out.
template <typename R>
__device__ __forceinline__ void llt_backward(const Blk2<R>& B, R s_e, R (&ab)[2], R (&pb)[3],
                                             int lane, R& ge, R& gq1, R& gq2, R* rbar) {
  // ---- abar: lane composite, reverse scan ----
  Rv2<R> L; L.g00 = 1; L.g01 = 0; L.g10 = 0; L.g11 = 1; L.c0 = 0; L.c1 = 0; L.c2 = 0;
#pragma unroll
  for (int k = KS - 1; k >= 0; --k) {
    Rv2<R> s; s.g00 = (R)1 - B.K1[k] - B.K2[k]; s.g01 = 1; s.g10 = -B.K2[k]; s.g11 = 1;
    s.c0 = B.v[k] * B.rF[k]; s.c1 = 0; s.c2 = 0;
    L = rv2_combine<R, false>(s, L);
  }
  Rv2<R> S = L;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const Rv2<R> O = rv2_shfl_down(S, off);
    if (lane + off < 32) S = rv2_combine<R, false>(S, O);
  }
  Rv2<R> X = rv2_shfl_down(S, 1);
  if (lane == 31) { X.g00 = 1; X.g01 = 0; X.g10 = 0; X.g11 = 1; X.c0 = 0; X.c1 = 0; X.c2 = 0; }
  R a0 = X.g00 * ab[0] + X.g10 * ab[1] + X.c0;      // abar after this lane's last step
  R a1 = X.g01 * ab[0] + X.g11 * ab[1] + X.c1;
  R n0[KS], n1[KS];
#pragma unroll
  for (int k = KS - 1; k >= 0; --k) {
    n0[k] = a0; n1[k] = a1;
    const R u0 = ((R)1 - B.K1[k] - B.K2[k]) * a0 - B.K2[k] * a1, u1 = a0 + a1;
    a0 = u0 + B.v[k] * B.rF[k]; a1 = u1;
  }
  ab[0] = __shfl_sync(FULL, a0, 0); ab[1] = __shfl_sync(FULL, a1, 0);
  // ---- Pbar: congruence elements ----
  R e00[KS], e01[KS], dF[KS];
  L.g00 = 1; L.g01 = 0; L.g10 = 0; L.g11 = 1; L.c0 = 0; L.c1 = 0; L.c2 = 0;
#pragma unroll
  for (int k = KS - 1; k >= 0; --k) {
    const R g00 = (R)1 - B.K1[k] - B.K2[k], g10 = -B.K2[k];
    const R u0 = g00 * n0[k] + g10 * n1[k], u1 = n0[k] + n1[k];
    const R vr = B.v[k] * B.rF[k];
    dF[k] = (R)-0.5 * (B.rF[k] - vr * vr);
    e00[k] = fma(u0, vr, dF[k]); e01[k] = (R)0.5 * u1 * vr;
    Rv2<R> s; s.g00 = g00; s.g01 = 1; s.g10 = g10; s.g11 = 1; s.c0 = e00[k]; s.c1 = e01[k]; s.c2 = 0;
    L = rv2_combine<R, true>(s, L);
  }
  S = L;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const Rv2<R> O = rv2_shfl_down(S, off);
    if (lane + off < 32) S = rv2_combine<R, true>(S, O);
  }
  X = rv2_shfl_down(S, 1);
  if (lane == 31) { X.g00 = 1; X.g01 = 0; X.g10 = 0; X.g11 = 1; X.c0 = 0; X.c1 = 0; X.c2 = 0; }
  R p00, p01, p11;
  {   // X.G' pb X.G + X.E
    const R t00 = pb[0] * X.g00 + pb[1] * X.g10, t01 = pb[0] * X.g01 + pb[1] * X.g11;
    const R t10 = pb[1] * X.g00 + pb[2] * X.g10, t11 = pb[1] * X.g01 + pb[2] * X.g11;
    p00 = X.g00 * t00 + X.g10 * t10 + X.c0;
    p01 = X.g00 * t01 + X.g10 * t11 + X.c1;
    p11 = X.g01 * t01 + X.g11 * t11 + X.c2;
  }
  R lge = 0, l1 = 0, l2 = 0;
#pragma unroll
  for (int k = KS - 1; k >= 0; --k) {
    const R K1 = B.K1[k], K2 = B.K2[k];
    l1 += p00; l2 += p11;                         // d/dq of P_{t+1} = ... + Q
    // abar+ = A' abar_{t+1},  Pbar+ = A' Pbar_{t+1} A
    const R ap0 = n0[k], ap1 = n0[k] + n1[k];
    const R q00 = p00, q01 = p00 + p01, q11 = p00 + (R)2 * p01 + p11;
    const R vr = B.v[k] * B.rF[k];
    rbar[k] = K1 * ap0 + K2 * ap1 - vr;
    lge += dF[k] - (ap0 * K1 + ap1 * K2) * vr + K1 * (q00 * K1 + q01 * K2) + K2 * (q01 * K1 + q11 * K2);
    // Pbar_t = G' Pbar_{t+1} G + E_t,  G = [[g00, 1], [g10, 1]]
    const R g00 = (R)1 - K1 - K2, g10 = -K2;
    const R n00 = g00 * g00 * p00 + (R)2 * g00 * g10 * p01 + g10 * g10 * p11 + e00[k];
    const R n01 = g00 * (p00 + p01) + g10 * (p01 + p11) + e01[k];
    const R n11 = p00 + (R)2 * p01 + p11;
    p00 = n00; p01 = n01; p11 = n11;
  }
  pb[0] = __shfl_sync(FULL, p00, 0); pb[1] = __shfl_sync(FULL, p01, 0);
  pb[2] = __shfl_sync(FULL, p11, 0);
  ge += lge; gq1 += l1; gq2 += l2;
}

// Whole-chain evaluation (one warp, tiles walked sequentially, any T).
template <typename R>
__device__ __forceinline__ void chain_eval_llt(TilePipe<R>& pipe, const ProbDev<R>& pr,
                                               const LltDev<R>& ld2, const WarpScratch<R>& ws,
                                               R s_e, R q1, R q2, bool want_grad, int lane,
                                               double& ll, double& g_se, double& g_q1,
                                               double& g_q2, R (&gw)[JS]) {
  const int p = pr.p, ld = pr.ld, NB = pr.NB;
  double acc = 0.0;
  int n_obs = 0;
  St2<R> st; st.b0 = 0; st.b1 = 0; st.c00 = 0; st.c01 = 0; st.c11 = 0;
  R* ck = ws.ckpt;                      // 5 values per tile (needs ckpt size >= 5*NB)
  for (int b = 0; b < NB; ++b) {
    const R* tile = pipe.acquire(b);
    Blk<R> B1;
    blk_residuals(B1, tile, ws.w, p, ld, lane);
    Blk2<R> B;
#pragma unroll
    for (int k = 0; k < KS; ++k) B.r[k] = B1.r[k];
    B.obs = B1.obs;
    if (want_grad && lane == 0) {
      ck[5 * b] = st.b0; ck[5 * b + 1] = st.b1; ck[5 * b + 2] = st.c00; ck[5 * b + 3] = st.c01;
      ck[5 * b + 4] = st.c11;
    }
    R lt;
    llt_forward(B, s_e, q1, q2, pr, ld2, st, b == 0, lane, lt);
    acc += (double)lt;
    n_obs += __popc(B.obs);
    pipe.release(lane);
  }
  acc += 1.8378770664093453 * (double)n_obs;
  ll = -0.5 * warp_sum(acc);
  g_se = 0.0; g_q1 = 0.0; g_q2 = 0.0;
#pragma unroll
  for (int s = 0; s < JS; ++s) gw[s] = 0;
  if (!want_grad) return;
  __syncwarp();
  const XtMap xm = xt_map(p, lane);
  const bool small_p = p <= PSMALL;
  R accw[PSMALL];
#pragma unroll
  for (int j = 0; j < PSMALL; ++j) accw[j] = 0;
  R g1[1] = {0};
  R ab[2] = {0, 0}, pb[3] = {0, 0, 0};
  double ge = 0.0, g1d = 0.0, g2d = 0.0;
  for (int b = NB - 1; b >= 0; --b) {
    const R* tile = pipe.acquire(b);
    Blk<R> B1;
    blk_residuals(B1, tile, ws.w, p, ld, lane);
    Blk2<R> B;
#pragma unroll
    for (int k = 0; k < KS; ++k) B.r[k] = B1.r[k];
    B.obs = B1.obs;
    st.b0 = ck[5 * b]; st.b1 = ck[5 * b + 1]; st.c00 = ck[5 * b + 2]; st.c01 = ck[5 * b + 3];
    st.c11 = ck[5 * b + 4];
    R lt;
    llt_forward(B, s_e, q1, q2, pr, ld2, st, b == 0, lane, lt);
    R rbar[KS];
    R lge = 0, l1 = 0, l2 = 0;
    llt_backward(B, s_e, ab, pb, lane, lge, l1, l2, rbar);
    ge += (double)lge; g1d += (double)l1; g2d += (double)l2;
    if (small_p) {
      blk_xt_rbar_small(tile, rbar, p, ld, lane, accw);
    } else {
#pragma unroll
      for (int k = 0; k < KS; ++k) ws.rbuf[lane * KS + k + (lane >> 2)] = rbar[k];
      __syncwarp();
      if (p <= 32) blk_xt_rbar<R, 1>(tile, ws.rbuf, p, ld, xm.jj, xm.part, xm.nparts, g1);
      else blk_xt_rbar<R, JS>(tile, ws.rbuf, p, ld, xm.jj, xm.part, xm.nparts, gw);
      __syncwarp();
    }
    pipe.release(lane);
  }
  g_se = warp_sum(ge); g_q1 = warp_sum(g1d); g_q2 = warp_sum(g2d);
  if (small_p) {
#pragma unroll
    for (int j = 0; j < PSMALL; ++j) {
      if (j < p) {
        const R tot = warp_sum(accw[j]);
        if (lane == j) gw[0] = -tot;
      }
    }
  } else {
    if (p <= 32) gw[0] = g1[0];
#pragma unroll
    for (int s = 0; s < JS; ++s) {
      R a = gw[s];
      for (int o = xm.PJ; o < 32; o <<= 1) a += __shfl_xor_sync(FULL, a, o);
      gw[s] = -a;
    }
  }
}

// slope prior + Jacobian (extension; mirrors the level prior of lib.py:424-432)
template <typename R>
__device__ __forceinline__ double llt_slope_prior(const LltDev<R>& ld2, R s, R q2, double& g_s) {
  const R rq = (R)1 / q2;
  g_s += -((double)ld2.q_conc + 1.0) + (double)ld2.q_scale * rq + 1.0;
  const bool ok = (q2 <= ld2.q_ub) && (s == s);
  const double lp = -((double)ld2.q_conc + 1.0) * s - (double)ld2.q_scale * rq + s;
  return ok ? lp : -CUDART_INF;
}

}  // namespace ci
