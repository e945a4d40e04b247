// ci_llt_predict.cuh -- K4 for the local linear trend (state = (level, slope), d = 2): simulation
// smoother (forward filter + backward sampling) and the one-step predictive draw, one warp per
// posterior draw.  Completes BASELINE.json configs[2] end to end; the reference itself has no
// slope (causalimpact/causalimpact_lib.py:496), so what this replaces is the same pair of
// reference steps as ci_predict.cuh -- the latent resampling inside the sampler (lib.py:365-388)
// and _get_posterior_means_and_trajectories (lib.py:609-632) -- for the extended model.
//
// Forward sweep: ci_llt.cuh's filtering elements, one 5-value checkpoint (filtered mean and
// covariance) per tile.  Backward sweep, per tile: replay the forward scan from the checkpoint; at
// every step turn the filtered moments (m_t, C_t) into the sampling map
//     x_t = J_t x_{t+1} + (m_t - J_t A m_t) + chol(C_t - J_t R_t J_t') z_t,
//     R_t = A C_t A' + Q,  J_t = C_t A' R_t^-1          (J = 0 at the last step)
// -- an AFFINE map of x_{t+1} with a 2x2 multiplier -- and resolve the whole tile with the reverse
// matrix-affine scan the adjoint already uses (Rv2, ci_llt.cuh).  oracle/smoother_np.py
// (posterior_predict_llt) restates it with the same Philox streams: one call per step, counter
// word c3 = 1, normals (z_state0, z_state1, z_predict, unused).
#pragma once
#include "ci_llt.cuh"
#include "ci_kernels.cuh"

namespace ci {

// Forward over one tile like llt_forward, but instead of the likelihood terms it leaves, per
// step, the sampling map (J, o) of the backward recursion.  zs0 / zs1: the step's state normals.
template <typename R>
__device__ __forceinline__ void llt_forward_sampling(const Blk2<R>& B, R s_e, R q1, R q2,
                                                     const ProbDev<R>& pr, const LltDev<R>& ld2,
                                                     St2<R>& st, bool first_tile, int lane, int t0,
                                                     const R (&zs0)[KS], const R (&zs1)[KS],
                                                     R (&j00)[KS], R (&j01)[KS], R (&j10)[KS],
                                                     R (&j11)[KS], R (&o0)[KS], R (&o1)[KS]) {
  const R rS = Num<R>::rcp(q1 + s_e), g = s_e * rS, kap = q1 * rS;
  El2<R> E;
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const bool o = (B.obs >> k) & 1u;
    El2<R> s;
    if (first_tile && lane == 0 && k == 0) {
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
  St2<R> f = el2_apply(X, st);
  if (!first_tile && lane == 0) f = st;
  const bool at_prior = first_tile && lane == 0;
  const int T = pr.T;
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
    const R rF = o ? Num<R>::rcp(P00 + s_e) : (R)0;
    const R v = o ? (B.r[k] - a0) : (R)0;
    const R K1 = P00 * rF, K2 = P01 * rF;
    f.b0 = fma(K1, v, a0); f.b1 = fma(K2, v, a1);
    f.c00 = fma(-K1, P00, P00); f.c01 = fma(-K1, P01, P01); f.c11 = fma(-K2, P01, P11);
    // ---- sampling map of this step from the filtered moments (m, C) = (f.b, f.c) ----
    const R m0 = f.b0, m1 = f.b1, c00 = f.c00, c01 = f.c01, c11 = f.c11;
    // With R = A C A' + Q and W = Q R^-1 the textbook forms J = C A' R^-1, V = C - J R J',
    // o = m - J A m are rewritten WITHOUT the differences of nearly equal quantities that lose
    // every float32 digit over a long forecast (C grows with t while V stays at the scale of Q):
    //     J = A^-1 (I - W),   o = A^-1 W (A m),   V = A^-1 (Q - W Q) A^-T
    // (identical in exact arithmetic; the float64 oracle keeps the textbook forms).
    R J00 = 0, J01 = 0, J10 = 0, J11 = 0, V00 = c00, V01 = c01, V11 = c11;
    R od0 = m0, od1 = m1;
    if (t0 + k < T - 1) {
      const R R00 = c00 + (R)2 * c01 + c11 + q1, R01 = c01 + c11, R11 = c11 + q2;   // A C A' + Q
      const R rdet = Num<R>::rcp(R00 * R11 - R01 * R01);
      const R i00 = R11 * rdet, i01 = -R01 * rdet, i11 = R00 * rdet;                 // R^-1
      const R w00 = q1 * i00, w01 = q1 * i01, w10 = q2 * i01, w11 = q2 * i11;        // W = Q R^-1
      const R n00 = (R)1 - w00, n01 = -w01, n10 = -w10, n11 = (R)1 - w11;            // I - W
      J00 = n00 - n10; J01 = n01 - n11; J10 = n10; J11 = n11;                        // A^-1 (I - W)
      const R am0 = m0 + m1, am1 = m1;                                               // A m
      const R u0 = w00 * am0 + w01 * am1, u1 = w10 * am0 + w11 * am1;                // W A m
      od0 = u0 - u1; od1 = u1;
      const R S00 = q1 * n00, S01 = -q1 * w10, S11 = q2 * n11;                       // Q - W Q (symmetric)
      V00 = S00 - (R)2 * S01 + S11; V01 = S01 - S11; V11 = S11;
    }
    const R l00 = Num<R>::sqrt(V00 > (R)0 ? V00 : (R)0);
    const R l10 = l00 > (R)0 ? V01 / l00 : (R)0;
    const R d11 = V11 - l10 * l10;
    const R l11 = Num<R>::sqrt(d11 > (R)0 ? d11 : (R)0);
    j00[k] = J00; j01[k] = J01; j10[k] = J10; j11[k] = J11;
    o0[k] = od0 + l00 * zs0[k];
    o1[k] = od1 + l10 * zs0[k] + l11 * zs1[k];
  }
  st.b0 = __shfl_sync(FULL, f.b0, 31); st.b1 = __shfl_sync(FULL, f.b1, 31);
  st.c00 = __shfl_sync(FULL, f.c00, 31); st.c01 = __shfl_sync(FULL, f.c01, 31);
  st.c11 = __shfl_sync(FULL, f.c11, 31);
}

template <typename R>
__global__ void __launch_bounds__(32 * (MAXG + 1), 1)
k_predict_llt(ProbDev<R> pr, LltDev<R> ld2, SmemCfg cfg, const R* __restrict__ theta, int S,
              uint64_t seed, uint64_t draw_id0, R* __restrict__ level, R* __restrict__ slope,
              R* __restrict__ traj) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (blockDim.x >> 5) - 1;
  const int s0 = blockIdx.x * G;
  const int nactive = min(G, S - s0);
  const CtaShared<R> cs = cta_prologue(smem, cfg, pr, nactive);
  if (warp == G) {
    if (lane == 0)
      tile_producer(pr.tiles, cs.stage0, cs.full, cs.empty, cfg.stage_elems, cfg.nstage, pr.NB,
                    cfg.resident != 0, 2LL, [](long long s) { return (s & 1) == 0; });
    return;
  }
  if (warp >= nactive) return;
  const int s = s0 + warp;
  const int p = pr.p, dim = pr.dim, ld = pr.ld, NB = pr.NB, T = pr.T;
  const R* th = theta + (size_t)s * dim;
  const WarpScratch<R> ws = warp_scratch<R>(smem, cfg, warp);
  for (int j = lane; j < p; j += 32) ws.w[j] = th[j];
  const R s_e = Num<R>::exp(th[p]), q1 = Num<R>::exp(th[p + 1]), q2 = Num<R>::exp(th[p + 2]);
  const R sig_e = Num<R>::sqrt(s_e);
  __syncwarp();
  TilePipe<R> pipe = make_pipe(cs, cfg);
  R* ck = ws.ckpt;                        // 5 values per tile

  // ---- forward sweep: checkpoints only ----
  St2<R> st; st.b0 = 0; st.b1 = 0; st.c00 = 0; st.c01 = 0; st.c11 = 0;
  for (int b = 0; b < NB; ++b) {
    const R* tile = pipe.acquire(b);
    Blk<R> B1;
    blk_residuals(B1, tile, ws.w, p, ld, lane);
    Blk2<R> B;
#pragma unroll
    for (int k = 0; k < KS; ++k) B.r[k] = B1.r[k];
    B.obs = B1.obs;
    if (lane == 0) {
      ck[5 * b] = st.b0; ck[5 * b + 1] = st.b1; ck[5 * b + 2] = st.c00; ck[5 * b + 3] = st.c01;
      ck[5 * b + 4] = st.c11;
    }
    R lt;
    llt_forward(B, s_e, q1, q2, pr, ld2, st, b == 0, lane, lt);
    pipe.release(lane);
  }
  __syncwarp();

  // ---- backward sweep: sample ----
  const uint64_t gid = draw_id0 + (uint64_t)s;
  const uint32_t c0 = (uint32_t)gid, c1 = RNG_SMOOTH | ((uint32_t)(gid >> 32) << 8);
  R xc0 = 0, xc1 = 0;
  for (int b = NB - 1; b >= 0; --b) {
    const R* tile = pipe.acquire(b);
    Blk<R> B1;
    R xw[KS];
    blk_residuals_xw(B1, xw, tile, ws.w, p, ld, lane);
    Blk2<R> B;
#pragma unroll
    for (int k = 0; k < KS; ++k) B.r[k] = B1.r[k];
    B.obs = B1.obs;
    st.b0 = ck[5 * b]; st.b1 = ck[5 * b + 1]; st.c00 = ck[5 * b + 2]; st.c01 = ck[5 * b + 3];
    st.c11 = ck[5 * b + 4];
    const int t0 = b * TB + lane * KS;
    R zs0[KS], zs1[KS], zp[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const uint4 x = Philox::gen(seed, c0, c1, (uint32_t)(t0 + k), 1u);
      R unused;
      box_muller<R>(x.x, x.y, zs0[k], zs1[k]);
      box_muller<R>(x.z, x.w, zp[k], unused);
    }
    R j00[KS], j01[KS], j10[KS], j11[KS], o0[KS], o1[KS];
    llt_forward_sampling(B, s_e, q1, q2, pr, ld2, st, b == 0, lane, t0, zs0, zs1, j00, j01, j10, j11,
                         o0, o1);
    // lane composite of x_t = J_t x_{t+1} + o_t (reverse), stored as G = J' so that the adjoint's
    // combine rule (c = G_e' c_l + c_e) applies unchanged
    Rv2<R> L; L.g00 = 1; L.g01 = 0; L.g10 = 0; L.g11 = 1; L.c0 = 0; L.c1 = 0; L.c2 = 0;
#pragma unroll
    for (int k = KS - 1; k >= 0; --k) {
      Rv2<R> e; e.g00 = j00[k]; e.g01 = j10[k]; e.g10 = j01[k]; e.g11 = j11[k];
      e.c0 = o0[k]; e.c1 = o1[k]; e.c2 = 0;
      L = rv2_combine<R, false>(e, L);
    }
    Rv2<R> Sc = L;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const Rv2<R> O = rv2_shfl_down(Sc, off);
      if (lane + off < 32) Sc = rv2_combine<R, false>(Sc, O);
    }
    Rv2<R> X = rv2_shfl_down(Sc, 1);
    if (lane == 31) { X.g00 = 1; X.g01 = 0; X.g10 = 0; X.g11 = 1; X.c0 = 0; X.c1 = 0; X.c2 = 0; }
    R x0 = X.g00 * xc0 + X.g10 * xc1 + X.c0;       // state at the step after this lane's last one
    R x1 = X.g01 * xc0 + X.g11 * xc1 + X.c1;
    R lv[KS], sl[KS], tr[KS];
#pragma unroll
    for (int k = KS - 1; k >= 0; --k) {
      const R n0 = j00[k] * x0 + j01[k] * x1 + o0[k];
      const R n1 = j10[k] * x0 + j11[k] * x1 + o1[k];
      x0 = n0; x1 = n1;
      lv[k] = x0; sl[k] = x1;
      tr[k] = x0 + xw[k] + sig_e * zp[k];
    }
    xc0 = __shfl_sync(FULL, x0, 0); xc1 = __shfl_sync(FULL, x1, 0);
    const size_t row = (size_t)s * T;
    if (level) store_run(level + row, t0, T, lv);
    if (slope) store_run(slope + row, t0, T, sl);
    store_run(traj + row, t0, T, tr);
    pipe.release(lane);
  }
}

}  // namespace ci
