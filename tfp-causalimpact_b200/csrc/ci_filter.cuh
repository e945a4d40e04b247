// ci_filter.cuh -- warp-cooperative, time-parallel Kalman filter for the local
// level model: value, adjoint gradient, and the pieces the simulation smoother
// reuses.  ONE WARP PER CHAIN; within a tile of TB = 256 steps lane L owns the
// KS = 8 consecutive steps 8L..8L+7.
//
// Replaces (reference, relative to /root/reference): the LGSSM arithmetic TFP
// performs inside gibbs_sampler.fit_with_gibbs_sampling, call site
// causalimpact/causalimpact_lib.py:365-388.  Conventions follow TFP's
// LinearGaussianStateSpaceModel: N(m0,P0) is the prior of the state AT t=0,
// every step updates with y_t (unless masked) and then predicts.
//
// Time-parallel formulation (restated in oracle/scan_np.py):
//   variance  P' = (aP+b)/(cP+d)   -> 2x2 Moebius matrices, warp scan
//   mean      a' = (1-K)a + K r     -> affine maps, warp scan
//   adjoints  abar, Pbar            -> affine maps, reverse warp scans
// State that crosses tiles (a, P, abar, Pbar) is carried in registers; the
// forward pass checkpoints (a,P) per tile and the backward pass recomputes the
// tile from its checkpoint, so [X|y] is streamed exactly twice per
// value+gradient evaluation and nothing per-step is ever written to HBM.
#pragma once
#include "ci_common.cuh"

namespace ci {

template <typename R> struct Mob { R a, b, c, d; };

template <typename R> __device__ __forceinline__ Mob<R> mob_shfl_up(const Mob<R>& m, int off) {
  Mob<R> o;
  o.a = __shfl_up_sync(FULL, m.a, off); o.b = __shfl_up_sync(FULL, m.b, off);
  o.c = __shfl_up_sync(FULL, m.c, off); o.d = __shfl_up_sync(FULL, m.d, off);
  return o;
}
// later * earlier, rescaled (a Moebius map is projective: any scale is the same map)
template <typename R> __device__ __forceinline__ Mob<R> mob_mul(const Mob<R>& L, const Mob<R>& E) {
  Mob<R> o;
  o.a = L.a * E.a + L.b * E.c; o.b = L.a * E.b + L.b * E.d;
  o.c = L.c * E.a + L.d * E.c; o.d = L.c * E.b + L.d * E.d;
  const R s = Num<R>::pow2_rescale(o.a + o.b + o.c + o.d);   // entries stay in [0, 2)
  o.a *= s; o.b *= s; o.c *= s; o.d *= s;
  return o;
}

// ---------------------------------------------------------------------------
// Branch-free Kogge-Stone scan steps: a lane without a partner combines with the
// identity instead of skipping the step, so there is no divergence region
// (BSSY/BSYNC) per level.
// ---------------------------------------------------------------------------
template <typename R> __device__ __forceinline__ void affine_scan_up(R& m, R& c, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    R mo = __shfl_up_sync(FULL, m, off), co = __shfl_up_sync(FULL, c, off);
    const bool ok = lane >= off;
    mo = ok ? mo : (R)1; co = ok ? co : (R)0;
    c = fma(m, co, c); m = m * mo;
  }
}
template <typename R> __device__ __forceinline__ void affine_scan_down(R& m, R& c, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    R mo = __shfl_down_sync(FULL, m, off), co = __shfl_down_sync(FULL, c, off);
    const bool ok = lane + off < 32;
    mo = ok ? mo : (R)1; co = ok ? co : (R)0;
    c = fma(m, co, c); m = m * mo;
  }
}
// float32: the combine is PREDICATED instead of selecting the identity (one setp + two predicated
// ops per level instead of a compare, two selects and two ops); still no divergence region.
template <> __device__ __forceinline__ void affine_scan_up<float>(float& m, float& c, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float mo = __shfl_up_sync(FULL, m, off), co = __shfl_up_sync(FULL, c, off);
    asm("{\n\t.reg .pred q;\n\tsetp.ge.s32 q, %2, %3;\n\t@q fma.rn.f32 %0, %1, %5, %0;\n\t"
        "@q mul.f32 %1, %1, %4;\n\t}"
        : "+f"(c), "+f"(m) : "r"(lane), "r"(off), "f"(mo), "f"(co));
  }
}
template <> __device__ __forceinline__ void affine_scan_down<float>(float& m, float& c, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float mo = __shfl_down_sync(FULL, m, off), co = __shfl_down_sync(FULL, c, off);
    asm("{\n\t.reg .pred q;\n\tsetp.lt.s32 q, %2, %3;\n\t@q fma.rn.f32 %0, %1, %5, %0;\n\t"
        "@q mul.f32 %1, %1, %4;\n\t}"
        : "+f"(c), "+f"(m) : "r"(lane), "r"(32 - off), "f"(mo), "f"(co));
  }
}
template <typename R> __device__ __forceinline__ void mob_scan_up(Mob<R>& M, int lane) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    Mob<R> O = mob_shfl_up(M, off);
    const bool ok = lane >= off;
    O.a = ok ? O.a : (R)1; O.b = ok ? O.b : (R)0; O.c = ok ? O.c : (R)0; O.d = ok ? O.d : (R)1;
    M = mob_mul(M, O);
  }
}

// ---------------------------------------------------------------------------
// consumer side of the tile pipeline
// ---------------------------------------------------------------------------
template <typename R> struct TilePipe {
  const R* stage0;
  uint64_t* full;
  uint64_t* empty;
  uint32_t stage_elems;
  uint32_t nstage;
  uint32_t it;        // tiles consumed so far (streaming mode)
  uint32_t cur;
  bool resident;      // all tiles fit: loaded once, never released

  __device__ __forceinline__ const R* acquire(int tile) {
    uint32_t par;
    if (resident) { cur = (uint32_t)tile; par = 0u; }
    else { cur = it % nstage; par = (it / nstage) & 1u; }
    mbar_wait(&full[cur], par);
    return stage0 + (size_t)cur * stage_elems;
  }
  __device__ __forceinline__ void release(int lane) {
    if (!resident) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[cur]);
    }
    ++it;
  }
};

// Source address of tile b.  Written as an explicit 64-bit mad.wide: with plain
// pointer arithmetic ptxas 12.9 (sm_100a) lowered the unrolled resident loop to
// `UMOV UR7, URZ ; ULEA UR6, UR6, UR10, 0x2` -- dropping the high 32 bits of the
// pointer (caught by compute-sanitizer, round 1 GPU run 4).
__device__ __forceinline__ const void* tile_src(const void* base, uint32_t b, uint32_t bytes) {
  unsigned long long a;
  asm volatile("mad.wide.u32 %0, %1, %2, %3;"
               : "=l"(a) : "r"(b), "r"(bytes), "l"((unsigned long long)base));
  return reinterpret_cast<const void*>(a);
}

// Producer (one elected thread): issues the bulk copies for `n_pass` sweeps.
// Sweep s walks tiles 0..NB-1 when dir[s] > 0 and NB-1..0 otherwise; the
// caller describes the schedule through the functor `sweep_dir(s)`.
template <typename R, typename DirFn>
__device__ void tile_producer(const R* gtiles, R* stage0, uint64_t* full, uint64_t* empty,
                              uint32_t stage_elems, uint32_t nstage, int NB, bool resident,
                              long long n_sweeps, DirFn sweep_dir) {
  const uint32_t bytes = stage_elems * (uint32_t)sizeof(R);
  if (resident) {
    for (int b = 0; b < NB; ++b) {
      mbar_expect_tx(&full[b], bytes);
      bulk_g2s(stage0 + (size_t)b * stage_elems, tile_src(gtiles, (uint32_t)b, bytes), bytes,
               &full[b]);
    }
    return;
  }
  uint32_t it = 0;
  for (long long s = 0; s < n_sweeps; ++s) {
    const bool fwd = sweep_dir(s);
    for (int i = 0; i < NB; ++i, ++it) {
      const int b = fwd ? i : NB - 1 - i;
      const uint32_t st = it % nstage;
      if (it >= nstage) mbar_wait(&empty[st], ((it / nstage) - 1u) & 1u);
      mbar_expect_tx(&full[st], bytes);
      bulk_g2s(stage0 + (size_t)st * stage_elems, tile_src(gtiles, (uint32_t)b, bytes), bytes,
               &full[st]);
    }
  }
}

// ---------------------------------------------------------------------------
// one tile's worth of per-lane filter state
// ---------------------------------------------------------------------------
template <typename R> struct Blk {
  R r[KS];    // residual y - x.w   (0 where masked)
  R P[KS];    // predicted variance at the step
  R K[KS];    // gain               (0 where masked)
  R rF[KS];   // 1/F                (0 where masked)
  R v[KS];    // innovation         (0 where masked)
  uint32_t obs;
};

// r_k = y_k - sum_j x_kj w_j for the lane's KS rows; conflict-free shared reads.
template <typename R>
__device__ __forceinline__ void blk_residuals(Blk<R>& B, const R* __restrict__ tile,
                                              const R* __restrict__ w_s, int p, int ld, int lane) {
  const R* row0 = tile + tile_off(lane * KS, ld);
  R acc[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) acc[k] = row0[k * ld + p];
#pragma unroll 4
  for (int j = 0; j < p; ++j) {
    const R wj = w_s[j];
#pragma unroll
    for (int k = 0; k < KS; ++k) acc[k] = fma(-row0[k * ld + j], wj, acc[k]);
  }
  uint32_t obs = 0;
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const bool o = (acc[k] == acc[k]);
    obs |= (o ? 1u : 0u) << k;
    B.r[k] = o ? acc[k] : (R)0;
  }
  B.obs = obs;
}

// Same, but also returns the regression term xw_k = x_k . w (needed where the
// step is masked and y is NaN): r_k = y_k - xw_k.
template <typename R>
__device__ __forceinline__ void blk_residuals_xw(Blk<R>& B, R (&xw)[KS],
                                                 const R* __restrict__ tile,
                                                 const R* __restrict__ w_s, int p, int ld,
                                                 int lane) {
  const R* row0 = tile + tile_off(lane * KS, ld);
#pragma unroll
  for (int k = 0; k < KS; ++k) xw[k] = 0;
#pragma unroll 4
  for (int j = 0; j < p; ++j) {
    const R wj = w_s[j];
#pragma unroll
    for (int k = 0; k < KS; ++k) xw[k] = fma(row0[k * ld + j], wj, xw[k]);
  }
  uint32_t obs = 0;
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const R y = row0[k * ld + p];
    const bool o = (y == y);
    obs |= (o ? 1u : 0u) << k;
    B.r[k] = o ? (y - xw[k]) : (R)0;
  }
  B.obs = obs;
}

// Forward over one tile.  (a_c, P_c) carry the predicted moments in and out.
// FILT = false: B.v[k] = innovation v_k (0 where masked).
// FILT = true : B.v[k] = FILTERED mean m_k = a_k + K_k v_k (the smoother's input).
template <typename R, bool FILT = false>
__device__ __forceinline__ void blk_forward(Blk<R>& B, R s_e, R s_h, R& a_c, R& P_c, int lane) {
  // ---- variance path: Moebius scan ----
  const R alpha = s_e + s_h, beta = s_e * s_h;
  Mob<R> M{(R)1, (R)0, (R)0, (R)1};
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    // observed: [[alpha, beta], [1, s_e]]   masked: [[1, s_h], [0, 1]]   (select, no branch)
    const bool o = (B.obs >> k) & 1u;
    const R e1 = o ? alpha : (R)1, e2 = o ? beta : s_h;
    const R f1 = o ? (R)1 : (R)0, f2 = o ? s_e : (R)1;
    Mob<R> N;
    N.a = fma(e1, M.a, e2 * M.c); N.b = fma(e1, M.b, e2 * M.d);
    N.c = fma(f1, M.a, f2 * M.c); N.d = fma(f1, M.b, f2 * M.d);
    M = N;
  }
  {
    const R s = Num<R>::rcp_fast(M.a + M.b + M.c + M.d);
    M.a *= s; M.b *= s; M.c *= s; M.d *= s;
  }
mob_scan_up(M, lane);
  Mob<R> E = mob_shfl_up(M, 1);
  if (lane == 0) { E.a = 1; E.b = 0; E.c = 0; E.d = 1; }
  R Pc = fma(E.a, P_c, E.b) * Num<R>::rcp(fma(E.c, P_c, E.d));
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    B.P[k] = Pc;
    const bool o = (B.obs >> k) & 1u;
    const R rF = o ? Num<R>::rcp(Pc + s_e) : (R)0;
    const R K = Pc * rF;
    B.rF[k] = rF; B.K[k] = K;
    Pc = fma(-K, Pc, Pc) + s_h;
  }
  P_c = __shfl_sync(FULL, Pc, 31);
  // ---- mean path: affine scan ----
  R m = 1, c = 0;
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const R omk = (R)1 - B.K[k];
    c = fma(omk, c, B.K[k] * B.r[k]);
    m = omk * m;
  }
affine_scan_up(m, c, lane);
  R me = __shfl_up_sync(FULL, m, 1), ce = __shfl_up_sync(FULL, c, 1);
  if (lane == 0) { me = 1; ce = 0; }
  R ac = fma(me, a_c, ce);
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const R v = ((B.obs >> k) & 1u) ? (B.r[k] - ac) : (R)0;
    ac = fma(B.K[k], v, ac);
    B.v[k] = FILT ? ac : v;
  }
  a_c = __shfl_sync(FULL, ac, 31);
}

// sum over the lane's observed steps of  log F + v^2/F.  The logs are taken of
// products of 4 innovation variances (2 logs per lane per tile instead of 8).
template <typename R> __device__ __forceinline__ R blk_loglik_terms(const Blk<R>& B, R s_e) {
  R s = 0;
  R prod[KS / 4];
#pragma unroll
  for (int h = 0; h < KS / 4; ++h) prod[h] = 1;
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const bool o = (B.obs >> k) & 1u;
    prod[k >> 2] *= o ? (B.P[k] + s_e) : (R)1;
    s = fma(B.v[k] * B.v[k], B.rF[k], s);
  }
#pragma unroll
  for (int h = 0; h < KS / 4; ++h) s += Num<R>::log(prod[h]);
  return s;
}

// Reverse (adjoint) sweep over one tile.  (ab_c, pb_c) carry d ll/d a, d ll/d P
// of the step just after the tile in, and of the tile's first step out.
// rbar[k] = d ll / d r_k.
template <typename R>
__device__ __forceinline__ void blk_backward(const Blk<R>& B, R s_e, R& ab_c, R& pb_c, int lane,
                                             R& ge, R& gh, R* rbar) {
  // ---- abar: abar_t = (1-K) abar_{t+1} + v/F ----
  R m = 1, c = 0;
#pragma unroll
  for (int k = KS - 1; k >= 0; --k) {
    const R omk = (R)1 - B.K[k];
    c = fma(omk, c, B.v[k] * B.rF[k]);
    m = omk * m;
  }
affine_scan_down(m, c, lane);
  R me = __shfl_down_sync(FULL, m, 1), ce = __shfl_down_sync(FULL, c, 1);
  if (lane == 31) { me = 1; ce = 0; }
  R ab = fma(me, ab_c, ce);
  R abn[KS];
#pragma unroll
  for (int k = KS - 1; k >= 0; --k) {
    abn[k] = ab;
    ab = fma((R)1 - B.K[k], ab, B.v[k] * B.rF[k]);
  }
  ab_c = __shfl_sync(FULL, ab, 0);
  // ---- Pbar: Pbar_t = (1-K)^2 Pbar_{t+1} + abar_{t+1} v s_e/F^2 + dF ----
  R q[KS], dF[KS];
  m = 1; c = 0;
#pragma unroll
  for (int k = KS - 1; k >= 0; --k) {
    const R omk = (R)1 - B.K[k];
    const R mult = omk * omk;
    const R rF = B.rF[k], v = B.v[k];
    dF[k] = (R)-0.5 * (rF - v * v * rF * rF);
    q[k] = fma(abn[k] * v * s_e, rF * rF, dF[k]);
    c = fma(mult, c, q[k]);
    m = mult * m;
  }
affine_scan_down(m, c, lane);
  me = __shfl_down_sync(FULL, m, 1); ce = __shfl_down_sync(FULL, c, 1);
  if (lane == 31) { me = 1; ce = 0; }
  R pb = fma(me, pb_c, ce);
  R lge = 0, lgh = 0;
#pragma unroll
  for (int k = KS - 1; k >= 0; --k) {
    const R K = B.K[k], rF = B.rF[k], v = B.v[k];
    const R omk = (R)1 - K;
    lgh += pb;
    lge += fma(K * K, pb, dF[k]) - abn[k] * v * B.P[k] * rF * rF;
    rbar[k] = fma(K, abn[k], -v * rF);
    pb = fma(omk * omk, pb, q[k]);
  }
  pb_c = __shfl_sync(FULL, pb, 0);
  ge += lge; gh += lgh;
}

// acc[s] (+)= sum_tl rb[tl] * x[tl][j],  j = jj + 32 s, over this lane's share
// of the tile's rows (part, part+nparts, ...).  Transposed access: lanes <-> j.
template <typename R, int JS>
__device__ __forceinline__ void blk_xt_rbar(const R* __restrict__ tile, const R* __restrict__ rb,
                                            int p, int ld, int jj, int part, int nparts,
                                            R (&acc)[JS]) {
  // rows tl = part + nparts*i; 4 independent partial sums per slot hide the FMA latency
  const int nrow = TB / nparts;          // nparts divides 32, TB is a multiple of 32
  R a0[JS], a1[JS], a2[JS], a3[JS];
#pragma unroll
  for (int s = 0; s < JS; ++s) { a0[s] = 0; a1[s] = 0; a2[s] = 0; a3[s] = 0; }
  for (int i = 0; i < nrow; i += 4) {
    const int t0 = part + nparts * i, t1 = t0 + nparts, t2 = t1 + nparts, t3 = t2 + nparts;
    const R r0 = rb[t0 + (t0 >> 5)], r1 = rb[t1 + (t1 >> 5)];
    const R r2 = rb[t2 + (t2 >> 5)], r3 = rb[t3 + (t3 >> 5)];
    const R* w0 = tile + tile_off(t0, ld);
    const R* w1 = tile + tile_off(t1, ld);
    const R* w2 = tile + tile_off(t2, ld);
    const R* w3 = tile + tile_off(t3, ld);
#pragma unroll
    for (int s = 0; s < JS; ++s) {
      const int j = jj + 32 * s;
      if (j < p) {
        a0[s] = fma(r0, w0[j], a0[s]); a1[s] = fma(r1, w1[j], a1[s]);
        a2[s] = fma(r2, w2[j], a2[s]); a3[s] = fma(r3, w3[j], a3[s]);
      }
    }
  }
#pragma unroll
  for (int s = 0; s < JS; ++s) acc[s] += (a0[s] + a1[s]) + (a2[s] + a3[s]);
}

// Small-p variant (p <= PSMALL): keep the lane <-> time mapping of the residual
// pass and accumulate one register per covariate; conflict-free, no rbar
// round-trip through shared memory, 8 p FMAs per lane per tile.
constexpr int PSMALL = 16;
template <typename R>
__device__ __forceinline__ void blk_xt_rbar_small(const R* __restrict__ tile, const R (&rbar)[KS],
                                                  int p, int ld, int lane, R (&accw)[PSMALL]) {
  const R* row0 = tile + tile_off(lane * KS, ld);
#pragma unroll
  for (int j = 0; j < PSMALL; ++j) {
    if (j < p) {
#pragma unroll
      for (int k = 0; k < KS; ++k) accw[j] = fma(rbar[k], row0[k * ld + j], accw[j]);
    }
  }
}

}  // namespace ci
