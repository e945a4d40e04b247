// ci_kernels.cuh -- __global__ entry points (sm_100a).
#pragma once
#include "ci_device.cuh"

namespace ci {

// Common CTA prologue: barriers, Omega -> smem.  Returns pointers.
template <typename R> struct CtaShared {
  R* stage0; uint64_t* full; uint64_t* empty; R* omega;
  uint64_t* omega_bar;     // completes when the slab precision has landed in shared memory
};

// Issued by the producer thread: Omega travels as one bulk copy on its own
// barrier, so no consumer thread ever waits for it at start-up.
template <typename R>
__device__ __forceinline__ void omega_fetch(const CtaShared<R>& cs, const ProbDev<R>& pr) {
  const uint32_t bytes = ((uint32_t)(pr.p * pr.p) * (uint32_t)sizeof(R) + 15u) & ~15u;
  if (bytes == 0) { mbar_arrive(cs.omega_bar); return; }
  mbar_expect_tx(cs.omega_bar, bytes);
  bulk_g2s(cs.omega, pr.omega, bytes, cs.omega_bar);
}
template <typename R> __device__ __forceinline__ void omega_wait(const CtaShared<R>& cs) {
  mbar_wait(cs.omega_bar, 0u);
}

template <typename R>
__device__ __forceinline__ CtaShared<R> cta_prologue(unsigned char* smem, const SmemCfg& cfg,
                                                     const ProbDev<R>& pr, int n_consumers) {
  CtaShared<R> cs;
  cs.stage0 = reinterpret_cast<R*>(smem);
  cs.full = reinterpret_cast<uint64_t*>(smem + cfg.off_full);
  cs.empty = reinterpret_cast<uint64_t*>(smem + cfg.off_empty);
  cs.omega = reinterpret_cast<R*>(smem + cfg.off_omega);
  cs.omega_bar = cs.empty + cfg.nstage;
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < cfg.nstage; ++s) {
      mbar_init(&cs.full[s], 1);
      mbar_init(&cs.empty[s], (uint32_t)n_consumers);
    }
    mbar_init(cs.omega_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  return cs;
}

template <typename R>
__device__ __forceinline__ TilePipe<R> make_pipe(const CtaShared<R>& cs, const SmemCfg& cfg) {
  TilePipe<R> pipe;
  pipe.stage0 = cs.stage0; pipe.full = cs.full; pipe.empty = cs.empty;
  pipe.stage_elems = cfg.stage_elems; pipe.nstage = cfg.nstage; pipe.it = 0; pipe.cur = 0;
  pipe.resident = cfg.resident != 0;
  return pipe;
}

// ---------------------------------------------------------------------------
// K1/K2/K3: batched Kalman log-prob (+ gradient), associative-scan variant.
// grid = ceil(C / G) CTAs, block = 32*(G+1): G consumer warps (one chain each)
// + 1 producer warp feeding [X|y] tiles through the mbarrier pipeline.
// ---------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(32 * (MAXG + 1), 1)
k_logpost_scan(ProbDev<R> pr, SmemCfg cfg, const R* __restrict__ theta, int C,
               R* __restrict__ value, R* __restrict__ grad, int flags) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (blockDim.x >> 5) - 1;
  const int chain0 = blockIdx.x * G;
  const int nactive = min(G, C - chain0);
  const CtaShared<R> cs = cta_prologue(smem, cfg, pr, nactive);
  const bool want_grad = grad != nullptr;

  if (warp == G) {
    if (lane == 0) {
      omega_fetch(cs, pr);
      tile_producer(pr.tiles, cs.stage0, cs.full, cs.empty, cfg.stage_elems, cfg.nstage, pr.NB,
                    cfg.resident != 0, want_grad ? 2LL : 1LL,
                    [](long long s) { return (s & 1) == 0; });
    }
    return;
  }
  if (warp >= nactive) return;

  const int c = chain0 + warp;
  const int p = pr.p, dim = pr.dim;
  const R* th = theta + (size_t)c * dim;
  const WarpScratch<R> ws = warp_scratch<R>(smem, cfg, warp);
  for (int j = lane; j < p; j += 32) ws.w[j] = th[j];
  const R u = th[p], l = th[p + 1];
  const R s_e = Num<R>::exp(u), s_h = Num<R>::exp(l);
  __syncwarp();

  TilePipe<R> pipe = make_pipe(cs, cfg);
  double ll, g_se, g_sh;
  R gw[JS];
  chain_eval(pipe, pr, ws, s_e, s_h, want_grad, lane, ll, g_se, g_sh, gw);

  double val = ll;
  double g_u = g_se * (double)s_e, g_l = g_sh * (double)s_h;
  if (flags & 1) {
    omega_wait(cs);
    val += chain_prior(pr, cs.omega, ws.w, u, l, s_e, s_h, lane, gw, g_u, g_l);
  }
  if (lane == 0) value[c] = (R)val;
  if (want_grad) {
    R* g = grad + (size_t)c * dim;
#pragma unroll
    for (int s = 0; s < JS; ++s) {
      const int j = lane + 32 * s;
      if (j < p) g[j] = gw[s];
    }
    if (lane == 0) { g[p] = (R)g_u; g[p + 1] = (R)g_l; }
  }
}

}  // namespace ci
