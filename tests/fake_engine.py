"""Oracle-backed stand-in for causalimpact_b200.Engine -- TESTS ONLY.

Lets the host logic (chain sharding, the single all-gather, draw ordering,
post-processing) run on CPU under gloo.  It answers the same methods as
Engine with the float64 oracle, keyed by the same global chain / draw ids."""
import numpy as np

from oracle import c_port
from oracle import hmc_np as H
from oracle import impact_np
from oracle import kalman_np as K
from oracle import quantiles_np
from oracle import smoother_np as SM


class FakeEngine:
  def __init__(self, device=0):
    self.device, self.spec, self.prob, self.seasonal = device, None, None, None

  def set_data(self, spec):
    self.spec = spec
    self.prob = K.Problem(
        model=spec.model, y=np.asarray(spec.y, float),
        X=None if spec.X is None else np.asarray(spec.X, float),
        Omega=None if spec.Omega is None else np.asarray(spec.Omega, float),
        m0=spec.m0, P0=spec.P0, obs_conc=spec.obs_conc, obs_scale=spec.obs_scale,
        obs_ub=spec.obs_ub, lvl_conc=spec.lvl_conc, lvl_scale=spec.lvl_scale, lvl_ub=spec.lvl_ub,
        ub_on_scale=bool(getattr(spec, "ub_on_scale", False)))

  def hmc_run(self, theta0, *, n_warmup, n_results, seed, chain_id0=0, max_leapfrog=8,
              init_step=0.05, target_accept=0.8, adapt_mass=True):
    f = lambda th: c_port.logpost_grad(self.prob, th)[:2]
    draws, st = H.run(f, theta0, n_warmup=n_warmup, n_results=n_results, seed=seed,
                      chain_id0=chain_id0, max_leapfrog=max_leapfrog, init_step=init_step,
                      target_accept=target_accept, adapt_mass=adapt_mass)
    stats = {k: np.asarray(st[k]) for k in ("accept_rate", "step_size", "n_divergent",
                                            "n_leapfrog")}
    return draws.astype(self.spec.np_dtype), stats

  def gibbs_run(self, n_chains, *, n_warmup, n_results, seed, chain_id0=0, sparse=True,
                nonzero_prob=None, want_level=True, want_traj=True, ssvs_order="random"):
    from oracle import gibbs_np as G
    sp, dt = self.spec, self.spec.np_dtype
    draws = np.empty((n_results, n_chains, sp.dim), dt)
    level = np.empty((n_results, n_chains, sp.T), dt)
    traj = np.empty((n_results, n_chains, sp.T), dt)
    incl = np.zeros((n_chains, max(sp.p, 1)), np.float32)
    for c in range(n_chains):
      g = G.run(self.prob, n_results=n_results, n_warmup=n_warmup,
                seed=(int(seed) % (2 ** 31)) * 1000003 + chain_id0 + c, sparse=sparse)
      draws[:, c, :sp.p] = g["w"]
      draws[:, c, sp.p] = np.log(g["s_e"]); draws[:, c, sp.p + 1] = np.log(g["s_h"])
      level[:, c] = g["level"]
      rng = np.random.default_rng(chain_id0 + c)
      loc = g["level"] + (g["w"] @ self.prob.X.T if sp.p else 0.0)
      traj[:, c] = loc + np.sqrt(g["s_e"])[:, None] * rng.normal(size=loc.shape)
      if sp.p:
        incl[c, :sp.p] = (g["w"] != 0).mean(0)
    return draws, level, traj, incl[:, :sp.p]

  def posterior_predict(self, theta_draws, *, seed, draw_id0=0, want_level=True):
    l, t, m = SM.posterior_predict(self.prob, theta_draws, seed, draw_id0)
    dt = self.spec.np_dtype
    return l.astype(dt), t.astype(dt), m.astype(dt)

  def row_quantiles(self, a, q):
    return quantiles_np.row_quantiles(a, q)

  # ---- the tensor-in / tensor-out surface api._train_causalimpact_sts drives (CPU tensors) ----
  def hmc_run_t(self, theta0, **kw):
    import torch
    draws, stats = self.hmc_run(np.asarray(theta0), **kw)
    rec = np.zeros(draws.shape[1], dtype=[("accept_rate", "f4"), ("step_size", "f4"),
                                          ("n_divergent", "i4"), ("n_leapfrog", "i4")])
    for k in rec.dtype.names:
      rec[k] = stats[k]
    return torch.from_numpy(np.ascontiguousarray(draws)), rec

  def gibbs_run_t(self, n_chains, **kw):
    import torch
    draws, level, traj, incl = self.gibbs_run(n_chains, **kw)
    cm = lambda a: torch.from_numpy(np.ascontiguousarray(a.transpose(1, 0, 2)).reshape(
        a.shape[0] * a.shape[1], a.shape[2]))
    return cm(draws), cm(level), cm(traj), incl

  def posterior_predict_t(self, theta_draws, *, seed, draw_id0=0):
    import torch
    l, t, _ = self.posterior_predict(np.asarray(theta_draws), seed=seed, draw_id0=draw_id0)
    return torch.from_numpy(l), torch.from_numpy(t)

  def predictive_mean_t(self, theta_draws, level):
    import torch
    th, lv = np.asarray(theta_draws, np.float64), np.asarray(level, np.float64)
    m = lv.mean(axis=0)
    if self.spec.p:
      m = m + self.prob.X @ th[:, :self.spec.p].mean(axis=0)
    return torch.from_numpy(m.astype(self.spec.np_dtype))

  # ---- seasonal components: the restated seasonal sweep (oracle/seasonal_np.py) ----
  def set_seasonal(self, sched):
    self.seasonal = sched

  def gibbs_seasonal_run_t(self, n_chains, *, n_warmup, n_results, seed, chain_id0=0, sparse=True,
                           nonzero_prob=None, ssvs_order="random"):
    import torch
    from oracle import seasonal_np as S
    sp, dt, sc = self.spec, self.spec.np_dtype, self.seasonal
    ssp = S.SeasonalSpec(n=list(sc.num_seasons), idx=np.asarray(sc.active, np.int64),
                         ends=np.asarray(sc.ends, bool), init_sd=sc.init_sd,
                         drift_conc=sc.drift_conc, drift_scale=sc.drift_scale, drift_ub=sc.drift_ub)
    K, R = ssp.K, n_chains * n_results
    th = np.empty((R, sp.dim), dt); lv = np.empty((R, sp.T), dt); la = np.empty((R, sp.T), dt)
    tr = np.empty((R, sp.T), dt); se = np.empty((R, sp.T, K), dt); dr = np.empty((R, K), dt)
    incl = np.zeros((n_chains, max(sp.p, 1)), np.float32)
    for c in range(n_chains):
      g = S.run(self.prob, ssp, n_results=n_results, n_warmup=n_warmup,
                seed=(int(seed) % (2 ** 31)) * 1000003 + chain_id0 + c, sparse=sparse and sp.p > 3)
      rows = slice(c * n_results, (c + 1) * n_results)
      th[rows, :sp.p] = g["w"]; th[rows, sp.p] = np.log(g["s_e"]); th[rows, sp.p + 1] = np.log(g["s_h"])
      lv[rows] = g["level"]; se[rows] = g["seasonal"]; dr[rows] = np.log(g["s_d"])
      la[rows] = g["level"] + g["seasonal"].sum(-1)
      rng = np.random.default_rng(chain_id0 + c)
      loc = la[rows] + (g["w"] @ self.prob.X.T if sp.p else 0.0)
      tr[rows] = loc + np.sqrt(g["s_e"])[:, None] * rng.normal(size=loc.shape)
      if sp.p:
        incl[c, :sp.p] = (g["w"] != 0).mean(0)
    t = torch.from_numpy
    return t(th), t(lv), t(la), t(tr), t(se), t(dr), incl[:, :sp.p]

  # ---- batches of independent series (fit_causalimpact_many) ----
  def set_data_batch(self, specs):
    self.batch_specs = list(specs)
    self.set_data(self.batch_specs[0])

  def batch_select(self, i, spec=None):
    self.set_data(spec if spec is not None else self.batch_specs[i])

  def gibbs_run_batch_t(self, n_chains, series_stride=0, chain_id0=0, **kw):
    import torch
    outs = []
    for i in range(len(self.batch_specs)):
      self.batch_select(i)
      # series i uses the global chain ids chain_id0 + i * series_stride + c (ci_gibbs_opts)
      outs.append(self.gibbs_run_t(n_chains, chain_id0=chain_id0 + i * series_stride, **kw))
    self.batch_select(0)
    return (torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs]),
            torch.stack([o[2] for o in outs]), np.stack([o[3] for o in outs]))

  # ---- panel prep + batched mean / impact (ci_set_panel, ci_*_batch_d): the host restatement ----
  def set_panel(self, values, *, row0, n_pre, standardize=True, dtype=np.float32,
                prior_level_sd=0.01, ub_on_scale=False):
    import causalimpact_b200 as cib
    from causalimpact_b200 import panel
    values = np.asarray(values, dtype=np.float64)
    N, T, _ = values.shape
    prep = panel.prepare_panel(values, np.arange(T), (row0, row0 + n_pre - 1),
                               (row0 + n_pre, T - 1), standardize, dtype)
    specs = [cib.build_problem(prep["y_ext"][i], None if prep["design"] is None else prep["design"][i],
                               prior_level_sd=prior_level_sd, outcome_sd=float(prep["outcome_sd"][i]),
                               dtype=dtype, ub_on_scale=ub_on_scale) for i in range(N)]
    self.set_data_batch(specs)
    stats = np.zeros((N, 8))
    stats[:, 0], stats[:, 1], stats[:, 2] = prep["y_scale"], prep["y_offset"], prep["outcome_sd"]
    return stats

  def predictive_mean_batch_t(self, theta, level):
    import torch
    out = []
    for i in range(theta.shape[0]):
      self.batch_select(i)
      out.append(self.predictive_mean_t(theta[i], level[i]))
    self.batch_select(0)
    return torch.stack(out)

  def impact_batch_t(self, traj, mean, *, scale, offset, obs_sum, observed, period, q_lo, q_hi):
    import torch
    ser, summ = [], []
    for i in range(traj.shape[0]):
      s9, sm = impact_np.impact_arrays(np.asarray(traj[i]), np.asarray(mean[i]), observed[i], period,
                                       float(scale[i]), float(offset[i]), q_lo, q_hi,
                                       float(obs_sum[i]))
      ser.append(s9.reshape(-1)); summ.append(sm)
    return torch.from_numpy(np.stack(ser)), torch.from_numpy(np.stack(summ))

  # ---- the two halves of ci_impact_d for sharded draws (ci_impact_rows_d / ci_impact_cols_d) ----
  def impact_rows_t(self, traj, mean, meta, out=None):
    import torch
    raw = np.asarray(traj)
    S, T = raw.shape
    per, obs = np.asarray(meta.period), np.asarray(meta.observed, dtype=np.float64)
    t_c0 = int(np.argmax(per != 0)) if np.any(per != 0) else T
    x = raw.astype(np.float64) * meta.scale + meta.offset
    point = obs[None, :] - x
    cum = impact_np._nan_cumsum(np.where((per == 0)[None, :], 0.0, point), axis=1)
    in_post = per == 1
    pred_sum = x[:, in_post].sum(axis=1)
    with np.errstate(invalid="ignore"), __import__("warnings").catch_warnings():
      __import__("warnings").simplefilter("ignore")
      eff_mean = np.nanmean(point[:, in_post], axis=1)
    stats = np.stack([x[:, in_post].mean(axis=1), pred_sum, eff_mean,
                      np.nansum(point[:, in_post], axis=1), meta.obs_sum / pred_sum - 1.0])
    packed = torch.from_numpy(np.ascontiguousarray(np.concatenate([stats, cum[:, t_c0:].T])))
    if mean is not None:
      m = np.asarray(mean, dtype=np.float64).reshape(-1) * meta.scale + meta.offset
      o = out.numpy()
      ser = o[:T * 9].reshape(T, 9)
      ser[:, 0], ser[:, 3] = m, obs - m
      ser[:, 6] = impact_np._nan_cumsum(np.where(per == 0, 0.0, obs - m), axis=0)
      o[T * 9 + 18], o[T * 9 + 19] = m[in_post].mean(), m[in_post].sum()
    return torch.from_numpy(np.ascontiguousarray(raw.T)), packed[5:], packed[:5], packed

  def impact_cols_t(self, trT, t_begin, cumT, c_begin, stats, meta, out):
    from oracle import quantiles_np
    per, obs = np.asarray(meta.period), np.asarray(meta.observed, dtype=np.float64)
    T = obs.shape[0]
    t_c0 = int(np.argmax(per != 0)) if np.any(per != 0) else T
    q = np.array([meta.q_lo, meta.q_hi])
    o = out.numpy()
    ser, summ = o[:T * 9].reshape(T, 9), o[T * 9:]
    nt, nc = trT.shape[0], cumT.shape[0]
    if nt:
      x = np.asarray(trT, dtype=np.float64).T * meta.scale + meta.offset       # [S, nt]
      ser[t_begin:t_begin + nt, 1:3] = quantiles_np.row_quantiles(x, q)
      ser[t_begin:t_begin + nt, 4:6] = quantiles_np.row_quantiles(obs[None, t_begin:t_begin + nt] - x, q)
      early = np.arange(t_begin, t_begin + nt) < t_c0
      ser[t_begin:t_begin + nt][early, 7:9] = 0.0
    if nc:
      ser[t_c0 + c_begin:t_c0 + c_begin + nc, 7:9] = quantiles_np.row_quantiles(np.asarray(cumT).T, q)
    if stats is not None:
      st = np.asarray(stats)                                                   # [5, S]
      S = st.shape[1]
      summ[0:10] = quantiles_np.row_quantiles(st.T, q).reshape(-1)
      summ[10:15] = st.std(axis=1, ddof=1) if S > 1 else np.nan
      summ[15] = st[4].mean()
      summ[16], summ[17] = np.sum(meta.obs_sum <= st[1]), np.sum(meta.obs_sum >= st[1])

  def to_host(self, t):
    return t.detach().cpu().numpy()

  def impact(self, traj, mean, meta, out=None):
    s9, summ = impact_np.impact_arrays(np.asarray(traj), np.asarray(mean), meta.observed,
                                       meta.period, meta.scale, meta.offset, meta.q_lo, meta.q_hi,
                                       meta.obs_sum)
    if out is None:
      return s9, summ
    import torch
    out.copy_(torch.from_numpy(np.concatenate([s9.reshape(-1), summ])))
    return None
