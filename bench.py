#!/usr/bin/env python
"""Benchmark of the CausalImpact hot path on B200 ("Measurement").

  python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one pass of the hot path over one batch: one Kalman log-prob +
gradient evaluation of every chain of BASELINE.json configs[1] (local level +
10 covariates, T=1000, 256 chains per GPU).  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200"), os.path.join(ROOT, "tests")):
  if _p not in sys.path:
    sys.path.insert(0, _p)

WORKLOAD = dict(T=1000, n_cov=10, chains=256)       # BASELINE.json configs[1]
METRIC = "kalman_logprob_grad_evals_per_sec"
UNIT = "evals/s"


def bytes_per_eval(T, p, d=1):
  """B_vg = 8 T (p+1) + 4 (2 (p+1+d) + 1), float32."""
  return 8 * T * (p + 1) + 4 * (2 * (p + 1 + d) + 1)


def make_inputs(cfg, seed=20242):
  from conftest import make_series, make_thetas
  y, X, _ = make_series(cfg["T"], cfg["n_cov"], seed)
  p = 0 if X is None else X.shape[1]
  th = make_thetas(p + 2, p, cfg["chains"], seed + 1)
  return y, X, th


def measured_peak_hbm():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    try:
      return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:   # pylint: disable=broad-except
      pass
  return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
  """SM clock / throttle reasons sampled DURING the timed region (NVML every 5 ms;
  nvidia-smi as a fallback)."""
  Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
       "clocks_event_reasons.sw_power_cap")

  def __init__(self, index):
    self.index, self.rows, self._stop, self._th = index, [], threading.Event(), None
    self._nvml = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self._nvml = pynvml
    except Exception:   # pylint: disable=broad-except
      self._nvml = None

  def _sample_nvml(self):
    n = self._nvml
    sm = n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)
    mx = n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)
    try:
      r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
    except Exception:   # pylint: disable=broad-except
      r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
    act = lambda bit: "Active" if (r & bit) else "Not Active"
    self.rows.append([str(sm), str(mx), act(n.nvmlClocksThrottleReasonHwSlowdown),
                      act(n.nvmlClocksThrottleReasonHwThermalSlowdown),
                      act(n.nvmlClocksThrottleReasonSwThermalSlowdown),
                      act(n.nvmlClocksThrottleReasonSwPowerCap)])

  def _run(self):
    while not self._stop.is_set():
      try:
        if self._nvml is not None:
          self._sample_nvml()
        else:
          out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                "--format=csv,noheader,nounits"], capture_output=True, text=True,
                               timeout=5).stdout.strip()
          if out:
            self.rows.append([c.strip() for c in out.split(",")])
      except Exception:   # pylint: disable=broad-except
        pass
      self._stop.wait(0.005 if self._nvml is not None else 0.1)

  def __enter__(self):
    self._th = threading.Thread(target=self._run, daemon=True)
    self._th.start()
    return self

  def __exit__(self, *a):
    self._stop.set()
    self._th.join(timeout=6)

  def summary(self):
    if not self.rows:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
    mx = max(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(r[2 + i] == "Active" for r in self.rows)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
            "samples": len(self.rows), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


_BEST_THREADS = None


def usable_cpus():
  """CPUs this process may actually use: affinity mask, capped by a cgroup quota."""
  n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
  try:
    quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
    if quota != "max":
      n = max(1, min(n, int(np.ceil(int(quota) / int(period)))))
  except Exception:   # pylint: disable=broad-except
    pass
  return n


def cpu_port_rate(cfg, seconds=10.0, nthreads=None):
  """Times oracle/_ref (C port of the oracle, float64, OpenMP over chains) on a
  bounded sample of the SAME workload; returns evals/s and a description.

  The thread count is CALIBRATED once (0.3 s each for n, n/2, n/4, ... of the
  usable CPUs, best kept): torchrun exports OMP_NUM_THREADS=1, and on a shared
  box 128 spinning threads ran 180x slower than 64 (run 12) -- the baseline must
  be the fastest the host can do, not an accident of the environment."""
  global _BEST_THREADS
  os.environ.setdefault("OMP_WAIT_POLICY", "passive")
  from oracle import c_port
  from oracle import kalman_np as K
  y, X, th = make_inputs(cfg)
  prob = K.default_problem(y, X)
  c_port.logpost_grad(prob, th[:8])                     # load + warm
  if nthreads is None:
    if _BEST_THREADS is None:
      best, n = (0.0, 1), usable_cpus()
      cand = []
      while n >= 1:
        cand.append(n); n //= 2
      for nt in cand:
        t0 = time.perf_counter(); k = 0
        while time.perf_counter() - t0 < 0.3:
          c_port.logpost_grad(prob, th, nthreads=nt); k += th.shape[0]
        rate = k / (time.perf_counter() - t0)
        if rate > best[0]:
          best = (rate, nt)
      _BEST_THREADS = best[1]
    nthreads = _BEST_THREADS
  t0 = time.perf_counter(); n = 0; used = 1
  while True:
    _, _, used = c_port.logpost_grad(prob, th, nthreads=nthreads)
    n += th.shape[0]
    dt = time.perf_counter() - t0
    if dt >= seconds:
      break
  return n / dt, used, (f"{n} value+grad evals of the workload ({n // th.shape[0]} passes, "
                        f"{dt:.1f} s; {used} threads chosen by calibration out of "
                        f"{usable_cpus()} usable CPUs)")


def run_reference(args):
  """--impl reference: the reference's CPU path for this metric.  The reference
  delegates to TensorFlow Probability, which is not installable here (no
  network, not in /opt/wheelhouse), so per the task's tier rules this arm times
  the oracle's C port of that algorithm on the host cores."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  cfg = dict(WORKLOAD)
  per_step = max(0.2, min(8.0, 90.0 / max(1, args.steps + args.warmup)))   # whole arm <= ~1.5 min
  rates = []
  for i in range(args.warmup + args.steps):
    r, used, sample = cpu_port_rate(cfg, seconds=per_step)
    if i >= args.warmup:
      rates.append(r)
  val = float(np.mean(rates))
  print(json.dumps({
      "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
      "steps": args.steps, "warmup": args.warmup,
      "ms_per_step": 1e3 * cfg["chains"] / val, "higher_is_better": True, "scaling": "weak",
      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
      "config": {"workload": "local-level + 10 covariates, T=1000, 256 chains (configs[1])",
                 "note": "TFP is not installable here; C port of the oracle, OpenMP over chains"},
      "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": "port",
                       "sample": sample},
      "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=200)
  ap.add_argument("--warmup", type=int, default=20)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--cpu-seconds", type=float, default=10.0)
  args = ap.parse_args()
  if args.impl == "reference":
    return run_reference(args)
  if args.warmup < 3:
    args.warmup = 3

  import torch
  import torch.distributed as dist
  import causalimpact_b200 as cib
  from causalimpact_b200 import _engine

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a B200: the CUDA path has no CPU fallback")
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  dev = torch.device("cuda", local)

  cfg = dict(WORKLOAD)
  # weak scaling: every rank evaluates its own 256 chains (global chain ids
  # rank*256 .. rank*256+255); the series is replicated.
  y, X, th_all = make_inputs(dict(cfg, chains=cfg["chains"] * world))
  C = cfg["chains"]
  th_np = np.ascontiguousarray(th_all[rank * C:(rank + 1) * C], dtype=np.float32)
  spec = cib.build_problem(y, X)
  eng = cib.Engine(local)
  eng.set_data(spec)
  dim, p = spec.dim, spec.p

  theta = torch.from_numpy(th_np).to(dev)
  value = torch.empty(C, dtype=torch.float32, device=dev)
  grad = torch.empty(C, dim, dtype=torch.float32, device=dev)
  flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MB > L2
  stream = torch.cuda.current_stream()

  def step():
    eng.logprob_grad_ptr(theta.data_ptr(), C, value.data_ptr(), grad.data_ptr(),
                         _engine.VARIANT_SCAN, _engine.WITH_PRIOR, stream.cuda_stream)

  def sync_all():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  for _ in range(args.warmup):
    flush.zero_(); step()
  sync_all()
  l0 = eng.launch_count

  # ---- device-resident throughput: per-step CUDA events, L2 flushed between ----
  starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
  stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
  with ClockSampler(local) as clk:
    sync_all()
    for i in range(args.steps):
      flush.zero_()
      starts[i].record(stream); step(); stops[i].record(stream)
    sync_all()
    # hot-L2 back-to-back (no flush) for reference
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
      step()
    e1.record(stream)
    sync_all()
  launches = eng.launch_count - l0
  ms = np.array([a.elapsed_time(b) for a, b in zip(starts, stops)])
  t_step = float(ms.sum())                      # ms over K steps, kernel only
  t_hot = float(e0.elapsed_time(e1))

  # ---- end to end through the host-pointer C ABI, pinned host buffers ----
  th_pin = torch.from_numpy(th_np).pin_memory()
  val_pin = torch.empty(C, dtype=torch.float32).pin_memory()
  grad_pin = torch.empty(C, dim, dtype=torch.float32).pin_memory()

  def step_e2e():
    eng.logprob_grad_ptr(th_pin.data_ptr(), C, val_pin.data_ptr(), grad_pin.data_ptr(),
                         _engine.VARIANT_SCAN, _engine.WITH_PRIOR, host=True)

  for _ in range(args.warmup):
    step_e2e()
  sync_all()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    step_e2e()
  torch.cuda.synchronize()
  t_e2e = (time.perf_counter() - t0) * 1e3
  # parity guard on the numbers just produced
  if not np.all(np.isfinite(val_pin.numpy())):
    raise SystemExit("non-finite log-prob in bench")

  # ---- secondary metrics (BASELINE.json names both): HMC leapfrog evals/s inside the
  # persistent kernel, and posterior draws/s (simulation smoother + predictive) ----
  hmc_kw = dict(n_warmup=40, n_results=20, seed=20242, max_leapfrog=8, init_step=0.02)
  th0 = th_np.astype(np.float64)
  eng.hmc_run(th0, chain_id0=rank * C, **dict(hmc_kw, n_warmup=5, n_results=2))     # warm
  sync_all()
  t0 = time.perf_counter()
  _, hstats = eng.hmc_run(th0, chain_id0=rank * C, **hmc_kw)
  t_hmc = (time.perf_counter() - t0) * 1e3
  hmc_evals = int(hstats["n_leapfrog"].sum())
  S_pred = 4096
  thp = torch.from_numpy(np.ascontiguousarray(np.tile(th_np, (S_pred // C + 1, 1))[:S_pred])).to(dev)
  lvl = torch.empty(S_pred, cfg["T"], dtype=torch.float32, device=dev)
  trj = torch.empty_like(lvl)
  mean_d = torch.empty(cfg["T"], dtype=torch.float32, device=dev)
  lib, ctx = eng._lib, eng._ctx                       # raw _d entry for device-resident timing
  def predict():
    rc = lib.ci_posterior_predict_d(ctx, thp.data_ptr(), S_pred, 7, rank * S_pred, lvl.data_ptr(),
                                    trj.data_ptr(), mean_d.data_ptr(), stream.cuda_stream)
    assert rc == 0, lib.ci_last_error()
  for _ in range(3):
    predict()
  sync_all()
  p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  p0.record(stream)
  for _ in range(10):
    predict()
  p1.record(stream)
  sync_all()
  t_pred = float(p0.elapsed_time(p1)) / 10.0

  # ---- impact series + summary on the device (ci_impact_d) on those draws ----
  import types
  Tn = cfg["T"]
  t_pre = int(0.7 * Tn)
  per = np.zeros(Tn, np.uint8); per[t_pre:] = 1
  obs = np.random.Generator(np.random.PCG64(5)).normal(size=Tn)
  meta = types.SimpleNamespace(observed=obs, period=per, scale=2.0, offset=100.0, q_lo=0.025,
                               q_hi=0.975, obs_sum=float(obs[t_pre:].sum()))
  out_d = torch.empty(Tn * 9 + 20, dtype=torch.float64, device=dev)
  iargs = _engine.CiImpactArgs(S=S_pred, T=Tn, dtype=0, reserved=0, scale=2.0, offset=100.0,
                               q_lo=0.025, q_hi=0.975, obs_sum=meta.obs_sum)
  import ctypes
  def impact_step():
    rc = lib.ci_impact_d(ctx, ctypes.byref(iargs), trj.data_ptr(), mean_d.data_ptr(),
                         obs.ctypes.data_as(ctypes.c_void_p), per.ctypes.data_as(ctypes.c_void_p),
                         out_d.data_ptr(), out_d.data_ptr() + 8 * Tn * 9, stream.cuda_stream)
    assert rc == 0, lib.ci_last_error()
  for _ in range(3):
    impact_step()
  sync_all()
  p0.record(stream)
  for _ in range(10):
    impact_step()
  p1.record(stream)
  sync_all()
  t_impact = float(p0.elapsed_time(p1)) / 10.0

  # ---- BASELINE configs[4] shape on this GPU: 10 000-draw posterior forecast, T = 2000, 10
  # covariates -- simulation smoother + predictive draws + mean + the whole impact stage,
  # device resident (the second headline quantity: posterior draws/s) ----
  from conftest import make_series as _mk
  y5, X5, _ = _mk(2000, 10, 20245)
  eng5 = cib.Engine(local)
  eng5.set_data(cib.build_problem(y5, X5))
  S5 = 10000
  th5 = torch.from_numpy(np.ascontiguousarray(np.tile(th_np, (S5 // C + 1, 1))[:S5])).to(dev)
  per5 = np.zeros(2000, np.uint8); per5[1400:] = 1
  obs5 = np.random.Generator(np.random.PCG64(6)).normal(size=2000)
  meta5 = types.SimpleNamespace(observed=obs5, period=per5, scale=2.0, offset=100.0, q_lo=0.025,
                                q_hi=0.975, obs_sum=float(obs5[1400:].sum()))
  def forecast():
    lv5, tr5 = eng5.posterior_predict_t(th5, seed=11, draw_id0=rank * S5)
    mu5 = eng5.predictive_mean_t(th5, lv5)
    return eng5.impact(tr5, mu5, meta5)          # ends with the D2H of series + summary
  for _ in range(3):
    forecast()
  sync_all()
  tq = time.perf_counter()
  for _ in range(5):
    forecast()
  sync_all()
  t_fc = (time.perf_counter() - tq) / 5 * 1e3
  eng5.close()

  # ---- batch of independent series: one launch, grid.y = series ----
  specs_b = []
  for sb in range(128):
    yb, Xb, _ = _mk(300, 2, 3000 + sb)
    specs_b.append(cib.build_problem(yb, Xb))
  engb = cib.Engine(local)
  engb.set_data_batch(specs_b)
  engb.gibbs_run_batch_t(8, n_warmup=2, n_results=2, seed=1)
  sync_all()
  tq = time.perf_counter()
  engb.gibbs_run_batch_t(8, n_warmup=100, n_results=50, seed=1)
  sync_all()
  t_batch = (time.perf_counter() - tq) * 1e3
  engb.close()

  # ---- the whole product call for a panel of series (fit_causalimpact_panel): vectorised data
  # prep + batched sampler + queued mean / impact + one read-back; every rank fits its own panel
  # (series are independent: no collective) ----
  rp = np.random.Generator(np.random.PCG64(77 + rank))
  Np, Tp = 128, 300
  xs_p = 100 + np.cumsum(rp.normal(size=(Np, Tp, 2)), axis=1) * 0.3
  y_p = xs_p[:, :, 0] + rp.normal(size=(Np, Tp)); y_p[:, 210:] += 3.0
  vals_p = np.concatenate([y_p[:, :, None], xs_p], axis=2)
  kw_p = dict(seed=1, inference_options=cib.InferenceOptions(num_results=400),
              engine_options=cib.EngineOptions(num_chains=8, device=local))
  # (fit_causalimpact_panel shards over an initialised group; here every rank times its OWN
  # full panel, so run it with the group hidden)
  _w = cib.shard.world
  cib.shard.world = lambda: (0, 1)
  try:
    cib.fit_causalimpact_panel(vals_p[:4], np.arange(Tp), (0, 209), (210, Tp - 1), **kw_p)
    sync_all()
    tq = time.perf_counter()
    cib.fit_causalimpact_panel(vals_p, np.arange(Tp), (0, 209), (210, Tp - 1), **kw_p)
    t_panel = (time.perf_counter() - tq) * 1e3
  finally:
    cib.shard.world = _w

  # ---- the reference's own sampler on the GPU: Gibbs sweeps/s (spike-and-slab, 256 chains)
  eng.gibbs_run(C, n_warmup=2, n_results=2, seed=1, chain_id0=rank * C, want_level=False,
                want_traj=False)
  sync_all()
  t0 = time.perf_counter()
  eng.gibbs_run(C, n_warmup=30, n_results=30, seed=1, chain_id0=rank * C, want_level=False,
                want_traj=False)
  t_gibbs = (time.perf_counter() - t0) * 1e3

  # ---- the product call itself: fit_causalimpact on the quickstart shape (configs[0]:
  # T=100, 1 covariate, defaults = 900 draws).  The reference's only published number is
  # 5.17 s wall for this call on an unspecified notebook CPU (docs/quickstart.ipynb:361-362).
  # (every rank calls it: with a process group the fit shards its chains over the ranks
  # and ends with one all-gather, so a rank-0-only call would wait forever)
  import pandas as pd
  rs = np.random.Generator(np.random.PCG64(20241))
  xq = 100 + np.cumsum(rs.normal(size=100)) * 0.3
  yq = 1.2 * xq + rs.normal(size=100)
  yq[71:] += 10
  dfq = pd.DataFrame({"y": yq, "x": xq})
  t_fit = None
  for _ in range(2):
    sync_all()
    tq = time.perf_counter()
    cib.fit_causalimpact(dfq, (0, 70), (71, 99), seed=1,
                         engine_options=cib.EngineOptions(device=local))
    t_fit = (time.perf_counter() - tq) * 1e3

  if world > 1:
    t = torch.tensor([t_step, t_e2e, t_hot, t_hmc, t_pred, t_gibbs], dtype=torch.float64,
                     device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_step, t_e2e, t_hot, t_hmc, t_pred, t_gibbs = (float(x) for x in t.tolist())
    he = torch.tensor([hmc_evals], dtype=torch.float64, device=dev)
    dist.all_reduce(he, op=dist.ReduceOp.SUM)
    hmc_evals = int(he.item())

  team = os.environ.get("CI_B200_TEAM", "1") != "0" and 2 <= -(-cfg["T"] // 256) <= 8
  if rank == 0:
    total = C * world * args.steps
    val = total / (t_step * 1e-3)
    B = bytes_per_eval(cfg["T"], p)
    peak, peak_src = measured_peak_hbm()
    kern_ms = float(ms.mean())
    achieved = C * B / (kern_ms * 1e-3) / 1e9
    cpu_val, cores, sample = cpu_port_rate(cfg, seconds=args.cpu_seconds)
    out = {
        "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_step / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "local-level + 10 covariates, T=1000, 256 chains per GPU "
                               "(BASELINE.json configs[1]); value+gradient, prior included",
                   "variant": ("associative scan, TEAM mode: warp-shuffle scan per 256-step tile, one "
                               "warp per tile (4 warps per chain), tile aggregates exchanged via smem")
                              if team else "associative scan, one warp per chain (CI_B200_TEAM=0)",
                   "l2": "flushed (256 MB memset) between timed steps",
                   "timing": "CUDA events per step on the launching stream"},
        "e2e": {"value": total / (t_e2e * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(th_np.nbytes),
                "d2h_bytes_per_step": int(C * 4 + C * dim * 4)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     # dram__bytes_read+write per launch, ncu --set full (profiles/r01_k_logpost_*_ncu.md, run 33)
                     "traffic": 140288 if team else 133888,
                     "kernel": "k_logpost_team<float>" if team else "k_logpost_scan<float>",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": C * B,
                     "kernel_ms": kern_ms},
        "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "clocks": clk.summary(),
        "extra": {"value_hot_l2": total / (t_hot * 1e-3),
                  "ms_per_step_hot_l2": t_hot / args.steps,
                  "hmc_leapfrog_evals_per_sec": hmc_evals / (t_hmc * 1e-3),
                  "hmc_run": {"chains_per_gpu": C, "iterations": 60, "wall_ms": t_hmc,
                              "note": "host-pointer ci_hmc_run incl. copies + sync"},
                  "gibbs_sweeps_per_sec": C * world * 60 / (t_gibbs * 1e-3),
                  "gibbs_run": {"chains_per_gpu": C, "sweeps": 60, "wall_ms": t_gibbs,
                                "note": "ci_gibbs_run, spike-and-slab (inclusion prob 3/11), "
                                        "one sweep = SSVS + FFBS + 2 InvGamma draws; the reference "
                                        "publishes <= 5.2 ms per sweep (1 chain, T=100)"},
                  "fit_causalimpact_quickstart_ms": t_fit,
                  "fit_note": "T=100, 1 covariate, 900 draws, 64 chains, 2nd call; reference "
                              "publishes 5170 ms for this call (other hardware, incl. tracing)",
                  "forecast_10000_draws_T2000_ms": t_fc,
                  "forecast_draws_per_sec": S5 * world / (t_fc * 1e-3),
                  "forecast_note": "BASELINE configs[4] shape per GPU: ci_posterior_predict_d + "
                                   "ci_predictive_mean_d + ci_impact_d for 10 000 draws, T=2000, 10 "
                                   "covariates, wall clock incl. the D2H of series + summary",
                  "batch_gibbs": {"series": 128, "T": 300, "chains_per_series": 8, "sweeps": 150,
                                  "wall_ms": t_batch,
                                  "sweeps_per_sec": 128 * 8 * 150 * world / (t_batch * 1e-3),
                                  "note": "ci_gibbs_run_batch_d: every chain of every series in "
                                          "one launch (rank 0's time)"},
                  "panel_fit": {"series": Np, "T": Tp, "covariates": 2, "chains_per_series": 8,
                                "draws": 400, "wall_ms": t_panel,
                                "series_per_sec": Np * world / (t_panel * 1e-3),
                                "note": "fit_causalimpact_panel: the whole call (data prep, sampler, "
                                        "predictive mean, impact, result arrays) for a panel of "
                                        "independent series; the reference takes ~5 s per series"},
                  "impact_ms": t_impact,
                  "impact_note": f"ci_impact_d on {S_pred} draws x T={cfg['T']}: effect paths, 3 per-time "
                                 "quantile families, post-period summary; device resident "
                                 "(rank 0's time)",
                  "posterior_draws_per_sec": S_pred * world / (t_pred * 1e-3),
                  "posterior_draws": {"draws_per_gpu": S_pred, "T": cfg["T"], "ms": t_pred,
                                      "note": "ci_posterior_predict_d: level + trajectory + mean, "
                                              "device resident, CUDA events"}},
    }
    print(json.dumps(out))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
