"""ctypes binding of the C ABI in include/ci_b200.h (libci_b200.so).

This is the only place Python touches the native library.  There is no CPU
fallback: if the shared library is missing or no B200 is visible the calls
raise ``EngineError`` -- loudly, never silently.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
from typing import Optional, Tuple

import numpy as np

from . import _build

F32, F64 = 0, 1
MODEL_LOCAL_LEVEL, MODEL_LOCAL_LINEAR_TREND = 0, 1
VARIANT_SEQ, VARIANT_SCAN = 0, 1
WITH_PRIOR = 1


class EngineError(RuntimeError):
  """A C-ABI call failed (message comes from ci_last_error())."""


class CiProblem(C.Structure):
  _fields_ = [("model", C.c_int32), ("dtype", C.c_int32), ("T", C.c_int32), ("p", C.c_int32),
              ("obs_conc", C.c_double), ("obs_scale", C.c_double), ("obs_ub", C.c_double),
              ("lvl_conc", C.c_double), ("lvl_scale", C.c_double), ("lvl_ub", C.c_double),
              ("slope_conc", C.c_double), ("slope_scale", C.c_double), ("slope_ub", C.c_double),
              ("m0", C.c_double), ("P0", C.c_double),
              ("m0_slope", C.c_double), ("P0_slope", C.c_double),
              ("ub_on_scale", C.c_int32), ("reserved", C.c_int32)]


class CiHmcOpts(C.Structure):
  _fields_ = [("n_warmup", C.c_int32), ("n_results", C.c_int32), ("max_leapfrog", C.c_int32),
              ("adapt_mass", C.c_int32), ("init_step", C.c_double), ("target_accept", C.c_double)]


class CiGibbsOpts(C.Structure):
  _fields_ = [("n_warmup", C.c_int32), ("n_results", C.c_int32), ("sparse", C.c_int32),
              ("chain_major", C.c_int32), ("nonzero_prob", C.c_double),
              ("ssvs_order", C.c_int32), ("reserved", C.c_int32), ("series_stride", C.c_uint64)]

SSVS_ORDER = {"random": 0, "index": 1}


class CiImpactArgs(C.Structure):
  _fields_ = [("S", C.c_int32), ("T", C.c_int32), ("dtype", C.c_int32), ("reserved", C.c_int32),
              ("scale", C.c_double), ("offset", C.c_double), ("q_lo", C.c_double),
              ("q_hi", C.c_double), ("obs_sum", C.c_double)]


class CiPanelArgs(C.Structure):
  _fields_ = [("n_series", C.c_int32), ("T_total", C.c_int32), ("n_cols", C.c_int32),
              ("row0", C.c_int32), ("n_pre", C.c_int32), ("standardize", C.c_int32),
              ("dtype", C.c_int32), ("ub_on_scale", C.c_int32), ("prior_level_sd", C.c_double)]


PANEL_STATS = 8
IMPACT_SERIES_COLS, IMPACT_SUMMARY_LEN = 9, 20
MAX_SEASONAL = 7


class CiSeasonal(C.Structure):
  _fields_ = [("n_components", C.c_int32), ("num_seasons", C.c_int32 * MAX_SEASONAL),
              ("active", C.c_void_p), ("ends", C.c_void_p), ("init_sd", C.c_double),
              ("drift_conc", C.c_double), ("drift_scale", C.c_double), ("drift_ub", C.c_double)]


class CiHmcStats(C.Structure):
  _fields_ = [("accept_rate", C.c_float), ("step_size", C.c_float),
              ("n_divergent", C.c_int32), ("n_leapfrog", C.c_int32)]


HMC_STATS_DTYPE = np.dtype([("accept_rate", np.float32), ("step_size", np.float32),
                            ("n_divergent", np.int32), ("n_leapfrog", np.int32)])

# every symbol include/ci_b200.h declares
EXPORTS = (
    "ci_version", "ci_last_error", "ci_device_count", "ci_ctx_create", "ci_ctx_destroy",
    "ci_launch_count", "ci_set_data", "ci_logprob", "ci_logprob_grad", "ci_logprob_grad_d",
    "ci_hmc_run", "ci_hmc_run_d", "ci_gibbs_run", "ci_gibbs_run_d", "ci_posterior_predict", "ci_posterior_predict_d",
    "ci_row_quantiles", "ci_row_quantiles_d", "ci_predictive_mean_d", "ci_impact", "ci_impact_d",
    "ci_set_seasonal", "ci_gibbs_seasonal_run", "ci_gibbs_seasonal_run_d",
    "ci_set_data_batch", "ci_batch_select", "ci_gibbs_run_batch_d", "ci_set_seasonal_batch",
    "ci_gibbs_seasonal_run_batch_d",
    "ci_comm_get_unique_id", "ci_comm_create", "ci_allgather", "ci_comm_destroy",
    "ci_set_panel", "ci_predictive_mean_batch_d", "ci_impact_batch_d",
    "ci_impact_rows_d", "ci_impact_cols_d", "ci_impact_sharded_d",
)

_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
  """dlopen the in-tree libci_b200.so (built by __graft_entry__.build())."""
  global _lib
  if _lib is not None and path is None:
    return _lib
  path = path or os.environ.get("CI_B200_LIB") or _build.LIB_PATH
  if not os.path.exists(path):
    raise EngineError(
        f"{path} not found: the CUDA extension is not built. Run "
        "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
        "There is no CPU fallback.")
  lib = C.CDLL(path)
  vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
  lib.ci_version.restype = i32
  lib.ci_last_error.restype = C.c_char_p
  lib.ci_device_count.restype = i32
  lib.ci_ctx_create.argtypes = [i32, C.POINTER(vp)]
  lib.ci_ctx_destroy.argtypes = [vp]
  lib.ci_launch_count.argtypes = [vp]
  lib.ci_launch_count.restype = C.c_int64
  lib.ci_set_data.argtypes = [vp, C.POINTER(CiProblem), vp, vp, vp]
  lib.ci_logprob.argtypes = [vp, vp, i32, vp, i32, i32]
  lib.ci_logprob_grad.argtypes = [vp, vp, i32, vp, vp, i32, i32]
  lib.ci_logprob_grad_d.argtypes = [vp, vp, i32, vp, vp, i32, i32, vp]
  lib.ci_hmc_run.argtypes = [vp, C.POINTER(CiHmcOpts), u64, u64, vp, i32, vp, vp]
  lib.ci_hmc_run_d.argtypes = [vp, C.POINTER(CiHmcOpts), u64, u64, vp, i32, vp, vp, vp]
  lib.ci_gibbs_run.argtypes = [vp, C.POINTER(CiGibbsOpts), u64, u64, i32, vp, vp, vp, vp]
  lib.ci_gibbs_run_d.argtypes = [vp, C.POINTER(CiGibbsOpts), u64, u64, i32, vp, vp, vp, vp, vp]
  lib.ci_posterior_predict.argtypes = [vp, vp, i32, u64, u64, vp, vp, vp]
  lib.ci_posterior_predict_d.argtypes = [vp, vp, i32, u64, u64, vp, vp, vp, vp]
  lib.ci_row_quantiles.argtypes = [vp, vp, i32, i32, i32, C.POINTER(C.c_double), i32, vp]
  lib.ci_row_quantiles_d.argtypes = [vp, vp, i32, i32, i32, C.POINTER(C.c_double), i32, vp, vp]
  lib.ci_predictive_mean_d.argtypes = [vp, vp, vp, i32, vp, vp]
  lib.ci_impact.argtypes = [vp, C.POINTER(CiImpactArgs), vp, vp, vp, vp, vp, vp]
  lib.ci_impact_d.argtypes = [vp, C.POINTER(CiImpactArgs), vp, vp, vp, vp, vp, vp, vp]
  lib.ci_set_seasonal.argtypes = [vp, C.POINTER(CiSeasonal)]
  lib.ci_gibbs_seasonal_run.argtypes = [vp, C.POINTER(CiGibbsOpts), u64, u64, i32, vp, vp, vp, vp,
                                        vp, vp, vp]
  lib.ci_gibbs_seasonal_run_d.argtypes = [vp, C.POINTER(CiGibbsOpts), u64, u64, i32, vp, vp, vp,
                                          vp, vp, vp, vp, vp]
  lib.ci_set_seasonal_batch.argtypes = [vp, C.POINTER(CiSeasonal), vp, vp, vp]
  lib.ci_gibbs_seasonal_run_batch_d.argtypes = [vp, C.POINTER(CiGibbsOpts), u64, u64, i32, vp, vp,
                                                vp, vp, vp, vp, vp, vp]
  lib.ci_set_data_batch.argtypes = [vp, C.POINTER(CiProblem), i32, vp, vp, vp]
  lib.ci_batch_select.argtypes = [vp, i32]
  lib.ci_gibbs_run_batch_d.argtypes = [vp, C.POINTER(CiGibbsOpts), u64, u64, i32, vp, vp, vp, vp, vp]
  lib.ci_set_panel.argtypes = [vp, C.POINTER(CiPanelArgs), vp, vp]
  lib.ci_predictive_mean_batch_d.argtypes = [vp, vp, vp, i32, vp, vp]
  lib.ci_impact_batch_d.argtypes = [vp, C.POINTER(CiImpactArgs), i32, vp, vp, vp, vp, vp, vp, vp, vp,
                                    vp, vp]
  lib.ci_impact_rows_d.argtypes = [vp, C.POINTER(CiImpactArgs), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
  lib.ci_impact_cols_d.argtypes = [vp, C.POINTER(CiImpactArgs), vp, i32, i32, vp, i32, i32, vp, vp,
                                   vp, vp, vp, vp]
  lib.ci_impact_sharded_d.argtypes = [vp, vp, C.POINTER(CiImpactArgs), vp, vp, vp, vp, vp, vp, vp, vp]
  lib.ci_comm_get_unique_id.argtypes = [vp]
  lib.ci_comm_create.argtypes = [vp, vp, i32, i32, C.POINTER(vp)]
  lib.ci_allgather.argtypes = [vp, vp, vp, C.c_size_t, vp]
  lib.ci_comm_destroy.argtypes = [vp]
  _lib = lib
  return lib


@dataclasses.dataclass
class ProblemSpec:
  """Host-side twin of ``ci_problem`` plus the arrays ci_set_data uploads."""
  model: int
  dtype: int
  y: np.ndarray                       # [T], NaN == masked
  X: Optional[np.ndarray]             # [T, p] or None
  Omega: Optional[np.ndarray]         # [p, p] or None
  m0: float
  P0: float
  obs_conc: float
  obs_scale: float
  obs_ub: float
  lvl_conc: float
  lvl_scale: float
  lvl_ub: float
  slope_conc: float = 16.0
  slope_scale: float = 0.0
  slope_ub: float = float("inf")
  m0_slope: float = 0.0
  P0_slope: float = 1.0
  ub_on_scale: bool = False           # False: the *_ub bound VARIANCES (TFP), True: scales

  def ub_variance(self, ub: float) -> float:
    """The bound `ub` expressed on the variance."""
    return float(ub) ** 2 if self.ub_on_scale else float(ub)

  @property
  def T(self) -> int:
    return int(self.y.shape[0])

  @property
  def p(self) -> int:
    return 0 if self.X is None else int(self.X.shape[1])

  @property
  def d(self) -> int:
    return 2 if self.model == MODEL_LOCAL_LINEAR_TREND else 1

  @property
  def dim(self) -> int:
    return self.p + 1 + self.d

  @property
  def np_dtype(self):
    return np.float64 if self.dtype == F64 else np.float32


def _ptr(a: Optional[np.ndarray]):
  return None if a is None else a.ctypes.data_as(C.c_void_p)


class DeviceArray:
  """An array that lives in the engine's HBM (a torch CUDA tensor inside -- torch is
  only the allocator).  Quacks enough like the ndarray / tf.Tensor the reference
  returns (``.shape``, ``.numpy()``, ``np.asarray``) that callers which do want the
  values on the host get them, copied on demand; ``fit_causalimpact`` itself never
  asks: the trajectories go from the sampler kernels to ``ci_impact`` without leaving
  the device."""

  def __init__(self, tensor):
    self.tensor = tensor

  @property
  def shape(self):
    return tuple(self.tensor.shape)

  @property
  def dtype(self):
    return np.dtype(str(self.tensor.dtype).replace("torch.", ""))

  def __len__(self):
    return int(self.tensor.shape[0])

  def numpy(self) -> np.ndarray:
    return self.tensor.detach().cpu().numpy()

  def __array__(self, dtype=None, copy=None):
    a = self.numpy()
    return a if dtype is None else a.astype(dtype, copy=False)


COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
  """ci_comm_get_unique_id: the 128-byte NCCL id rank 0 hands to the other ranks."""
  lib = load_library()
  buf = C.create_string_buffer(COMM_ID_BYTES)
  rc = lib.ci_comm_get_unique_id(buf)
  if rc != 0:
    raise EngineError(f"ci_b200 error {rc}: {lib.ci_last_error().decode()}")
  return buf.raw


class Comm:
  """``ci_comm``: the path's one collective without torch.distributed (NCCL under the C ABI)."""

  def __init__(self, engine: "Engine", unique_id: bytes, rank: int, nranks: int):
    if len(unique_id) != COMM_ID_BYTES:
      raise ValueError(f"unique_id must be {COMM_ID_BYTES} bytes")
    self._eng, self.rank, self.nranks = engine, rank, nranks
    self._h = C.c_void_p()
    engine._check(engine._lib.ci_comm_create(engine._ctx, unique_id, rank, nranks, C.byref(self._h)))

  def allgather_rows(self, local):
    """[rows, width] device tensor of every rank (same shape) -> [nranks * rows, width]."""
    torch, dev = self._eng._torch_dev()
    local = local.contiguous()
    out = torch.empty((self.nranks * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                      device=dev)
    self._eng._check(self._eng._lib.ci_allgather(
        self._h, local.data_ptr(), out.data_ptr(), local.numel() * local.element_size(),
        self._eng._stream(torch)))
    return out

  def impact_sharded_t(self, traj_local, mean_part, meta, counts):
    """ci_impact_sharded_d: the impact stage over draws sharded ``counts[r]`` per rank, the
    exchange inside (one grouped send/receive + one all-reduce on the current stream).
    traj_local [counts[rank], T] and mean_part [T] (this rank's ci_predictive_mean_d) are device
    tensors.  Returns (out [T*9 + 20] float64, mean [T]) device tensors, the same on every rank;
    nothing is synchronised."""
    eng = self._eng
    torch, dev = eng._torch_dev()
    traj_local, mean_part = traj_local.contiguous(), mean_part.contiguous().reshape(-1)
    S, T = traj_local.shape
    cnt = np.ascontiguousarray(counts, dtype=np.int32)
    if cnt.shape != (self.nranks,) or int(cnt[self.rank]) != S:
      raise ValueError("counts must be [nranks] with counts[rank] == traj_local.shape[0]")
    obs = np.ascontiguousarray(meta.observed, dtype=np.float64)
    per = np.ascontiguousarray(meta.period, dtype=np.uint8)
    if obs.shape != (T,) or per.shape != (T,):
      raise ValueError(f"observed / period must be [{T}]")
    args = eng._impact_args(meta, S, T, traj_local.dtype)
    out = torch.empty(T * IMPACT_SERIES_COLS + IMPACT_SUMMARY_LEN, dtype=torch.float64, device=dev)
    mean = torch.empty(T, dtype=traj_local.dtype, device=dev)
    part = mean_part.to(traj_local.dtype)
    eng._check(eng._lib.ci_impact_sharded_d(
        eng._ctx, self._h, C.byref(args), _ptr(cnt), traj_local.data_ptr() if S else None,
        part.data_ptr(), _ptr(obs), _ptr(per), mean.data_ptr(),
        out.data_ptr(), eng._stream(torch)))
    return out, mean

  def close(self):
    if self._h:
      self._eng._lib.ci_comm_destroy(self._h)
      self._h = C.c_void_p()


class Engine:
  """One ``ci_ctx``: a device, its uploaded problem and workspaces."""

  def __init__(self, device: int = 0):
    self._lib = load_library()
    self._ctx = C.c_void_p()
    self.device = device
    self.spec: Optional[ProblemSpec] = None
    self.seasonal = None
    self.batch_specs = None
    self._check(self._lib.ci_ctx_create(device, C.byref(self._ctx)))

  # -- plumbing ------------------------------------------------------------
  def _check(self, rc: int):
    if rc != 0:
      msg = self._lib.ci_last_error()
      raise EngineError(f"ci_b200 error {rc}: {msg.decode() if msg else '?'}")

  def close(self):
    comm = getattr(self, "_shard_comm", None)      # shard.engine_comm: destroyed before its context
    if comm is not None:
      self._shard_comm = None
      comm.close()
    if self._ctx:
      self._lib.ci_ctx_destroy(self._ctx)
      self._ctx = C.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:   # pylint: disable=broad-except
      pass

  @property
  def launch_count(self) -> int:
    return int(self._lib.ci_launch_count(self._ctx))

  def _arr(self, a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=self.spec.np_dtype)
    if shape is not None and tuple(a.shape) != tuple(shape):
      raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a

  # -- problem -------------------------------------------------------------
  def set_data(self, spec: ProblemSpec):
    self.spec = spec
    self.seasonal = None
    self.batch_specs = None
    dt = spec.np_dtype
    y = np.ascontiguousarray(spec.y, dtype=dt)
    X = None if spec.X is None else np.ascontiguousarray(spec.X, dtype=dt)
    Om = None if spec.Omega is None else np.ascontiguousarray(spec.Omega, dtype=dt)
    pb = self._ci_problem(spec)
    self._check(self._lib.ci_set_data(self._ctx, C.byref(pb), _ptr(y), _ptr(X), _ptr(Om)))

  @staticmethod
  def _ci_problem(spec: ProblemSpec) -> CiProblem:
    return CiProblem(model=spec.model, dtype=spec.dtype, T=spec.T, p=spec.p,
                     obs_conc=spec.obs_conc, obs_scale=spec.obs_scale, obs_ub=spec.obs_ub,
                     lvl_conc=spec.lvl_conc, lvl_scale=spec.lvl_scale, lvl_ub=spec.lvl_ub,
                     slope_conc=spec.slope_conc, slope_scale=spec.slope_scale,
                     slope_ub=min(spec.slope_ub, 1e300), m0=spec.m0, P0=spec.P0,
                     m0_slope=spec.m0_slope, P0_slope=spec.P0_slope,
                     ub_on_scale=int(bool(spec.ub_on_scale)), reserved=0)

  # -- batches of independent series -----------------------------
  def set_data_batch(self, specs):
    """ci_set_data_batch: N series sharing (T, p, dtype); series 0 becomes current."""
    specs = list(specs)
    s0 = specs[0]
    dt = s0.np_dtype
    for sp in specs:
      if (sp.T, sp.p, sp.dtype, sp.model) != (s0.T, s0.p, s0.dtype, s0.model):
        raise ValueError("the series of a batch must share T, the number of covariates and dtype")
    probs = (CiProblem * len(specs))(*[self._ci_problem(sp) for sp in specs])
    y = np.ascontiguousarray(np.stack([sp.y for sp in specs]), dtype=dt)
    X = Om = None
    if s0.p:
      X = np.ascontiguousarray(np.stack([sp.X for sp in specs]), dtype=dt)
      Om = np.ascontiguousarray(np.stack([sp.Omega for sp in specs]), dtype=dt)
    self.spec, self.seasonal, self.batch_specs = s0, None, specs
    self._check(self._lib.ci_set_data_batch(self._ctx, probs, len(specs), _ptr(y), _ptr(X), _ptr(Om)))

  def set_panel(self, values, *, row0: int, n_pre: int, standardize: bool = True, dtype=np.float32,
                prior_level_sd: float = 0.01, ub_on_scale: bool = False) -> np.ndarray:
    """ci_set_panel: data prep of a whole panel ON THE DEVICE.  values [N, T_total, 1 + k]
    float64 (column 0 = outcome); rows [row0, row0 + n_pre) are the pre-period, rows from row0
    on the modelled span.  Returns stats [N, 8]: y_scale, y_offset, outcome_sd, n_obs, y'y, m0,
    error code, reserved.  Afterwards the engine holds the batch (as after set_data_batch)."""
    values = np.ascontiguousarray(values, dtype=np.float64)
    if values.ndim != 3:
      raise ValueError("values must be [n_series, T, 1 + n_covariates]")
    N, T_total, ncol = values.shape
    np_dt = np.dtype(dtype)
    args = CiPanelArgs(n_series=N, T_total=T_total, n_cols=ncol, row0=int(row0), n_pre=int(n_pre),
                       standardize=int(bool(standardize)), dtype=F64 if np_dt == np.float64 else F32,
                       ub_on_scale=int(bool(ub_on_scale)), prior_level_sd=float(prior_level_sd))
    stats = np.empty((N, PANEL_STATS), dtype=np.float64)
    rc = self._lib.ci_set_panel(self._ctx, C.byref(args), _ptr(values), _ptr(stats))
    if rc != 0:
      msg = self._lib.ci_last_error().decode()
      # the reference's own input errors are ValueErrors (data.py:140-190)
      if rc == -1 and ("Input " in msg or "observed" in msg):
        raise ValueError(msg)
      raise EngineError(f"ci_b200 error {rc}: {msg}")
    Tm, p = T_total - int(row0), (ncol if ncol > 1 else 0)
    # a shape-only spec: the batched entry points need T / p / dim / dtype, not the arrays
    self.spec = ProblemSpec(model=MODEL_LOCAL_LEVEL, dtype=args.dtype, y=np.empty(Tm),
                            X=np.empty((Tm, p)) if p else None, Omega=None, m0=0.0, P0=1.0,
                            obs_conc=0.0, obs_scale=0.0, obs_ub=0.0, lvl_conc=0.0, lvl_scale=0.0,
                            lvl_ub=0.0, ub_on_scale=bool(ub_on_scale))
    self.seasonal = None
    self.batch_specs = [self.spec] * N
    return stats

  def predictive_mean_batch_t(self, theta, level):
    """ci_predictive_mean_batch_d: theta [N,S,dim], level [N,S,T] device tensors -> mean [N,T]."""
    torch, dev = self._torch_dev()
    th, lv = theta.contiguous(), level.contiguous()
    N, S = th.shape[0], th.shape[1]
    mean = torch.empty((N, self.spec.T), dtype=th.dtype, device=dev)
    self._check(self._lib.ci_predictive_mean_batch_d(self._ctx, th.data_ptr(), lv.data_ptr(), S,
                                                     mean.data_ptr(), self._stream(torch)))
    return mean

  def impact_batch_t(self, traj, mean, *, scale, offset, obs_sum, observed, period, q_lo, q_hi):
    """ci_impact_batch_d: traj [N,S,T], mean [N,T] device tensors; scale / offset / obs_sum [N],
    observed [N,T], period [T] host arrays.  Returns ONE float64 device tensor [N, T*9 + 20]
    (series then summary per series); nothing is synchronised."""
    torch, dev = self._torch_dev()
    traj, mean = traj.contiguous(), mean.contiguous()
    N, S, T = traj.shape
    dt = F64 if traj.dtype == torch.float64 else F32
    args = CiImpactArgs(S=S, T=T, dtype=dt, reserved=0, scale=1.0, offset=0.0, q_lo=q_lo, q_hi=q_hi,
                        obs_sum=0.0)
    sc = np.ascontiguousarray(scale, dtype=np.float64); of = np.ascontiguousarray(offset, dtype=np.float64)
    os_ = np.ascontiguousarray(obs_sum, dtype=np.float64)
    obs = np.ascontiguousarray(observed, dtype=np.float64)
    per = np.ascontiguousarray(period, dtype=np.uint8)
    if obs.shape != (N, T) or per.shape != (T,) or sc.shape != (N,):
      raise ValueError("observed must be [N,T], period [T], scale / offset / obs_sum [N]")
    series = torch.empty((N, T * IMPACT_SERIES_COLS), dtype=torch.float64, device=dev)
    summ = torch.empty((N, IMPACT_SUMMARY_LEN), dtype=torch.float64, device=dev)
    self._check(self._lib.ci_impact_batch_d(
        self._ctx, C.byref(args), N, _ptr(sc), _ptr(of), _ptr(os_), traj.data_ptr(), mean.data_ptr(),
        _ptr(obs), _ptr(per), series.data_ptr(), summ.data_ptr(), self._stream(torch)))
    return series, summ

  def batch_select(self, i: int, spec: Optional[ProblemSpec] = None):
    """ci_batch_select: series i of the batch becomes the current problem."""
    self._check(self._lib.ci_batch_select(self._ctx, int(i)))
    self.spec = spec if spec is not None else self.batch_specs[i]

  def gibbs_run_batch_t(self, n_chains: int, *, n_warmup: int, n_results: int, seed: int,
                        chain_id0: int = 0, sparse: bool = True,
                        nonzero_prob: Optional[float] = None, ssvs_order: str = "random",
                        series_stride: int = 0):
    """ci_gibbs_run_batch_d, chain-major: device tensors theta [N, R, dim], level [N, R, T],
    traj [N, R, T] (R = n_chains * n_results) and incl [N, n_chains, p] ndarray."""
    torch, dev = self._torch_dev()
    sp, dt, N = self.spec, self._tdtype(torch), len(self.batch_specs)
    if nonzero_prob is None:
      nonzero_prob = min(1.0, 3.0 / sp.p) if sp.p else 1.0     # lib.py:449-450
    rows = n_chains * n_results
    draws = torch.empty((N, rows, sp.dim), dtype=dt, device=dev)
    level = torch.empty((N, rows, sp.T), dtype=dt, device=dev)
    traj = torch.empty((N, rows, sp.T), dtype=dt, device=dev)
    incl = torch.zeros((N, n_chains, max(sp.p, 1)), dtype=torch.float32, device=dev)
    opts = CiGibbsOpts(n_warmup=n_warmup, n_results=n_results, sparse=int(sparse), chain_major=1,
                       nonzero_prob=float(nonzero_prob), ssvs_order=SSVS_ORDER[ssvs_order],
                       reserved=0, series_stride=int(series_stride))
    self._check(self._lib.ci_gibbs_run_batch_d(
        self._ctx, C.byref(opts), seed & (2**64 - 1), chain_id0, n_chains, draws.data_ptr(),
        level.data_ptr(), traj.data_ptr(), incl.data_ptr(), self._stream(torch)))
    return draws, level, traj, incl.cpu().numpy()[:, :, :sp.p]

  def set_seasonal_batch(self, scheds, init_sd=None, drift_scale=None, drift_ub=None):
    """ci_set_seasonal_batch: one season calendar for the panel + per-series prior scales.
    ``scheds``: a list of model.SeasonalSchedule (one per series), or ONE schedule together with
    the per-series arrays init_sd / drift_scale / drift_ub [N]."""
    if isinstance(scheds, (list, tuple)):
      s0 = scheds[0]
      init_sd = [sc.init_sd for sc in scheds]
      drift_scale = [sc.drift_scale for sc in scheds]
      drift_ub = [sc.drift_ub for sc in scheds]
    else:
      s0 = scheds
    act = np.ascontiguousarray(s0.active, dtype=np.uint8)
    ends = np.ascontiguousarray(s0.ends, dtype=np.uint8)
    cs = CiSeasonal(n_components=s0.K, active=act.ctypes.data, ends=ends.ctypes.data,
                    init_sd=s0.init_sd, drift_conc=s0.drift_conc, drift_scale=s0.drift_scale,
                    drift_ub=s0.drift_ub)
    for k, n in enumerate(s0.num_seasons):
      cs.num_seasons[k] = int(n)
    a = np.ascontiguousarray(init_sd, dtype=np.float64)
    b = np.ascontiguousarray(drift_scale, dtype=np.float64)
    u = np.ascontiguousarray(drift_ub, dtype=np.float64)
    self._check(self._lib.ci_set_seasonal_batch(self._ctx, C.byref(cs), _ptr(a), _ptr(b), _ptr(u)))
    self.seasonal = s0

  def gibbs_seasonal_run_batch_t(self, n_chains: int, *, n_warmup: int, n_results: int, seed: int,
                                 chain_id0: int = 0, sparse: bool = True,
                                 nonzero_prob: Optional[float] = None, ssvs_order: str = "random",
                                 series_stride: int = 0):
    """ci_gibbs_seasonal_run_batch_d, chain-major device tensors with a leading series axis:
    (theta [N,R,dim], level, latent, traj [N,R,T], seasonal [N,R,T,K], log drift variance
    [N,R,K], incl [N,C,p] ndarray)."""
    torch, dev = self._torch_dev()
    sp, dt, K, N = self.spec, self._tdtype(torch), self.seasonal.K, len(self.batch_specs)
    if nonzero_prob is None:
      nonzero_prob = min(1.0, 3.0 / sp.p) if sp.p else 1.0
    rows = n_chains * n_results
    mk = lambda *shape: torch.empty(shape, dtype=dt, device=dev)
    draws, level, latent, traj = mk(N, rows, sp.dim), mk(N, rows, sp.T), mk(N, rows, sp.T), mk(N, rows, sp.T)
    seas, drift = mk(N, rows, sp.T, K), mk(N, rows, K)
    incl = torch.zeros((N, n_chains, max(sp.p, 1)), dtype=torch.float32, device=dev)
    opts = CiGibbsOpts(n_warmup=n_warmup, n_results=n_results, sparse=int(sparse), chain_major=1,
                       nonzero_prob=float(nonzero_prob), ssvs_order=SSVS_ORDER[ssvs_order],
                       reserved=0, series_stride=int(series_stride))
    self._check(self._lib.ci_gibbs_seasonal_run_batch_d(
        self._ctx, C.byref(opts), seed & (2**64 - 1), chain_id0, n_chains, draws.data_ptr(),
        level.data_ptr(), traj.data_ptr(), incl.data_ptr(), latent.data_ptr(), seas.data_ptr(),
        drift.data_ptr(), self._stream(torch)))
    return draws, level, latent, traj, seas, drift, incl.cpu().numpy()[:, :, :sp.p]

  def set_seasonal(self, sched):
    """ci_set_seasonal: ``sched`` is a model.SeasonalSchedule (or None to remove)."""
    self.seasonal = sched
    if sched is None or sched.K == 0:
      self.seasonal = None
      self._check(self._lib.ci_set_seasonal(self._ctx, None))
      return
    if sched.K > MAX_SEASONAL:
      raise EngineError(f"at most {MAX_SEASONAL} seasonal components are supported")
    act = np.ascontiguousarray(sched.active, dtype=np.uint8)
    ends = np.ascontiguousarray(sched.ends, dtype=np.uint8)
    if act.shape != (sched.K, self.spec.T) or ends.shape != act.shape:
      raise ValueError(f"seasonal schedule must be [{sched.K},{self.spec.T}]")
    cs = CiSeasonal(n_components=sched.K, active=act.ctypes.data, ends=ends.ctypes.data,
                    init_sd=sched.init_sd, drift_conc=sched.drift_conc,
                    drift_scale=sched.drift_scale, drift_ub=sched.drift_ub)
    for k, n in enumerate(sched.num_seasons):
      cs.num_seasons[k] = int(n)
    self._check(self._lib.ci_set_seasonal(self._ctx, C.byref(cs)))

  def gibbs_seasonal_run(self, n_chains: int, *, n_warmup: int, n_results: int, seed: int,
                         chain_id0: int = 0, sparse: bool = True,
                         nonzero_prob: Optional[float] = None, ssvs_order: str = "random"):
    """Host-buffer ci_gibbs_seasonal_run.  Returns dict: draws [n_results, C, dim], level,
    latent, traj [n_results, C, T], seasonal [n_results, C, T, K], drift [n_results, C, K]
    (drift SCALES), incl [C, p]."""
    sp, K = self.spec, self.seasonal.K
    if nonzero_prob is None:
      nonzero_prob = min(1.0, 3.0 / sp.p) if sp.p else 1.0
    dt = sp.np_dtype
    shp = (n_results, n_chains)
    out = dict(draws=np.empty(shp + (sp.dim,), dt), level=np.empty(shp + (sp.T,), dt),
               traj=np.empty(shp + (sp.T,), dt), latent=np.empty(shp + (sp.T,), dt),
               seasonal=np.empty(shp + (sp.T, K), dt), drift=np.empty(shp + (K,), dt))
    incl = np.zeros((n_chains, max(sp.p, 1)), dtype=np.float32)
    opts = CiGibbsOpts(n_warmup=n_warmup, n_results=n_results, sparse=int(sparse), chain_major=0,
                       nonzero_prob=float(nonzero_prob), ssvs_order=SSVS_ORDER[ssvs_order],
                       reserved=0, series_stride=0)
    self._check(self._lib.ci_gibbs_seasonal_run(
        self._ctx, C.byref(opts), seed & (2**64 - 1), chain_id0, n_chains, _ptr(out["draws"]),
        _ptr(out["level"]), _ptr(out["traj"]), _ptr(incl), _ptr(out["latent"]),
        _ptr(out["seasonal"]), _ptr(out["drift"])))
    out["drift"] = np.exp(0.5 * out["drift"].astype(np.float64)).astype(dt)
    out["incl"] = incl[:, :sp.p]
    return out

  def gibbs_seasonal_run_t(self, n_chains: int, *, n_warmup: int, n_results: int, seed: int,
                           chain_id0: int = 0, sparse: bool = True,
                           nonzero_prob: Optional[float] = None, ssvs_order: str = "random"):
    """ci_gibbs_seasonal_run_d, chain-major device tensors: (theta [R, dim], level [R, T],
    latent [R, T], traj [R, T], seasonal [R, T, K], log drift variance [R, K], incl ndarray),
    R = C * n_results."""
    torch, dev = self._torch_dev()
    sp, dt, K = self.spec, self._tdtype(torch), self.seasonal.K
    if nonzero_prob is None:
      nonzero_prob = min(1.0, 3.0 / sp.p) if sp.p else 1.0
    rows = n_chains * n_results
    mk = lambda *shape: torch.empty(shape, dtype=dt, device=dev)
    draws, level, latent, traj = mk(rows, sp.dim), mk(rows, sp.T), mk(rows, sp.T), mk(rows, sp.T)
    seas, drift = mk(rows, sp.T, K), mk(rows, K)
    incl = torch.zeros((n_chains, max(sp.p, 1)), dtype=torch.float32, device=dev)
    opts = CiGibbsOpts(n_warmup=n_warmup, n_results=n_results, sparse=int(sparse), chain_major=1,
                       nonzero_prob=float(nonzero_prob), ssvs_order=SSVS_ORDER[ssvs_order],
                       reserved=0, series_stride=0)
    self._check(self._lib.ci_gibbs_seasonal_run_d(
        self._ctx, C.byref(opts), seed & (2**64 - 1), chain_id0, n_chains, draws.data_ptr(),
        level.data_ptr(), traj.data_ptr(), incl.data_ptr(), latent.data_ptr(), seas.data_ptr(),
        drift.data_ptr(), self._stream(torch)))
    return draws, level, latent, traj, seas, drift, incl.cpu().numpy()[:, :sp.p]

  # -- K1/K2/K3 --------------------------------------------------------------
  def logprob(self, theta, variant: int = VARIANT_SCAN, with_prior: bool = False) -> np.ndarray:
    theta = self._arr(np.atleast_2d(theta))
    n = theta.shape[0]
    if theta.shape[1] != self.spec.dim:
      raise ValueError(f"theta must be [C,{self.spec.dim}]")
    val = np.empty(n, dtype=self.spec.np_dtype)
    self._check(self._lib.ci_logprob(self._ctx, _ptr(theta), n, _ptr(val), variant,
                                     WITH_PRIOR if with_prior else 0))
    return val

  def logprob_grad(self, theta, variant: int = VARIANT_SCAN,
                   with_prior: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    theta = self._arr(np.atleast_2d(theta))
    n = theta.shape[0]
    if theta.shape[1] != self.spec.dim:
      raise ValueError(f"theta must be [C,{self.spec.dim}]")
    val = np.empty(n, dtype=self.spec.np_dtype)
    grad = np.empty_like(theta)
    self._check(self._lib.ci_logprob_grad(self._ctx, _ptr(theta), n, _ptr(val), _ptr(grad),
                                          variant, WITH_PRIOR if with_prior else 0))
    return val, grad

  def logprob_grad_ptr(self, theta_ptr: int, n_chains: int, value_ptr: int, grad_ptr: int,
                       variant: int = VARIANT_SCAN, flags: int = 0, stream: int = 0,
                       host: bool = False):
    """Raw-pointer entry (device pointers + stream, or pinned host pointers)."""
    if host:
      self._check(self._lib.ci_logprob_grad(self._ctx, theta_ptr, n_chains, value_ptr,
                                            grad_ptr or None, variant, flags))
    else:
      self._check(self._lib.ci_logprob_grad_d(self._ctx, theta_ptr, n_chains, value_ptr,
                                              grad_ptr or None, variant, flags, stream or None))

  # -- K6 --------------------------------------------------------------------
  def hmc_run(self, theta0, *, n_warmup: int, n_results: int, seed: int, chain_id0: int = 0,
              max_leapfrog: int = 8, init_step: float = 0.05, target_accept: float = 0.8,
              adapt_mass: bool = True):
    theta0 = self._arr(np.atleast_2d(theta0))
    n = theta0.shape[0]
    draws = np.empty((n_results, n, self.spec.dim), dtype=self.spec.np_dtype)
    stats = np.zeros(n, dtype=HMC_STATS_DTYPE)
    opts = CiHmcOpts(n_warmup=n_warmup, n_results=n_results, max_leapfrog=max_leapfrog,
                     adapt_mass=int(adapt_mass), init_step=init_step,
                     target_accept=target_accept)
    self._check(self._lib.ci_hmc_run(self._ctx, C.byref(opts), seed & (2**64 - 1), chain_id0,
                                     _ptr(theta0), n, _ptr(draws), _ptr(stats)))
    return draws, stats

  # -- the reference's Gibbs sampler (spike-and-slab) --------------------------
  def gibbs_run(self, n_chains: int, *, n_warmup: int, n_results: int, seed: int,
                chain_id0: int = 0, sparse: bool = True, nonzero_prob: Optional[float] = None,
                want_level: bool = True, want_traj: bool = True, ssvs_order: str = "random"):
    """Returns draws [n_results, C, dim], level, traj [n_results, C, T], incl [C, p]."""
    sp = self.spec
    if nonzero_prob is None:
      nonzero_prob = min(1.0, 3.0 / sp.p) if sp.p else 1.0     # lib.py:449-450
    dt = sp.np_dtype
    draws = np.empty((n_results, n_chains, sp.dim), dtype=dt)
    level = np.empty((n_results, n_chains, sp.T), dtype=dt) if want_level else None
    traj = np.empty((n_results, n_chains, sp.T), dtype=dt) if want_traj else None
    incl = np.zeros((n_chains, max(sp.p, 1)), dtype=np.float32)
    opts = CiGibbsOpts(n_warmup=n_warmup, n_results=n_results, sparse=int(sparse), chain_major=0,
                       nonzero_prob=float(nonzero_prob), ssvs_order=SSVS_ORDER[ssvs_order],
                       reserved=0, series_stride=0)
    self._check(self._lib.ci_gibbs_run(self._ctx, C.byref(opts), seed & (2**64 - 1), chain_id0,
                                       n_chains, _ptr(draws), _ptr(level), _ptr(traj),
                                       _ptr(incl)))
    return draws, level, traj, incl[:, :sp.p]

  # -- K4 --------------------------------------------------------------------
  def posterior_predict(self, theta_draws, *, seed: int, draw_id0: int = 0,
                        want_level: bool = True):
    th = self._arr(np.atleast_2d(theta_draws))
    S, T = th.shape[0], self.spec.T
    level = np.empty((S, T), dtype=self.spec.np_dtype) if want_level else None
    traj = np.empty((S, T), dtype=self.spec.np_dtype)
    mean = np.empty(T, dtype=self.spec.np_dtype)
    self._check(self._lib.ci_posterior_predict(self._ctx, _ptr(th), S, seed & (2**64 - 1),
                                               draw_id0, _ptr(level), _ptr(traj), _ptr(mean)))
    return level, traj, mean

  # -- device-resident variants (torch CUDA tensors as buffers, *_d entry points) ---
  def _torch_dev(self):
    import torch
    return torch, torch.device("cuda", self.device)

  def synchronize(self):
    """Wait for everything queued on this engine's device."""
    import torch
    torch.cuda.synchronize(self.device)

  def torch_device(self):
    """torch.device of this engine's GPU."""
    return self._torch_dev()[1]

  def _tdtype(self, torch):
    return torch.float64 if self.spec.dtype == F64 else torch.float32

  def _stream(self, torch) -> int:
    return torch.cuda.current_stream(self.device).cuda_stream

  def _as_dev(self, a):
    """numpy / tensor -> contiguous tensor of the problem dtype on this device."""
    torch, dev = self._torch_dev()
    if not hasattr(a, "data_ptr"):
      a = torch.from_numpy(np.ascontiguousarray(a, dtype=self.spec.np_dtype))
    return a.to(device=dev, dtype=self._tdtype(torch)).contiguous()

  def hmc_run_t(self, theta0, *, n_warmup: int, n_results: int, seed: int, chain_id0: int = 0,
                max_leapfrog: int = 8, init_step: float = 0.05, target_accept: float = 0.8,
                adapt_mass: bool = True):
    """ci_hmc_run_d: returns (draws tensor [n_results, C, dim] on the device, stats ndarray)."""
    torch, dev = self._torch_dev()
    th = self._as_dev(np.atleast_2d(theta0) if not hasattr(theta0, "data_ptr") else theta0)
    n = th.shape[0]
    draws = torch.empty((n_results, n, self.spec.dim), dtype=th.dtype, device=dev)
    stats = torch.zeros(n * HMC_STATS_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    opts = CiHmcOpts(n_warmup=n_warmup, n_results=n_results, max_leapfrog=max_leapfrog,
                     adapt_mass=int(adapt_mass), init_step=init_step,
                     target_accept=target_accept)
    self._check(self._lib.ci_hmc_run_d(self._ctx, C.byref(opts), seed & (2**64 - 1), chain_id0,
                                       th.data_ptr(), n, draws.data_ptr(), stats.data_ptr(),
                                       self._stream(torch)))
    return draws, stats.cpu().numpy().view(HMC_STATS_DTYPE)

  def gibbs_run_t(self, n_chains: int, *, n_warmup: int, n_results: int, seed: int,
                  chain_id0: int = 0, sparse: bool = True, nonzero_prob: Optional[float] = None,
                  ssvs_order: str = "random"):
    """ci_gibbs_run_d, chain-major: returns (theta [C*n_results, dim], level [C*n_results, T],
    traj [C*n_results, T]) device tensors -- row c*n_results + i is kept sweep i of chain c --
    and incl [C, p] ndarray."""
    torch, dev = self._torch_dev()
    sp, dt = self.spec, self._tdtype(torch)
    if nonzero_prob is None:
      nonzero_prob = min(1.0, 3.0 / sp.p) if sp.p else 1.0     # lib.py:449-450
    rows = n_chains * n_results
    draws = torch.empty((rows, sp.dim), dtype=dt, device=dev)
    level = torch.empty((rows, sp.T), dtype=dt, device=dev)
    traj = torch.empty((rows, sp.T), dtype=dt, device=dev)
    incl = torch.zeros((n_chains, max(sp.p, 1)), dtype=torch.float32, device=dev)
    opts = CiGibbsOpts(n_warmup=n_warmup, n_results=n_results, sparse=int(sparse), chain_major=1,
                       nonzero_prob=float(nonzero_prob), ssvs_order=SSVS_ORDER[ssvs_order],
                       reserved=0, series_stride=0)
    self._check(self._lib.ci_gibbs_run_d(self._ctx, C.byref(opts), seed & (2**64 - 1), chain_id0,
                                         n_chains, draws.data_ptr(), level.data_ptr(),
                                         traj.data_ptr(), incl.data_ptr(), self._stream(torch)))
    return draws, level, traj, incl.cpu().numpy()[:, :sp.p]

  def posterior_predict_t(self, theta_draws, *, seed: int, draw_id0: int = 0):
    """ci_posterior_predict_d: (level [S,T], traj [S,T]) device tensors."""
    torch, dev = self._torch_dev()
    th = self._as_dev(theta_draws)
    S, T = th.shape[0], self.spec.T
    level = torch.empty((S, T), dtype=th.dtype, device=dev)
    traj = torch.empty((S, T), dtype=th.dtype, device=dev)
    self._check(self._lib.ci_posterior_predict_d(self._ctx, th.data_ptr(), S, seed & (2**64 - 1),
                                                 draw_id0, level.data_ptr(), traj.data_ptr(),
                                                 None, self._stream(torch)))
    return level, traj

  def predictive_mean_t(self, theta_draws, level):
    """ci_predictive_mean_d (lib.py:627): [T] device tensor."""
    torch, dev = self._torch_dev()
    th, lv = self._as_dev(theta_draws), self._as_dev(level)
    mean = torch.empty((self.spec.T,), dtype=th.dtype, device=dev)
    self._check(self._lib.ci_predictive_mean_d(self._ctx, th.data_ptr(), lv.data_ptr(),
                                               th.shape[0], mean.data_ptr(), self._stream(torch)))
    return mean

  def to_host(self, t) -> np.ndarray:
    """Device tensor -> ndarray.  Large arrays (the [S,T] level paths) go through pinned
    memory: 1.4 ms instead of 38 ms for 80 MB on the B200 box (run 21); the returned
    array aliases the pinned block, which torch's host allocator recycles once freed."""
    t = t.detach()
    if t.is_cuda and t.numel() * t.element_size() >= (1 << 20):
      import torch
      host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
      host.copy_(t)                      # synchronous for device -> pinned host
      return host.numpy()
    return t.cpu().numpy()

  # -- impact series + summary -----------------------------------
  def impact(self, traj, mean, meta, out=None) -> Tuple[np.ndarray, np.ndarray]:
    """ci_impact: (series [T,9], summary [20]) float64 from predictive draws [S,T] and mean [T]
    on the standardized scale.  ``traj`` / ``mean`` may be DeviceArray / torch CUDA tensors
    (no copy: ci_impact_d) or host arrays (ci_impact uploads them once).  ``meta``:
    impact.ImpactMeta.  With ``out`` (a float64 device tensor of T*9 + 20 elements, device
    inputs only) the result is left there and nothing is synchronised or returned -- the
    panel path queues one call per series and reads everything back once."""
    if isinstance(traj, DeviceArray):
      traj = traj.tensor
    if isinstance(mean, DeviceArray):
      mean = mean.tensor
    on_dev = hasattr(traj, "data_ptr")
    obs = np.ascontiguousarray(meta.observed, dtype=np.float64)
    per = np.ascontiguousarray(meta.period, dtype=np.uint8)
    if on_dev:
      torch, dev = self._torch_dev()
      if traj.dtype not in (torch.float32, torch.float64):
        traj = traj.to(torch.float32)
      traj = traj.to(dev).contiguous()
      mean_t = mean if hasattr(mean, "data_ptr") else torch.from_numpy(np.ascontiguousarray(mean))
      mean_t = mean_t.to(device=dev, dtype=traj.dtype).contiguous().reshape(-1)
      S, T = traj.shape
      dt = F64 if traj.dtype == torch.float64 else F32
    else:
      traj = np.ascontiguousarray(traj)
      if traj.dtype not in (np.float32, np.float64):
        traj = traj.astype(np.float64)
      mean_h = np.ascontiguousarray(np.asarray(mean).reshape(-1), dtype=traj.dtype)
      S, T = traj.shape
      dt = F64 if traj.dtype == np.float64 else F32
    if obs.shape != (T,) or per.shape != (T,):
      raise ValueError(f"observed / period must be [{T}]")
    args = CiImpactArgs(S=S, T=T, dtype=dt, reserved=0, scale=meta.scale, offset=meta.offset,
                        q_lo=meta.q_lo, q_hi=meta.q_hi, obs_sum=meta.obs_sum)
    if on_dev:
      queued = out is not None
      if not queued:
        out = torch.empty(T * IMPACT_SERIES_COLS + IMPACT_SUMMARY_LEN, dtype=torch.float64,
                          device=dev)
      elif out.numel() != T * IMPACT_SERIES_COLS + IMPACT_SUMMARY_LEN or \
          out.dtype != torch.float64 or not out.is_contiguous():
        raise ValueError("out must be a contiguous float64 tensor of T*9 + 20 elements")
      self._check(self._lib.ci_impact_d(
          self._ctx, C.byref(args), traj.data_ptr(), mean_t.data_ptr(), _ptr(obs), _ptr(per),
          out.data_ptr(), out.data_ptr() + 8 * T * IMPACT_SERIES_COLS, self._stream(torch)))
      if queued:
        return None
      out = out.cpu().numpy()
      return out[:T * IMPACT_SERIES_COLS].reshape(T, IMPACT_SERIES_COLS), \
          out[T * IMPACT_SERIES_COLS:]
    series = np.empty((T, IMPACT_SERIES_COLS), dtype=np.float64)
    summ = np.empty(IMPACT_SUMMARY_LEN, dtype=np.float64)
    self._check(self._lib.ci_impact(self._ctx, C.byref(args), _ptr(traj), _ptr(mean_h), _ptr(obs),
                                    _ptr(per), _ptr(series), _ptr(summ)))
    return series, summ

  # -- the two halves of ci_impact_d for draws sharded over GPUs ----------
  def _impact_args(self, meta, S, T, torch_dtype):
    import torch
    return CiImpactArgs(S=S, T=T, dtype=F64 if torch_dtype == torch.float64 else F32, reserved=0,
                        scale=meta.scale, offset=meta.offset, q_lo=meta.q_lo, q_hi=meta.q_hi,
                        obs_sum=meta.obs_sum)

  def impact_rows_t(self, traj, mean, meta, out=None):
    """ci_impact_rows_d on this rank's draws: traj [S_local, T] device tensor; ``mean`` = the
    predictive mean over ALL draws ([T] device tensor) on the one rank that writes the
    mean-derived columns into ``out`` (float64 [T*9 + 20]), None elsewhere.  Returns device
    tensors (trT [T,S_local], cumT [T - t_c0, S_local], stats [5, S_local]); nothing is
    synchronised.  The
    fourth value is the one float64 block [5 + T - t_c0, S_local] that stats and cumT are views of."""
    torch, dev = self._torch_dev()
    traj = traj.contiguous()
    S, T = traj.shape
    obs = np.ascontiguousarray(meta.observed, dtype=np.float64)
    per = np.ascontiguousarray(meta.period, dtype=np.uint8)
    if obs.shape != (T,) or per.shape != (T,):
      raise ValueError(f"observed / period must be [{T}]")
    t_c0 = int(np.argmax(per != 0)) if np.any(per != 0) else T
    args = self._impact_args(meta, S, T, traj.dtype)
    trT = torch.empty((T, S), dtype=traj.dtype, device=dev)
    # one float64 block [5 + (T - t_c0), S]: the statistics ride in front of the cumulative paths
    # so that a sharded caller exchanges both with one collective (shard.impact_sharded)
    packed = torch.empty((5 + T - t_c0, S), dtype=torch.float64, device=dev)
    stats, cumT = packed[:5], packed[5:]
    if mean is not None:
      mean = mean.to(dtype=traj.dtype).contiguous().reshape(-1)
      if out is None or out.numel() != T * IMPACT_SERIES_COLS + IMPACT_SUMMARY_LEN:
        raise ValueError("out (float64, T*9 + 20 elements) is required with mean")
    self._check(self._lib.ci_impact_rows_d(
        self._ctx, C.byref(args), traj.data_ptr(), mean.data_ptr() if mean is not None else None,
        _ptr(obs), _ptr(per), trT.data_ptr(), cumT.data_ptr(), stats.data_ptr(),
        out.data_ptr() if mean is not None else None,
        out.data_ptr() + 8 * T * IMPACT_SERIES_COLS if mean is not None else None,
        self._stream(torch)))
    return trT, cumT, stats, packed

  def impact_cols_t(self, trT, t_begin, cumT, c_begin, stats, meta, out):
    """ci_impact_cols_d: the quantile columns of a time block.  trT [t_count, S] and cumT
    [c_count, S] hold ALL draws of prediction steps t_begin.. / cumulative steps c_begin..;
    ``stats`` ([5, S], all draws) on the one rank that writes summary[0..17], None elsewhere.
    Fills this block's entries of ``out`` (float64 [T*9 + 20]); nothing is synchronised."""
    torch, dev = self._torch_dev()
    obs = np.ascontiguousarray(meta.observed, dtype=np.float64)
    per = np.ascontiguousarray(meta.period, dtype=np.uint8)
    T = obs.shape[0]
    trT, cumT = trT.contiguous(), cumT.contiguous()
    S = trT.shape[1] if trT.shape[0] else (cumT.shape[1] if cumT.shape[0] else stats.shape[1])
    if stats is not None:
      stats = stats.contiguous()
    args = self._impact_args(meta, S, T, trT.dtype)
    self._check(self._lib.ci_impact_cols_d(
        self._ctx, C.byref(args), trT.data_ptr() if trT.shape[0] else None, int(t_begin),
        trT.shape[0], cumT.data_ptr() if cumT.shape[0] else None, int(c_begin), cumT.shape[0],
        stats.data_ptr() if stats is not None else None, _ptr(obs), _ptr(per), out.data_ptr(),
        out.data_ptr() + 8 * T * IMPACT_SERIES_COLS, self._stream(torch)))

  # -- K5 --------------------------------------------------------------------
  def row_quantiles(self, a, q) -> np.ndarray:
    """Per-time quantiles of a [S, T] draw-major array -> [T, len(q)].

    Columns of up to ~50 900 (float32) / ~25 000 (float64) draws are selected in shared memory;
    longer ones straight from global memory (slower per draw, no size limit)."""
    a = np.ascontiguousarray(a)
    if a.dtype not in (np.float32, np.float64):
      a = a.astype(np.float64)
    S, T = a.shape
    out_dtype = a.dtype
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.empty((T, q.shape[0]), dtype=a.dtype)
    self._check(self._lib.ci_row_quantiles(
        self._ctx, _ptr(a), S, T, F64 if a.dtype == np.float64 else F32,
        q.ctypes.data_as(C.POINTER(C.c_double)), q.shape[0], _ptr(out)))
    return out.astype(out_dtype, copy=False)
