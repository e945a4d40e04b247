"""TFP-conformance goldens: run on a box that HAS tensorflow + tensorflow_probability.

TEST INFRASTRUCTURE ONLY.  This container has neither (no network, not in /opt/wheelhouse), so
the Kalman / sampler half of the oracle is "parity unpinned" against the reference (DESIGN.md
section 3).  This script is the way out: on any machine where

    pip install tensorflow tensorflow-probability tfp-causalimpact     (or PYTHONPATH=/root/reference)

works, run

    python oracle/make_golden_tfp.py [--out tests/golden]

and commit the `tfp_*.npz` files it writes.  tests/test_tfp_conformance.py consumes them when
present (and skips LOUDLY when not):

  tfp_logprob_<case>.npz   tfd.LinearGaussianStateSpaceModel(...).log_prob of the masked residual
                           series for a batch of theta -- pins oracle/kalman_np.log_lik (and through
                           it the CUDA kernels) to TFP's filter: conventions (update-then-predict,
                           prior at t = 0), masking, float32 level.
  tfp_smoother_<case>.npz  posterior_marginals means / variances -- pins oracle/smoother_np.
  tfp_fit_<case>.npz       the UNMODIFIED reference `causalimpact.fit_causalimpact` on seeded inputs:
                           posterior-sample summaries (means, sds, inclusion frequencies, clamps
                           observed), the series and summary frames -- pins the sampler semantics
                           this repo had to guess (upper_bound on variance vs
                           scale, SSVS visiting order, weight adjustment) statistically.

Inputs are generated with numpy PCG64 (no TFP RNG), so the same arrays are rebuilt by the test.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "tests")):
  if p not in sys.path:
    sys.path.insert(0, p)

LOGPROB_CASES = [("quickstart", 100, 1, 8, 1), ("cfg2_small", 400, 10, 16, 2), ("gaps", 300, 3, 8, 3)]
FIT_CASES = [  # (name, T, n_cov, seed, num_results)
    ("quickstart", 100, 1, 11, 1000), ("sparse10", 300, 10, 12, 1000), ("nocov", 120, 0, 13, 1000),
    ("unstandardized", 150, 2, 14, 1000)]


def fit_inputs(T, n_cov, seed):
  """Raw (un-standardized) frame of the quickstart recipe (docs/quickstart.ipynb:280-295) with
  numpy's RNG: returns values [T, 1 + n_cov], pre = (0, t_pre - 1), post = (t_pre, T - 1)."""
  rng = np.random.Generator(np.random.PCG64(seed))
  t_pre = int(round(0.7 * T))
  xs = np.empty((T, n_cov))
  for j in range(n_cov):
    a = np.empty(T); a[0] = rng.normal()
    eps = rng.normal(size=T)
    for t in range(1, T):
      a[t] = 0.999 * a[t - 1] + eps[t]
    xs[:, j] = 100.0 + a
  beta = np.zeros(max(n_cov, 1)); beta[:3] = (1.2, 0.6, -0.4)[:min(3, max(n_cov, 1))]
  if n_cov:
    y = xs @ beta[:n_cov] + rng.normal(size=T)
  else:
    y = 100.0 + np.cumsum(0.05 * rng.normal(size=T)) + rng.normal(size=T)
  y[t_pre:] += 10.0
  return np.column_stack([y, xs]), (0, t_pre - 1), (t_pre, T - 1)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
  args = ap.parse_args()
  try:
    import tensorflow as tf
    import tensorflow_probability as tfp
  except ImportError as e:
    raise SystemExit(f"tensorflow / tensorflow_probability not importable here ({e}); run this "
                     "script on a TFP-equipped box") from e
  try:
    import causalimpact
  except ImportError:
    sys.path.insert(0, "/root/reference")
    import causalimpact
  import pandas as pd
  from conftest import make_series, make_thetas
  from oracle import kalman_np as K
  tfd = tfp.distributions
  os.makedirs(args.out, exist_ok=True)
  versions = np.array([tf.__version__, tfp.__version__])

  # ---- 1 + 2: TFP's own filter / smoother on the oracle's inputs ----
  for name, T, n_cov, C, seed in LOGPROB_CASES:
    y, X, _ = make_series(T, n_cov, seed, nan_frac=0.05 if name == "gaps" else 0.01)
    prob = K.default_problem(y, X)
    th = make_thetas(prob.dim, prob.p, C, seed + 100)
    lp = np.empty(C); sm_mean = np.empty((C, T)); sm_var = np.empty((C, T))
    mask = np.isnan(y)
    for c in range(C):
      w, s_e, s_h = th[c, :prob.p], np.exp(th[c, prob.p]), np.exp(th[c, prob.p + 1])
      r = np.where(mask, 0.0, y - (X @ w if prob.p else 0.0))
      model = tfd.LinearGaussianStateSpaceModel(
          num_timesteps=T,
          transition_matrix=tf.linalg.LinearOperatorIdentity(1, dtype=tf.float64),
          transition_noise=tfd.MultivariateNormalDiag(scale_diag=tf.constant([np.sqrt(s_h)], tf.float64)),
          observation_matrix=tf.linalg.LinearOperatorIdentity(1, dtype=tf.float64),
          observation_noise=tfd.MultivariateNormalDiag(scale_diag=tf.constant([np.sqrt(s_e)], tf.float64)),
          initial_state_prior=tfd.MultivariateNormalDiag(
              loc=tf.constant([prob.m0], tf.float64),
              scale_diag=tf.constant([np.sqrt(prob.P0)], tf.float64)))
      obs = tf.constant(r[:, None], tf.float64)
      lp[c] = float(model.log_prob(obs, mask=tf.constant(mask)))
      means, covs = model.posterior_marginals(obs, mask=tf.constant(mask))
      sm_mean[c] = means.numpy()[:, 0]; sm_var[c] = covs.numpy()[:, 0, 0]
    np.savez(os.path.join(args.out, f"tfp_logprob_{name}.npz"), T=T, n_cov=n_cov, C=C, seed=seed,
             nan_frac=0.05 if name == "gaps" else 0.01, theta=th, log_prob=lp, versions=versions)
    np.savez(os.path.join(args.out, f"tfp_smoother_{name}.npz"), T=T, n_cov=n_cov, C=C, seed=seed,
             nan_frac=0.05 if name == "gaps" else 0.01, theta=th, mean=sm_mean, var=sm_var,
             versions=versions)
    print("wrote tfp_logprob / tfp_smoother", name)

  # ---- 3: the unmodified reference fit ----
  for name, T, n_cov, seed, num_results in FIT_CASES:
    vals, pre, post = fit_inputs(T, n_cov, seed)
    df = pd.DataFrame(vals, columns=["y"] + [f"x{j}" for j in range(n_cov)])
    res = causalimpact.fit_causalimpact(
        df, pre, post, seed=(0, seed),
        data_options=causalimpact.DataOptions(standardize_data=name != "unstandardized"),
        inference_options=causalimpact.InferenceOptions(num_results=num_results))
    ps = res.posterior_samples
    num = lambda a: None if a is None else np.asarray(a)
    w = num(ps.weights)
    out = dict(T=T, n_cov=n_cov, seed=seed, num_results=num_results,
               standardize=name != "unstandardized", versions=versions,
               observation_noise_scale=num(ps.observation_noise_scale), level_scale=num(ps.level_scale),
               level_mean=num(ps.level).mean(0), level_sd=num(ps.level).std(0),
               series=res.series[[c for c in res.series.columns
                                  if not c.endswith(("_start", "_end"))]].values.astype(np.float64),
               series_columns=np.array([c for c in res.series.columns if not c.endswith(("_start", "_end"))]),
               summary=res.summary.values.astype(np.float64),
               summary_columns=np.array(list(res.summary.columns)))
    if w is not None:
      out.update(weights_mean=w.mean(0), weights_sd=w.std(0), inclusion=(w != 0).mean(0))
    np.savez(os.path.join(args.out, f"tfp_fit_{name}.npz"), **out)
    print("wrote tfp_fit", name)


if __name__ == "__main__":
  main()
