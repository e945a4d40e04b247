"""Per-time quantiles exactly as the reference computes them: pandas
``DataFrame.quantile(q, axis=1)`` on the [T, S] frame
(causalimpact/posterior_processing.py:56).  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import pandas as pd


def row_quantiles(a_st: np.ndarray, q) -> np.ndarray:
  """a_st [S, T] draw-major (as the engine takes it) -> [T, len(q)]."""
  frame = pd.DataFrame(np.asarray(a_st, dtype=np.float64).T)
  return frame.quantile(q=list(np.asarray(q, dtype=np.float64)), axis=1).transpose().values
