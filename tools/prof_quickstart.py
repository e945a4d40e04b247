"""Developer tool: cProfile of fit_causalimpact on the quickstart shape."""
import os, sys, time, cProfile, pstats, io
import numpy as np, pandas as pd
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tfp-causalimpact_b200")):
  sys.path.insert(0, p)
import causalimpact_b200 as cib
rs = np.random.Generator(np.random.PCG64(20241))
xq = 100 + np.cumsum(rs.normal(size=100)) * 0.3
yq = 1.2 * xq + rs.normal(size=100); yq[71:] += 10
df = pd.DataFrame({"y": yq, "x": xq})
for _ in range(3):
  cib.fit_causalimpact(df, (0, 70), (71, 99), seed=1)
t0 = time.perf_counter()
for _ in range(10):
  cib.fit_causalimpact(df, (0, 70), (71, 99), seed=1)
print("fit ms", (time.perf_counter() - t0) / 10 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
  cib.fit_causalimpact(df, (0, 70), (71, 99), seed=1)
pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative"); buf = io.StringIO(); st.stream = buf; st.print_stats(30)
print("\n".join(buf.getvalue().splitlines()[6:42]))
