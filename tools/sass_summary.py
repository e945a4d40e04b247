#!/usr/bin/env python
"""Developer tool: per-kernel SASS evidence of the in-tree library -> profiles/r02_sass_summary.txt
(cuobjdump -sass; what proves the sm_100a-native structure: UBLKCP = cp.async.bulk (TMA 1-D) tile
copies, SYNCS = mbarrier try_wait / arrive, BAR = named team barriers, SHFL = warp-shuffle scans,
no HMMA / tensor-core ops, LDL / STL = register spills)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "tfp-causalimpact_b200", "lib", "libci_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ["UBLKCP", "UTMALDG", "SYNCS", "BAR", "SHFL", "MUFU", "LDS", "STS", "LDG", "STG", "LDL", "STL",
       "FFMA", "DFMA", "HMMA", "REDUX", "ATOMS"]
rows, cur, arch = {}, None, set()
for ln in out.splitlines():
  m = re.search(r"Function : (\S+)", ln)
  if m:
    cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("ci::", "")
    rows[cur] = collections.Counter(); continue
  m = re.search(r"arch = (sm_\w+)", ln)
  if m:
    arch.add(m.group(1))
  m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", ln)
  if m and cur:
    rows[cur]["total"] += 1
    op = m.group(1)
    for o in OPS:
      if op == o or op.startswith(o + "."):
        rows[cur][o] += 1
lines = [f"SASS summary of tfp-causalimpact_b200/lib/libci_b200.so  (arch: {', '.join(sorted(arch))})",
         "kernel".ljust(44) + "".join(o.rjust(8) for o in ["total"] + OPS)]
tot = collections.Counter()
for k in sorted(rows):
  lines.append(k[:43].ljust(44) + "".join(str(rows[k][o]).rjust(8) for o in ["total"] + OPS))
  tot.update(rows[k])
lines.append("ALL".ljust(44) + "".join(str(tot[o]).rjust(8) for o in ["total"] + OPS))
txt = "\n".join(lines) + "\n"
open(os.path.join(ROOT, "profiles", "r02_sass_summary.txt"), "w").write(txt)
print(txt)
