#!/usr/bin/env python
"""Developer tool: per SOURCE LINE instruction counts and stall samples of one kernel in an
ncu report (ncu's csv source page prints metrics per SASS instruction only; nvdisasm -gi of
the in-tree library supplies the line of each instruction, inlining included).

  tools/ncu_lines.py gpurun_out/x.ncu-rep k_logpost_tstream [float] [top]
"""
import csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
tmpl = sys.argv[3] if len(sys.argv) > 3 else "float"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
lib = os.path.join(ROOT, "tfp-causalimpact_b200", "lib", "libci_b200.so")

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
sect, cur = {}, None
for r in rows:
  if r and r[0] == "Kernel Name":
    cur = r[1]; sect[cur] = []
  elif cur is not None:
    sect[cur].append(r)
name = [k for k in sect if kern in k and ("<%s" % tmpl) in k]
if not name:
  name = [k for k in sect if kern in k]
name = name[0]
body = sect[name]
hdr = body[0]
ix = {h: i for i, h in enumerate(hdr)}
insts = []
for r in body[1:]:
  if len(r) != len(hdr):
    continue
  insts.append((r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)))

with tempfile.TemporaryDirectory() as td:
  subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, capture_output=True)
  lines = None
  mangled = "%d%sI%s" % (len(kern), kern, {"float": "f", "double": "d"}[tmpl])
  extra = sys.argv[5] if len(sys.argv) > 5 else ""     # e.g. Li2ELi4 to pick one instantiation
  for f in sorted(os.listdir(td)):
    out = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(td, f)], capture_output=True, text=True).stdout
    if mangled not in out:
      continue
    # cut the function's text section
    m = re.search(r"\.section\s+\.text\.[^\n]*%s[^\n]*\n" % (re.escape(mangled) + re.escape(extra)), out)
    if not m:
      continue
    seg = out[m.end():]
    nxt = re.search(r"\n\s*\.section\s", seg)
    seg = seg[:nxt.start()] if nxt else seg
    lines = seg.splitlines()
    break
if lines is None:
  sys.exit("kernel not found in the library")
cur_line, per_inst, fresh, cur_ext = "?", [], True, False
for ln in lines:
  m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
  if m:
    # a group of //## lines precedes an instruction: innermost frame first, then its inline
    # callers; keep the innermost frame that is in our own sources
    ours = "/csrc/" in m.group(1)
    if fresh or (cur_ext and ours):
      cur_line = os.path.basename(m.group(1)) + ":" + m.group(2)
      cur_ext = not ours
    fresh = False
    continue
  if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
    per_inst.append(cur_line)
    fresh = True
if len(per_inst) != len(insts):
  print(f"warning: {len(per_inst)} disassembled vs {len(insts)} profiled instructions", file=sys.stderr)
agg = {}
for (src, n, s), where in zip(insts, per_inst):
  a = agg.setdefault(where, [0, 0])
  a[0] += n; a[1] += s
tot_n = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
print(f"{name}\n  {tot_n} warp instructions, {tot_s} stall samples")
srcs = {}
def text(where):
  f, l = where.split(":") if ":" in where else (where, "0")
  for d in ("tfp-causalimpact_b200/csrc",):
    p = os.path.join(ROOT, d, f)
    if os.path.exists(p):
      if p not in srcs:
        srcs[p] = open(p).read().splitlines()
      return srcs[p][int(l) - 1].strip()[:70] if int(l) - 1 < len(srcs[p]) else ""
  return ""
for where, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
  print(f"{100 * n / tot_n:5.1f}% inst {100 * s / tot_s:5.1f}% stall  {where:28s} {text(where)}")
