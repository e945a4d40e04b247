#!/usr/bin/env python
"""Developer tool: read ncu --set full reports and (re)write profiles/r02_traffic.json, the table
bench.py takes `roofline.traffic` from (DRAM bytes read + written per launch, per kernel), plus a
short markdown summary per kernel (duration, warps active, issue active, top stalls).

  tools/ncu_traffic.py gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]
"""
import csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_traffic.json")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short_name(full):
  m = re.match(r"(?:void )?(?:ci::)?(\w+)<(\w+)", full)
  return f"{m.group(1)}<{m.group(2)}>" if m else full


table = json.load(open(OUT)) if os.path.exists(OUT) else {}
for rep in sys.argv[1:]:
  raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(raw.splitlines()))
  hdr, units = rows[0], rows[1]
  ix = {h: i for i, h in enumerate(hdr)}
  for r in rows[2:]:
    if len(r) != len(hdr):
      continue
    def val(name, scale_unit=False):
      v = float(r[ix[name]].replace(",", "")) if r[ix[name]] not in ("", "n/a") else float("nan")
      return v * UNIT.get(units[ix[name]], 1) if scale_unit else v
    name = short_name(r[ix["Kernel Name"]])
    stalls = {h.split("issue_stalled_")[1].split("_per_")[0]: val(h) for h in hdr
              if "smsp__average_warps_issue_stalled_" in h and h.endswith("_per_issue_active.ratio")}
    top = sorted(((v, k) for k, v in stalls.items() if k != "selected"), reverse=True)[:4]
    table[name] = {
        "dram_bytes_read": val("dram__bytes_read.sum", True),
        "dram_bytes_write": val("dram__bytes_write.sum", True),
        "duration_us_under_ncu": val("gpu__time_duration.sum"),
        "grid": r[ix["Grid Size"]] if "Grid Size" in ix else None,
        "block": r[ix["Block Size"]] if "Block Size" in ix else None,
        "registers_per_thread": val("launch__registers_per_thread"),
        "warp_instructions": val("smsp__inst_executed.sum"),
        "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "top_stalls_per_issue": {k: round(v, 3) for v, k in top},
        "source": os.path.basename(rep),
    }
    print(name, json.dumps(table[name]))
json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)
