"""Turns the ncu launch list of one eager training step (scripts/ncu_round.sh) into
  profiles/<tag>_gemm_traffic.json   {"B": {"gemm_dram_bytes_per_step": ..., "source": ...}}   (read by bench.py)
  profiles/<tag>_launch_shares.md    per-kernel launches / time / share of the step / DRAM bytes
Usage: python scripts/ncu_traffic.py gpurun_out/r02_launches.csv r02 [config]"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

path, tag = sys.argv[1], sys.argv[2]
cfg = sys.argv[3] if len(sys.argv) > 3 else "B"
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
per = defaultdict(dict)
for r in rd:
    per[int(r["ID"])]["name"] = r["Kernel Name"]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    per[int(r["ID"])][r["Metric Name"]] = v * scale
ids = sorted(per)
# the bench runs 3 identical eager steps (warm-up, timed, e2e): take the middle third of the launches between the first
# and the last optimizer kernel
adam = [i for i in ids if "adamw_multi" in per[i]["name"]]
if len(adam) >= 3:
    lo, hi = adam[0] + 1, adam[1]
    step_ids = [i for i in ids if lo <= i <= hi]
else:
    step_ids = ids
agg = defaultdict(lambda: [0, 0.0, 0.0])
for i in step_ids:
    d = per[i]
    n = re.sub(r"\(.*", "", d["name"])[:80]
    a = agg[n]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
gemm_bytes = sum(a[2] for n, a in agg.items() if "gemm_tc" in n)
gemm_us = sum(a[1] for n, a in agg.items() if "gemm_tc" in n)
os.makedirs("profiles", exist_ok=True)
out = {}
jp = os.path.join("profiles", f"{tag}_gemm_traffic.json")
if os.path.isfile(jp):
    out = json.load(open(jp))
out[cfg] = {"gemm_dram_bytes_per_step": gemm_bytes, "gemm_us_under_ncu": gemm_us, "step_us_under_ncu": tot,
            "gemm_share_of_step_under_ncu": gemm_us / tot if tot else None,
            "source": f"ncu launch list profiles/{tag}_launches.csv (dram__bytes_read.sum + dram__bytes_write.sum over the "
                      f"gemm_tc launches of one eager step, scripts/ncu_traffic.py)"}
json.dump(out, open(jp, "w"), indent=1)
with open(os.path.join("profiles", f"{tag}_launch_shares.md"), "w") as f:
    f.write(f"# {tag}: one eager training step of config {cfg} under ncu (serialised, cold caches: SHARES are meaningful)\n\n")
    f.write("| kernel | launches | us | share | DRAM MB |\n|---|---|---|---|---|\n")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{n}` | {a[0]} | {a[1]:.0f} | {a[1] / tot:.3f} | {a[2] / 1e6:.1f} |\n")
    f.write(f"\ntotal {tot:.0f} us over {len(step_ids)} launches; GEMM launches {gemm_us:.0f} us ({gemm_us / tot:.3f}), "
            f"GEMM DRAM traffic {gemm_bytes / 1e9:.2f} GB\n")
print(json.dumps(out[cfg], indent=1))
