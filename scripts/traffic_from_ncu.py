#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report of this build: DRAM bytes read / written per launch of every kernel.

    python scripts/traffic_from_ncu.py gpurun_out/prof_<tag>.ncu-rep cfg2_sorted_channels_last [profiles/traffic.json]
Runs in the build container (ncu -i needs no GPU).  bench.py copies the entry of its workload into `roofline.traffic`."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

rep, key = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.OrderedDict()
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
    rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale[units[col["dram__bytes_read.sum"]]]
    wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale[units[col["dram__bytes_write.sum"]]]
    dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
    du = units[col["gpu__time_duration.sum"]]
    dur = dur / 1e3 if du in ("ns", "nsecond") else dur
    agg.setdefault(name, []).append((rd, wr, dur))
entry = {}
for name, v in agg.items():
    n = len(v)
    entry[name] = {"launches": n, "dram_read_bytes": round(sum(x[0] for x in v) / n), "dram_write_bytes": round(sum(x[1] for x in v) / n),
                   "duration_us_under_ncu": round(sum(x[2] for x in v) / n, 2)}
fwd = [k for k in entry if any(t in k for t in ("zero_flags", "prologue", "fwd_columns", "fwd_gather", "fwd_store", "lift"))]
entry["forward_op_total"] = sum(entry[k]["dram_read_bytes"] + entry[k]["dram_write_bytes"] for k in fwd)
entry["source"] = os.path.basename(rep)
try:
    allj = json.load(open(out))
except Exception:
    allj = {}
allj[key] = entry
json.dump(allj, open(out, "w"), indent=1)
print(json.dumps(entry, indent=1))
