"""Summarise `ncu -i X.ncu-rep --page source --csv`: top SASS lines by stall samples, with the dominant stall reason."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ci = {c: i for i, c in enumerate(h)}
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[ci["# Samples"]] or 0) for r in rows[2:] if len(r) > ci["# Samples"])
execd = sum(int(r[ci["Instructions Executed"]] or 0) for r in rows[2:] if len(r) > ci["# Samples"])
print("total samples", tot, "warp instructions executed", execd)
agg = {}
for r in rows[2:]:
    if len(r) <= ci["# Samples"]:
        continue
    for s_ in stalls:
        agg[s_] = agg.get(s_, 0) + int(r[ci[s_]] or 0)
print("stall mix:", {k: round(100 * v / max(tot, 1), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
top = sorted([r for r in rows[2:] if len(r) > ci["# Samples"]], key=lambda r: -int(r[ci["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    n = int(r[ci["# Samples"]])
    why = max(stalls, key=lambda s_: int(r[ci[s_]] or 0))
    print(f"{100*n/tot:5.1f}%  exec {r[ci['Instructions Executed']]:>8}  {why:22s} {r[ci['Source']].strip()[:90]}")
