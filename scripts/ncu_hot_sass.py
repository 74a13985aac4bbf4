"""Summarise `ncu -i X.ncu-rep --page source --csv`: per kernel section, top SASS lines by stall samples with the
dominant stall reason.  usage: ncu_hot_sass.py file.csv [top_n] [kernel-name substring]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else ""
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
for sec in secs:
    if want not in sec["name"] or sec["name"] in seen:
        continue
    seen.add(sec["name"])
    h = sec["rows"][0]
    ci = {c: i for i, c in enumerate(h)}
    body = [r for r in sec["rows"][1:] if len(r) > ci["# Samples"] and r[ci["# Samples"]].isdigit()]
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[ci["# Samples"]]) for r in body)
    execd = sum(int(r[ci["Instructions Executed"]] or 0) for r in body)
    print("==", sec["name"][:100])
    print("total samples", tot, "warp instructions executed", execd, "sass lines", len(body))
    agg = {s_: sum(int(r[ci[s_]] or 0) for r in body) for s_ in stalls}
    print("stall mix:", {k: round(100 * v / max(tot, 1), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for r in sorted(body, key=lambda r: -int(r[ci["# Samples"]]))[:top_n]:
        n = int(r[ci["# Samples"]])
        why = max(stalls, key=lambda s_: int(r[ci[s_]] or 0))
        print(f"{100*n/max(tot,1):5.1f}%  exec {r[ci['Instructions Executed']]:>8}  {why:22s} {r[ci['Source']].strip()[:90]}")
