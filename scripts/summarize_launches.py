"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean / min / max (us)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v = v / 1e3 if r[mu] in ("ns", "nsecond") else v
    agg.setdefault(r[kn].split("(")[0][:70], []).append(v)
print("| kernel | launches | mean us | min us | max us |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"| `{k}` | {len(v)} | {sum(v)/len(v):.2f} | {min(v):.2f} | {max(v):.2f} |")
