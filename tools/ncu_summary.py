"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST training step
(delimited by the adamw_clip_kernel launches)."""
import csv, sys, re, collections
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    rows.append((row["Kernel Name"], ns))
ends = [i for i, (k, _) in enumerate(rows) if "adamw_clip" in k]
if len(ends) >= 2:
    step = rows[ends[-2] + 1: ends[-1] + 1]
else:
    step = rows
tot = sum(ns for _, ns in step)
agg = collections.defaultdict(lambda: [0, 0.0])
for k, ns in step:
    k = re.sub(r"\(.*", "", k)
    agg[k][0] += 1
    agg[k][1] += ns
print(f"launches in step: {len(step)}  total kernel time: {tot/1e6:.2f} ms (serialised, cold cache)")
print(f"{'kernel':70s} {'n':>5s} {'ms':>9s} {'share':>7s}")
for k, (n, ns) in sorted(agg.items(), key=lambda t: -t[1][1])[:40]:
    print(f"{k[:70]:70s} {n:5d} {ns/1e6:9.3f} {100*ns/tot:6.1f}%")
