"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: share per kernel and the last decode step in order."""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    grid = r.get("Grid Size", "")
    rows.append((name, float(r["Metric Value"].replace(",", "")), grid))
tot = sum(t for _, t, _ in rows)
agg = collections.OrderedDict()
for n, t, _ in rows:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
print(f"{len(rows)} launches, {tot/1e3:.1f} us total")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100*t/tot:6.2f}%  {c:5d}  {t/c/1e3:8.2f} us  {n[:90]}")
# last full decode step: launches between the last two select_kernel launches
idx = [i for i, r in enumerate(rows) if "select_kernel" in r[0]]
if len(idx) >= 2:
    a, b = idx[-2] + 1, idx[-1] + 1
    print("one decode step:")
    s = 0.0
    for n, t, g in rows[a:b]:
        print(f"  {t/1e3:7.2f} us  {n.split('::')[-1][:40]:40s} {g}")
        s += t
    print(f"  {s/1e3:7.2f} us  sum")
