"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`): share per
kernel, DRAM bytes per kernel, and the last decode step in launch order."""
import csv, sys, collections, json
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
per = collections.OrderedDict()   # launch id -> dict
for r in csv.DictReader(lines):
    i = r["ID"]
    d = per.setdefault(i, {"name": r["Kernel Name"].split("(")[0], "grid": r.get("Grid Size", ""), "t": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        d["t"] = v * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(unit, 1.0)
    elif m == "dram__bytes_read.sum":
        d["rd"] = v * scale
    elif m == "dram__bytes_write.sum":
        d["wr"] = v * scale
rows = list(per.values())
tot = sum(r["t"] for r in rows)
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r["name"], [0, 0.0, 0.0]); a[0] += 1; a[1] += r["t"]; a[2] += r["rd"] + r["wr"]
print(f"{len(rows)} launches, {tot/1e3:.1f} us total (cold-cache, serialised: compare shares)")
print("| share | launches | avg us | avg DRAM MB | kernel |\n|---:|---:|---:|---:|---|")
for n, (c, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {100*t/tot:.2f}% | {c} | {t/c/1e3:.2f} | {b/c/1e6:.2f} | `{n[:80]}` |")
idx = [i for i, r in enumerate(rows) if "select" in r["name"]]
if len(idx) >= 2:
    a, b = idx[-2] + 1, idx[-1] + 1
    print("\none decode step (launch order):\n| us | DRAM MB | kernel | grid |\n|---:|---:|---|---|")
    st = sb = 0.0
    for r in rows[a:b]:
        print(f"| {r['t']/1e3:.2f} | {(r['rd']+r['wr'])/1e6:.2f} | `{r['name'].split('::')[-1][:40]}` | {r['grid']} |")
        st += r["t"]; sb += r["rd"] + r["wr"]
    print(f"| **{st/1e3:.2f}** | **{sb/1e6:.2f}** | sum | |")
    if len(sys.argv) > 2:
        json.dump({"dram_bytes_per_decode_loop": 20 * sb, "dram_bytes_per_step": sb,
                   "note": "dram__bytes_read+write summed over the launches of one decode step (ncu, cold-cache replay, graphs off) x 20 steps"},
                  open(sys.argv[2], "w"))
mega = [r for r in rows if "mega_decode_kernel" in r["name"]]
if mega and len(sys.argv) > 2:   # persistent decode kernel: one launch IS the decode loop of a call
    b = sum(r["rd"] + r["wr"] for r in mega) / len(mega)
    print(f"\npersistent decode kernel: {len(mega)} launches, avg {sum(r['t'] for r in mega) / len(mega) / 1e3:.1f} us, avg DRAM {b / 1e6:.1f} MB per launch")
    json.dump({"dram_bytes_per_decode_loop": b, "dram_bytes_per_step": b / 20,
               "note": "dram__bytes_read+write of one mega_decode_kernel launch (= the 20-step decode loop of one call), ncu launch list, "
                       "average over the captured launches"}, open(sys.argv[2], "w"))
