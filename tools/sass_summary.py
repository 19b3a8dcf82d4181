"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st ->
LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, legacy mma.sync -> HMMA).   python tools/sass_summary.py [libsubgc_b200.so] > profiles/rNN_sass.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sub-gc_b200", "subgc", "libsubgc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "SYNCS", "HMMA", "REDG", "ATOMG",
        "LDGSTS", "MUFU"]
cur, counts, size = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        size[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        size[cur] += 1
        op = m.group(1)
        for w in want:
            if op.startswith(w):
                counts[cur][w] += 1
demangle = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS evidence per kernel (`cuobjdump -sass sub-gc_b200/subgc/libsubgc_b200.so`, sm_100a)\n")
print("tcgen05.mma -> `UTCHMMA` (kind::f16 / tf32), tcgen05.ld -> `LDTM`, cp.async.bulk(.tensor) -> `UBLKCP` / `UTMALDG` / `UTMASTG`, mbarrier -> `SYNCS`;")
print("`HMMA` (legacy mma.sync) must be absent.\n")
cols = [w for w in want if any(c[w] for c in counts.values())]
print("| kernel | instructions | " + " | ".join(cols) + " |")
print("|---|---:|" + "---:|" * len(cols))
for (name, c), dm in zip(counts.items(), demangle):
    short = re.sub(r"\(.*", "", dm).replace("subgc::", "")
    if not any(c[w] for w in cols if w not in ("MUFU", "REDG", "ATOMG")):
        continue
    print(f"| `{short[:60]}` | {size[name]} | " + " | ".join(str(c[w]) for w in cols) + " |")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("\nwhole library: " + ", ".join(f"{w} {tot[w]}" for w in want if tot[w]) + f"; HMMA {tot['HMMA']}")
