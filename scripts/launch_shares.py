"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols = rows[hdr]; ki = cols.index("Kernel Name"); vi = cols.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg[r[ki].split("(")[0][:60]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>10s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:60s} {len(v):8d} {sum(v) / 1e6:10.3f} {sum(v) / tot * 100:6.1f}% {sum(v) / len(v) / 1e3:10.1f}")
