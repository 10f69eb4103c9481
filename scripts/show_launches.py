"""Per-kernel summary of an ncu launch list (gpu__time_duration.sum CSV): python scripts/show_launches.py file.csv ..."""
import collections, csv, sys
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    ki, mi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mi:
            continue
        n = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        agg.setdefault(n, []).append(float(r[mi]))
    print(f)
    for k, v in agg.items():
        print(f"  {k:40s} n={len(v):3d} avg {sum(v)/len(v)/1e3:8.1f} us  min {min(v)/1e3:8.1f}  max {max(v)/1e3:8.1f}")
