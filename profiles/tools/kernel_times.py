"""Per-kernel average duration from an `ncu --metrics gpu__time_duration.sum --csv` launch list: python kernel_times.py launches.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
ki, vi = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
d = collections.defaultdict(list)
for r in rows[h + 1:]:
    try:
        d[r[ki]].append(float(r[vi].replace(",", "")))
    except (ValueError, IndexError):
        pass
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{len(v):5d} x {sum(v) / len(v) / 1e3:9.1f} us  {100 * sum(v) / tot:5.1f} %  {k[:110]}")
