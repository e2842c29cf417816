"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = {}
for r in rows[1:]:
    k = r[ki].split("(")[0]; agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-28s n=%5d total %10.1f us  avg %9.2f us  %5.1f%%" % (k, v[0], v[1] / 1e3, v[1] / 1e3 / v[0], 100 * v[1] / tot))
