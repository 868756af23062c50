"""Per-kernel table (launches, average / total duration, share) from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <command>`)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows:
    n, t = agg.get(r[k], (0, 0.0))
    agg[r[k]] = (n + 1, t + float(r[v].replace(",", "")) / 1e3)
total = sum(t for _, t in agg.values())
print("| kernel | launches | avg us | total us | share |\n|---|---|---|---|---|")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name[:110]}` | {n} | {t / n:.2f} | {t:.0f} | {100 * t / total:.1f}% |")
