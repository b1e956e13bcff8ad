"""Kernel table (launches, avg us, share of the step) from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
t = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    t[name][0] += 1
    t[name][1] += float(r[14]) / 1e3
tot = sum(v[1] for v in t.values())
print("| kernel | launches | avg us (ncu, cold, serialised) | share |\n|---|---|---|---|")
for k, (n, us) in sorted(t.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {n} | {us / n:.2f} | {100 * us / tot:.1f}% |")
