"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = [(int(r["ID"]), re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("vv::", ""), float(r["Metric Value"].replace(",", "")) / 1e3,
         r["Grid Size"]) for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
tot = sum(r[2] for r in rows)
print("launches: %d   total kernel time: %.1f us" % (len(rows), tot))
agg = collections.defaultdict(lambda: [0, 0.0])
for _, k, v, _ in rows:
    agg[k][0] += 1
    agg[k][1] += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%10.1f us %5.1f%%  n=%4d avg=%7.2f us  %s" % (t, 100 * t / tot, n, t / n, k))
if len(sys.argv) > 2:
    for r in rows:
        print(r)
