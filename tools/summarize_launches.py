"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST forward."""
import collections
import csv
import re
import sys

path = sys.argv[1]
n_last = int(sys.argv[2]) if len(sys.argv) > 2 else 399
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
recs = [x for x in csv.DictReader(lines) if x["Metric Name"] == "gpu__time_duration.sum"]
last = recs[-n_last:]
tot = sum(float(x["Metric Value"].replace(",", "")) for x in last)
agg = collections.defaultdict(lambda: [0, 0.0])
for x in last:
    name = re.sub(r"\(.*", "", x["Kernel Name"])
    name = name.replace("void ", "").replace("sta::", "").replace("<unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += float(x["Metric Value"].replace(",", ""))
print("timed forward: %d launches, %.3f ms (serialised, cold caches, clocks not locked)" % (len(last), tot / 1e6))
print("%10s %6s %6s  %s" % ("ms", "share", "count", "kernel"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%10.3f %5.1f%% %6d  %s" % (v[1] / 1e6, 100 * v[1] / tot, v[0], k))
