"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST
forward in the log (or all launches with --all).  usage: python tools/launch_summary.py file.csv [n_last]"""
import collections
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
n_last = int(sys.argv[2]) if len(sys.argv) > 2 else len(rows) // 2
sel = rows[-n_last:]
tot = collections.OrderedDict()
for i, x in enumerate(sel):
    v = float(x["Metric Value"].replace(",", ""))
    us = v / 1000 if x["Metric Unit"].startswith("ns") else v
    k = x["Kernel Name"].split("(")[0].replace("void ", "").replace("dwmh::", "")[:48]
    tot.setdefault(k, [0.0, 0])
    tot[k][0] += us; tot[k][1] += 1
    if "-v" in sys.argv:
        print("%3d %-48s %10.1f us  grid %s" % (i, k, us, x["Grid Size"]))
s = sum(v[0] for v in tot.values())
for k, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print("%-50s %4d launches %10.3f ms  %5.1f %%" % (k, n, us / 1000, 100 * us / s))
print("total %.3f ms over %d launches" % (s / 1000, len(sel)))
