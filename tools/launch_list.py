"""Print the per-launch durations of the LAST forward in an ncu launch list (gpu__time_duration.sum CSV).
usage: python tools/launch_list.py launches.csv [n_last]"""
import csv
import sys

path = sys.argv[1]
n_last = int(sys.argv[2]) if len(sys.argv) > 2 else 47
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    ms = v / 1e6 if row["Metric Unit"] == "ns" else (v / 1e3 if row["Metric Unit"] == "us" else v)
    rows.append((row["Kernel Name"][:58], row["Grid Size"], ms))
for k in rows[-n_last:]:
    print("%-60s %-18s %.3f" % k)
print("total %.3f ms" % sum(k[2] for k in rows[-n_last:]))
