"""Summarise an `ncu --page source --csv --print-source sass` export of a warp-specialised kernel: warp-stall samples grouped by how often
an instruction ran (the roles of conv3_tc_kernel differ in that: loaders, epilogue warps, issuers, spin loops), then the hottest
instructions.  usage: python tools/ncu_src_roles.py file.csv [n_ctas*warps_per_cta] [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
unit = float(sys.argv[2]) if len(sys.argv) > 2 else 32768.0
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]


def f(r, k):
    try:
        return float(r[idx[k]])
    except (ValueError, KeyError, IndexError):
        return 0.0


stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(f(r, "# Samples") for r in data)
print("kernel:", rows[0][1][:110])
print("samples %d over %d instructions" % (tot, len(data)))
g = defaultdict(lambda: [0.0, 0, defaultdict(float)])
for r in data:
    key = round(f(r, "Instructions Executed") / unit, 1)
    g[key][0] += f(r, "# Samples")
    g[key][1] += 1
    for s in stalls:
        g[key][2][s] += f(r, s)
for k, v in sorted(g.items(), key=lambda kv: -kv[1][0])[:12]:
    top = sorted(v[2].items(), key=lambda kv: -kv[1])[:5]
    print("executed %8.1f x  instrs %5d  samples %8.0f (%4.1f%%)  " % (k, v[1], v[0], 100 * v[0] / tot), " ".join("%s %.0f" % (a[6:], b) for a, b in top))
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:ntop]:
    best = max(stalls, key=lambda s: f(r, s))
    print("%6s %-64s smp %6.0f exec/unit %8.1f  %s %.0f" % (r[idx["Address"]][-5:], r[idx["Source"]][:64], f(r, "# Samples"), f(r, "Instructions Executed") / unit, best[6:], f(r, best)))
