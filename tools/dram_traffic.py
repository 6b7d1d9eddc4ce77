"""From an ncu CSV with dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum per launch, write the
per-launch DRAM traffic of the conv3_tc_kernel launches of the LAST forward as JSON (read by bench.py for
roofline.traffic).  usage: python tools/dram_traffic.py metrics.csv n_last out.json"""
import csv
import json
import sys

path, n_last, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
lines = [l for l in open(path) if not l.startswith("==")]
per = {}
order = []
for row in csv.DictReader(lines):
    i = row["ID"]
    if i not in per:
        per[i] = {"kernel": row["Kernel Name"].split("(")[0].replace("void ", "").replace("dwmh::", ""), "grid": row["Grid Size"]}
        order.append(i)
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    name = row["Metric Name"]
    if name.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    elif name.startswith("gpu__time"):
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)          # -> ms
    per[i][name] = v
sel = [per[i] for i in order[-n_last:]]
tc = [x for x in sel if x["kernel"].startswith("conv3_tc_kernel")]
tot = sum(x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0) for x in tc)
res = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, tests/tools/profile_forward.py 32 1 (one batch of 32 tile-forwards)",
       "tc_launches": len(tc), "tc_dram_bytes_total": tot, "tc_dram_bytes_per_launch": tot / max(len(tc), 1),
       "all_dram_bytes_total": sum(x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0) for x in sel),
       "launches": [{"kernel": x["kernel"][:60], "grid": x["grid"], "ms": round(x.get("gpu__time_duration.sum", 0), 4),
                     "dram_read_GB": round(x.get("dram__bytes_read.sum", 0) / 1e9, 4), "dram_write_GB": round(x.get("dram__bytes_write.sum", 0) / 1e9, 4)} for x in sel]}
json.dump(res, open(out, "w"), indent=1)
print("tc launches %d, DRAM %.2f GB total, %.3f GB per launch" % (len(tc), tot / 1e9, tot / 1e9 / max(len(tc), 1)))
