"""Condense an `ncu --set full` report (exported with `ncu -i rep --page raw --csv`) to the handful of metrics the
roofline discussion uses.  usage: python tools/ncu_summary.py raw.csv > summary.txt"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("== %s  grid %s block %s" % (d.get("Kernel Name", "?")[:90], d.get("Grid Size", "?"), d.get("Block Size", "?")))
    for k in KEYS:
        if k in d:
            print("   %-75s %-14s %s" % (k, units[hdr.index(k)], d[k]))
