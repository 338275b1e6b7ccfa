#!/bin/bash
# usage: tools/ncu_full.sh <label> <kernel-regex> <skip> -- <ncu_layer.py args...>
# One `ncu --set full` capture of one launch of tools/ncu_layer.py (cold cache, serialised) and a text summary of the metrics
# DESIGN.md quotes; the .ncu-rep stays in gpurun_out/ for `ncu -i ... --page source`.
label=$1; regex=$2; skip=$3; shift 4
O=gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c 1 -o $O/$label -f \
    python tools/ncu_layer.py "$@" > $O/$label.log 2>&1
echo "== $label   (python tools/ncu_layer.py $*)"
ncu -i $O/$label.ncu-rep --page raw --csv 2>/dev/null | python -c '
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print(f"{w:82s} {vals[i][:120]} {units[i]}")
'
