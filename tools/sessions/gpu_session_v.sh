#!/bin/bash
# ncu --set full captures of the backward bricks and the HBM-bound glue kernels (evidence for DESIGN section 4)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
full() {  # label regex skip -- args
  label=$1; regex=$2; skip=$3; shift 4
  ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c 1 -o $O/$label -f python tools/ncu_backward.py "$@" > $O/$label.log 2>&1
  echo "== $label   (python tools/ncu_backward.py $*)"
  ncu -i $O/$label.ncu-rep --page raw --csv 2>/dev/null | python -c '
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print(f"{w:82s} {vals[i][:110]} {units[i]}")
'
}
{
echo "# ncu --set full --clock-control none --import-source on, one launch each (third launch of tools/ncu_backward.py; cold cache, serialised), B200, round 2"
full v_wgrad_3x3_256 conv_wgrad_kernel 2 -- 16 256 256 38 60 3 1
full v_wgrad_reduce wgrad_reduce_kernel 2 -- 16 256 256 38 60 3 1
full v_bn_bwd_reduce bn_act_bwd_reduce 2 -- 16 128 128 75 120 1 1
full v_bn_bwd_apply bn_act_bwd_apply 2 -- 16 128 128 75 120 1 1
full v_focus_pack focus_pack_kernel 2 -- 16 128 128 75 120 1 1
full v_head_pred head_pred_kernel 2 -- 16 128 128 75 120 1 1
} > $O/v_ncu_full_summary.txt 2>&1
grep -E "^==|time_duration|dram_throughput" $O/v_ncu_full_summary.txt
rm -f $O/v_*.ncu-rep
