#!/bin/bash
# usage: tools/ncu_one.sh <kernel-regex> <skip (matching launches)> <out-name>
ncu --set full --clock-control none --import-source on -k "regex:$1" -s $2 -c 1 -o gpurun_out/$3 -f \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/$3.log 2>&1
grep -E "PROF|WARN" gpurun_out/$3.log | head -5
