#!/bin/bash
# round-2 re-entry: validate HEAD on a B200 and collect fresh profiles (see profiles/README.md)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/a_smi.txt 2>&1
timeout 600 python -m pytest tests -q -m gpu -x > $O/a_pytest.txt 2>&1; echo "pytest rc $?" | tee -a $O/a_pytest.txt
tail -3 $O/a_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/a_bench.json 2> $O/a_bench.err; echo "bench rc $?"
head -c 600 $O/a_bench.json; echo
timeout 300 python tools/ab_step.py l 8 "base,fuse apply,skip every" > $O/a_ab.txt 2>&1; echo "ab rc $?"
cat $O/a_ab.txt | tail -12
timeout 300 python tools/layer_graph_bench.py l 8 > $O/a_layers.txt 2>&1; echo "layers rc $?"
tail -2 $O/a_layers.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv \
  --log-file $O/a_launches_step.csv python bench.py --steps 1 --warmup 3 --no-graph --no-train --no-extras --no-cpu-baseline > $O/a_ncu_step.log 2>&1; echo "ncu step rc $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
  --log-file $O/a_launches_train.csv python tools/bench_train.py --model l --batch 4 --steps 1 --warmup 1 --eager > $O/a_ncu_train.log 2>&1; echo "ncu train rc $?"
ls -la $O | head -30
