#!/bin/bash
# round-2 session B: validate the loss / head / pack / BatchNorm-backward / first-write changes and measure them
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 700 python -m pytest tests -q -m gpu -x > $O/b_pytest.txt 2>&1; echo "pytest rc $?" | tee -a $O/b_pytest.txt
tail -5 $O/b_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/b_bench.json 2> $O/b_bench.err; echo "bench rc $?"
head -c 400 $O/b_bench.json; echo
timeout 200 python tools/ab_step.py l 8 "base,head pred" > $O/b_ab.txt 2>&1; echo "ab rc $?"
cat $O/b_ab.txt | tail -6
for w in 2 1; do
  SY_WGRAD_WAVES=$w timeout 200 python tools/bench_train.py --model l --batch 4 --steps 20 --warmup 3 > $O/b_train_waves$w.txt 2>&1
  echo "waves $w: $(grep -o "\"ms_per_step\": [0-9.]*" $O/b_train_waves$w.txt | tail -1)"
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
  --log-file $O/b_launches_train.csv python tools/bench_train.py --model l --batch 4 --steps 1 --warmup 1 --eager > $O/b_ncu_train.log 2>&1; echo "ncu train rc $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
  --log-file $O/b_launches_step.csv python bench.py --steps 1 --warmup 3 --no-graph --no-train --no-extras --no-cpu-baseline > $O/b_ncu_step.log 2>&1; echo "ncu step rc $?"
