#!/bin/bash
# final validation of round 2: full GPU test suite, the bench line, the reference arm, the training-step launch list, smoke
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 700 python -m pytest tests -q -m gpu > $O/s_pytest.txt 2>&1; echo "pytest rc $?"; tail -3 $O/s_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/s_bench.json 2> $O/s_bench.err; echo "bench rc $?"
head -c 260 $O/s_bench.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/s_bench_reference.json 2> $O/s_bench_reference.err; echo "ref rc $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
  --log-file $O/s_launches_train.csv python tools/bench_train.py --model l --batch 4 --steps 1 --warmup 1 --eager > $O/s_ncu_train.log 2>&1; echo "ncu train rc $?"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv \
  --log-file $O/s_launches_step.csv python bench.py --steps 1 --warmup 3 --no-graph --no-train --no-extras --no-cpu-baseline > $O/s_ncu_step.log 2>&1; echo "ncu step rc $?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/s_smoke.txt 2>&1; echo "smoke rc $?"; tail -1 $O/s_smoke.txt
