#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 700 python -m pytest tests -q -m gpu > $O/k_pytest.txt 2>&1; echo "pytest rc $?"; tail -3 $O/k_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/k_bench.json 2> $O/k_bench.err; echo "bench rc $?"
head -c 300 $O/k_bench.json; echo; tail -3 $O/k_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/k_bench_reference.json 2> $O/k_bench_reference.err; echo "ref rc $?"; cat $O/k_bench_reference.json | head -c 400; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
  --log-file $O/k_launches_train.csv python tools/bench_train.py --model l --batch 4 --steps 1 --warmup 1 --eager > $O/k_ncu_train.log 2>&1; echo "ncu train rc $?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/k_smoke.txt 2>&1; echo "smoke rc $?"; tail -2 $O/k_smoke.txt
