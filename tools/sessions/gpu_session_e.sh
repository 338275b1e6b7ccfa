#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train.py -q -m gpu -x -k "wgrad or train or walk or pack" > $O/e_pytest.txt 2>&1; echo "pytest rc $?"; tail -2 $O/e_pytest.txt
timeout 200 python tools/bench_train.py --model l --batch 4 --steps 20 --warmup 3 > $O/e_train_l_b4.txt 2>&1
echo "train l b4: $(grep -o '"ms_per_step": [0-9.]*' $O/e_train_l_b4.txt | tail -1)"
timeout 200 python tools/bench_train.py --model l --batch 8 --steps 10 --warmup 3 > $O/e_train_l_b8.txt 2>&1
echo "train l b8: $(grep -o '"ms_per_step": [0-9.]*' $O/e_train_l_b8.txt | tail -1)"
timeout 200 python tools/wgrad_bench.py > $O/e_wgrad_bench.txt 2>&1; tail -12 $O/e_wgrad_bench.txt
