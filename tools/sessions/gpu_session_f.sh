#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_train.py tests/test_gpu_model.py -q -m gpu -x > $O/f_pytest.txt 2>&1; echo "pytest rc $?"; tail -2 $O/f_pytest.txt
timeout 200 python tools/bench_train.py --model l --batch 4 --steps 20 --warmup 3 > $O/f_train_l_b4.txt 2>&1
echo "train l b4: $(grep -o '"ms_per_step": [0-9.]*' $O/f_train_l_b4.txt | tail -1)"
timeout 400 python tools/ab_step.py l 8 "base,apply:,raw arena" > $O/f_ab.txt 2>&1; echo "ab rc $?"
cat $O/f_ab.txt | tail -14
