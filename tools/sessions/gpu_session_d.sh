#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_train.py -q -m gpu -x > $O/d_pytest.txt 2>&1; echo "pytest rc $?"; tail -2 $O/d_pytest.txt
for v in 00 11 22 33 10 20 01 02; do
  SY_BNBWD=$v timeout 200 python tools/bench_train.py --model l --batch 4 --steps 20 --warmup 3 > $O/d_train_$v.txt 2>&1
  echo "SY_BNBWD=$v: $(grep -o '"ms_per_step": [0-9.]*' $O/d_train_$v.txt | tail -1)" | tee -a $O/d_bnbwd_variants.txt
done
