#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train.py -q -m gpu -x -k "wgrad or train or walk or pack or bn_act_backward or backward" > $O/g_pytest.txt 2>&1; echo "pytest rc $?"; tail -2 $O/g_pytest.txt
for cfg in "default" "SY_WGRAD_SG1=1" "SY_PDL=0"; do
  if [ "$cfg" = "default" ]; then envs=""; else envs="$cfg"; fi
  env $envs timeout 200 python tools/bench_train.py --model l --batch 4 --steps 20 --warmup 3 > $O/g_train.txt 2>&1
  echo "train l b4 [$cfg]: $(grep -o '"ms_per_step": [0-9.]*' $O/g_train.txt | tail -1)" | tee -a $O/g_train_variants.txt
done
