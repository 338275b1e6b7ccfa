#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_train.py tests/test_gpu_ops.py -q -m gpu -x > $O/c_pytest.txt 2>&1; echo "pytest rc $?"; tail -2 $O/c_pytest.txt
timeout 200 python tools/bench_train.py --model l --batch 4 --steps 20 --warmup 3 > $O/c_train_l_b4.txt 2>&1
echo "train l b4: $(grep -o '"ms_per_step": [0-9.]*' $O/c_train_l_b4.txt | tail -1)"
{
echo "# ncu --set full --clock-control none --import-source on, one launch each (third launch of tools/ncu_layer.py; cold cache, serialised), B200, round 2 (final kernels: cta_group::2 pair mode)"
timeout 200 tools/ncu_full.sh c_head3x3_pair conv_tc_kernel 2 -- 8 256 256 75 120 3 1
timeout 200 tools/ncu_full.sh c_c3x3_128_halo_pair conv_tc_kernel 2 -- 16 128 128 75 120 3 1
timeout 200 tools/ncu_full.sh c_c3x3_256_38x60_pair conv_tc_kernel 2 -- 16 256 256 38 60 3 1
timeout 200 tools/ncu_full.sh c_c1x1_128 conv_tc_kernel 2 -- 16 128 128 75 120 1 1
timeout 200 tools/ncu_full.sh c_apply bn_act_apply 2 -- 16 128 128 75 120 1 1
} > $O/c_ncu_full_summary.txt 2>&1
grep -E "^==|tensor_cycles|time_duration" $O/c_ncu_full_summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
  --log-file $O/c_launches_train.csv python tools/bench_train.py --model l --batch 4 --steps 1 --warmup 1 --eager > $O/c_ncu_train.log 2>&1; echo "ncu train rc $?"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/c_bench.json 2> $O/c_bench.err; echo "bench rc $?"
head -c 300 $O/c_bench.json; echo
ls -la $O/*.ncu-rep
