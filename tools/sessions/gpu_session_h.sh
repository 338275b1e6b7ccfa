#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_train.py -q -m gpu -x > $O/h_pytest.txt 2>&1; echo "pytest rc $?"; tail -2 $O/h_pytest.txt
timeout 200 python tools/bench_train.py --model l --batch 4 --steps 20 --warmup 3 > $O/h_train.txt 2>&1
echo "train l b4: $(grep -o '"ms_per_step": [0-9.]*' $O/h_train.txt | tail -1)"
timeout 200 python tools/ab_step.py l 8 "base,no pdl" > $O/h_ab.txt 2>&1; tail -5 $O/h_ab.txt
SY_TL_BN=1 timeout 100 python tools/conv_timeline.py 16 256 256 38 60 1 1 > $O/h_tl_1x1_256_38x60.txt 2>&1; head -8 $O/h_tl_1x1_256_38x60.txt
SY_TL_BN=1 timeout 100 python tools/conv_timeline.py 16 512 512 19 30 1 1 > $O/h_tl_1x1_512_19x30.txt 2>&1; head -8 $O/h_tl_1x1_512_19x30.txt
SY_TL_BN=1 timeout 100 python tools/conv_timeline.py 8 256 256 19 30 3 1 > $O/h_tl_3x3_256_8x19x30.txt 2>&1; head -8 $O/h_tl_3x3_256_8x19x30.txt
