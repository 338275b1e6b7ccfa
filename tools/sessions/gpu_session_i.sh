#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 700 python -m pytest tests -q -m gpu -x > $O/i_pytest.txt 2>&1; echo "pytest rc $?"; tail -3 $O/i_pytest.txt
timeout 200 python tools/ab_step.py l 8 "base" > $O/i_ab.txt 2>&1; tail -4 $O/i_ab.txt
timeout 300 python tools/layer_graph_bench.py l 8 > $O/i_layers.txt 2>&1; tail -1 $O/i_layers.txt
SY_TL_BN=1 timeout 100 python tools/conv_timeline.py 16 256 256 38 60 1 1 > $O/i_tl_1x1_256_38x60.txt 2>&1; head -6 $O/i_tl_1x1_256_38x60.txt
timeout 200 python tools/bench_train.py --model l --batch 4 --steps 20 --warmup 3 > $O/i_train.txt 2>&1
echo "train l b4: $(grep -o '"ms_per_step": [0-9.]*' $O/i_train.txt | tail -1)"
