#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 700 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pair.py tests/test_gpu_parity_l.py tests/test_gpu_model.py -q -m gpu -x > $O/j_pytest.txt 2>&1; echo "pytest rc $?"; tail -3 $O/j_pytest.txt
timeout 200 python tools/ab_step.py l 8 "base" > $O/j_ab.txt 2>&1; tail -4 $O/j_ab.txt
SY_TL_BN=1 timeout 100 python tools/conv_timeline.py 16 256 256 38 60 1 1 > $O/j_tl_1x1_256_38x60.txt 2>&1; head -6 $O/j_tl_1x1_256_38x60.txt; grep " K " $O/j_tl_1x1_256_38x60.txt
timeout 300 python tools/layer_graph_bench.py l 8 > $O/j_layers.txt 2>&1; tail -1 $O/j_layers.txt
