cd /root/repo
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py -q --maxfail=60 -p no:cacheprovider 2>&1 | tail -40) > gpurun_out/t_ops.txt
(SY_STAGE_TILES=2 timeout 300 python -m pytest tests/test_gpu_ops.py -q --maxfail=60 -p no:cacheprovider -k "conv" 2>&1 | tail -20) > gpurun_out/t_ops_st2.txt
(SY_STAGE_TILES=1 timeout 300 python -m pytest tests/test_gpu_ops.py -q --maxfail=60 -p no:cacheprovider -k "conv" 2>&1 | tail -20) > gpurun_out/t_ops_st1.txt
(timeout 420 python -m pytest tests/test_gpu_model.py -q -x -p no:cacheprovider 2>&1 | tail -40) > gpurun_out/t_model.txt
(timeout 300 python tools/ab_step.py l 8 "base,staging" 2>&1 | tail -30) > gpurun_out/ab4.txt
(timeout 120 python tools/conv_timeline.py 16 128 128 75 120 1 1 2>&1 | head -120) > gpurun_out/tl3_1x1.txt
tail -3 gpurun_out/t_ops.txt; tail -3 gpurun_out/t_ops_st2.txt; tail -3 gpurun_out/t_ops_st1.txt; tail -3 gpurun_out/t_model.txt; cat gpurun_out/ab4.txt; head -6 gpurun_out/tl3_1x1.txt
