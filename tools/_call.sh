cd /root/repo
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py -q --maxfail=60 -p no:cacheprovider 2>&1 | tail -30) > gpurun_out/t_ops.txt
(timeout 420 python -m pytest tests/test_gpu_model.py -q -x -p no:cacheprovider 2>&1 | tail -40) > gpurun_out/t_model.txt
(SY_CONV_A=halo timeout 420 python -m pytest tests/test_gpu_model.py -q -x -p no:cacheprovider 2>&1 | tail -40) > gpurun_out/t_model_halo.txt
(timeout 300 python tools/ab_step.py l 8 "base,halo" 2>&1 | tail -30) > gpurun_out/ab7.txt
tail -3 gpurun_out/t_ops.txt; tail -3 gpurun_out/t_model.txt; tail -3 gpurun_out/t_model_halo.txt; cat gpurun_out/ab7.txt
