cd /root/repo
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py -q --maxfail=30 -p no:cacheprovider -k "wgrad or dgrad" 2>&1 | tail -40) > gpurun_out/t_bwd.txt
tail -30 gpurun_out/t_bwd.txt | cut -c1-220
