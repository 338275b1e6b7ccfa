cd /root/repo
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py -q --maxfail=30 -p no:cacheprovider -k "backward_chain" 2>&1 | tail -30) > gpurun_out/t_chain.txt
tail -30 gpurun_out/t_chain.txt | cut -c1-240
