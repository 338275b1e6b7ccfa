cd /root/repo
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_model.py -q -p no:cacheprovider -k "loss" 2>&1 | tail -25) > gpurun_out/t_lossbwd.txt
tail -25 gpurun_out/t_lossbwd.txt | cut -c1-220
