cd /root/repo
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py -q --maxfail=30 -p no:cacheprovider -k "stride2 or upsample_nearest_backward or head_pred_backward" 2>&1 | tail -40) > gpurun_out/t_bwd2.txt
tail -30 gpurun_out/t_bwd2.txt | cut -c1-240
