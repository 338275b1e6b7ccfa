cd /root/repo
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py -q --maxfail=30 -p no:cacheprovider -k "backward_chain or add_and_spp or bn_finalize" 2>&1 | tail -40) > gpurun_out/t_bwd3.txt
tail -30 gpurun_out/t_bwd3.txt | cut -c1-240
