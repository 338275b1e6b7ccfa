cd /root/repo
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_postprocess.py -q -p no:cacheprovider 2>&1 | tail -25) > gpurun_out/t_nms.txt
(timeout 300 python tools/wgrad_bench.py 2>&1 | tail -10) > gpurun_out/wgrad_bench.txt
tail -12 gpurun_out/t_nms.txt | cut -c1-200; cat gpurun_out/wgrad_bench.txt
