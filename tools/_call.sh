cd /root/repo
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu_r1f.json 2> gpurun_out/bench_2gpu_r1f.err
echo "stdout lines: $(wc -l < gpurun_out/bench_2gpu_r1f.json)"; head -c 200 gpurun_out/bench_2gpu_r1f.json; echo; ls /tmp/sy_nccl* 2>/dev/null | head -3
