cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1e.txt 2>&1
tail -5 gpurun_out/smoke_r1e.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_2gpu_r1e.json 2> gpurun_out/bench_2gpu_r1e.err
cat gpurun_out/bench_2gpu_r1e.json | cut -c1-400; tail -3 gpurun_out/bench_2gpu_r1e.err
