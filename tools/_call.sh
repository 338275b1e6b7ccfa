cd /root/repo
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8) > gpurun_out/pytest_gpu_r1c.txt
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 600 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
bash tools/ncu_one.sh conv_tc_kernel 454 r01c_conv_head3x3_full
python tools/bench_modes.py l > gpurun_out/bench_modes_r1c.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r1c.txt 2>&1
tail -3 gpurun_out/pytest_gpu_r1c.txt; cat gpurun_out/bench_r1c.json; tail -3 gpurun_out/bench_modes_r1c.txt; tail -2 gpurun_out/smoke_r1c.txt
