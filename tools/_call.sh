cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1d.txt 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err
tail -4 gpurun_out/smoke_r1d.txt; python -c "
import json;d=json.load(open('gpurun_out/bench_r1d.json'));print(d['value'],d['ms_per_step'],d['clocks'],d['e2e']['value'],d['roofline']['achieved'],d['cpu_baseline'])"; tail -3 gpurun_out/bench_r1d.err
