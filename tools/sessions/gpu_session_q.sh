#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
timeout 700 python -m pytest tests -q -m gpu > $O/q_pytest.txt 2>&1; echo "pytest rc $?"; tail -3 $O/q_pytest.txt
timeout 200 python tools/ab_step.py l 8 "base,epi cycles 1900,1900,1450" > $O/q_ab.txt 2>&1; tail -5 $O/q_ab.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/q_bench.json 2> $O/q_bench.err; echo "bench rc $?"
head -c 260 $O/q_bench.json; echo
timeout 300 python tools/layer_graph_bench.py l 8 > $O/q_layers.txt 2>&1; tail -1 $O/q_layers.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/q_smoke.txt 2>&1; echo "smoke rc $?"; tail -1 $O/q_smoke.txt
