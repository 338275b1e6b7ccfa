#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
export NCCL_DEBUG=WARN
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/l_bench_2gpu.json 2> $O/l_bench_2gpu.err; echo "bench 2gpu rc $?"
head -c 300 $O/l_bench_2gpu.json; echo; tail -3 $O/l_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_train.py --model l --batch 4 --steps 20 --warmup 3 > $O/l_train_2gpu.txt 2>&1; echo "train 2gpu rc $?"
grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"allreduce": {[^}]*}' $O/l_train_2gpu.txt | tail -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/l_ref_2gpu.json 2> $O/l_ref_2gpu.err; echo "ref 2gpu rc $?"; head -c 200 $O/l_ref_2gpu.json; echo
