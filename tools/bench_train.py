"""Training-step benchmark (BASELINE.json configs 2-4: forward + backward, optionally the DDP step) -- companion of bench.py,
which measures the headline forward+loss metric.  NOT YET RUN ON A GPU (written at the end of round 1).

    python tools/bench_train.py [--model s|m|l] [--batch 8] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py ...

One step = streamyolo_b200.train.train_step: recording forward, reverse walk, gradient mean over the ranks (NCCL),
SGD-nesterov step, EMA update.  Timed with CUDA events, max over ranks; prints one JSON line on rank 0.
Algorithmic work: 3 x the forward conv FLOPs."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from streamyolo_b200 import dist as sydist, synth, train


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="l", choices=list(bench.MODELS))
    ap.add_argument("--batch", type=int, default=8, help="frame pairs per GPU")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-ema", action="store_true")
    args = ap.parse_args()
    bench.guard_stdout()
    rank, local_rank, world = sydist.env_world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sydist.init("nccl")
    model = bench.build_model(args.model, dev)
    opt = train.build_optimizer(model, lr=0.01 / 64 * args.batch * world)
    ema = None if args.no_ema else train.ModelEMA(model)
    x = synth.synth_frames(args.batch, 600, 960, seed=1234 + rank).to(dev)
    fut, cur = synth.synth_labels(args.batch, 600, 960, seed=1 + rank)
    tg = (fut.to(dev), cur.to(dev))
    for _ in range(args.warmup):
        losses = train.train_step(model, opt, x, tg, ema)
    torch.cuda.synchronize()
    sydist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        losses = train.train_step(model, opt, x, tg, ema)
    e1.record()
    torch.cuda.synchronize()
    sydist.barrier()
    ms = sydist.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    if rank == 0:
        gf = bench.GFLOP_PER_PAIR[args.model] * 3.0
        pairs = world * args.batch / (ms * 1e-3)
        bench.emit(json.dumps({
            "metric": "frame-pairs/sec StreamYOLO-%s 600x960 fwd+bwd+step" % args.model, "value": round(pairs, 2), "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
            "scaling": "weak", "dtype": "bf16 activations / gradients, fp32 parameters", "data": "synthetic",
            "config": {"workload": "StreamYOLO-%s (random init) 600x960 frame pairs, training step, %d pairs/GPU" % (args.model, args.batch),
                       "parallelism": "dp%d, gradient mean over NCCL" % world, "ema": ema is not None},
            "tflops": round(pairs * gf / 1e3, 1), "loss": float(losses["total_loss"])}))
    sydist.shutdown()


if __name__ == "__main__":
    main()
