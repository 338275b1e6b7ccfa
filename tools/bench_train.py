"""Training-step benchmark (BASELINE.json configs 2-4: forward + backward + the DDP step) -- the same measurement
bench.py reports under its "train" key, as a stand-alone tool with more switches.

    python tools/bench_train.py [--model s|m|l] [--batch 8] [--steps 10] [--warmup 3] [--impl trainer|stock] [--no-overlap]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py ...

--impl trainer  streamyolo_b200.train.Trainer: recording forward, reverse walk writing into the flat gradient buffer, bucketed
                NCCL all-reduce launched from the walk, fused SGD-nesterov + EMA kernel
--impl stock    train.train_step: same walk, torch.optim.SGD + Python EMA + post-hoc all-reduce (round-1 path)
Timed with CUDA events, max over ranks; prints one JSON line on rank 0.  Algorithmic work: 3 x the forward conv FLOPs."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from streamyolo_b200 import dist as sydist, ops, synth, train


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="l", choices=list(bench.MODELS))
    ap.add_argument("--batch", type=int, default=8, help="frame pairs per GPU")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="trainer", choices=["trainer", "stock"])
    ap.add_argument("--no-ema", action="store_true")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--split", action="store_true", help="also time forward+backward and the optimiser step separately")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph: every kernel launched from Python each step")
    args = ap.parse_args()
    bench.guard_stdout()
    rank, local_rank, world = sydist.env_world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sydist.init("nccl")
    model = bench.build_model(args.model, dev)
    lr = 0.01 / 64 * args.batch * world
    x = synth.synth_frames(args.batch, 600, 960, seed=1234 + rank).to(dev)
    fut, cur = synth.synth_labels(args.batch, 600, 960, seed=1 + rank)
    tg = (fut.to(dev), cur.to(dev))
    if args.impl == "trainer":
        tr = train.Trainer(model, lr=lr, use_ema=not args.no_ema, overlap=not args.no_overlap)
        segments = 0
        if not args.eager:
            segments = tr.capture(x, tg)

        def step():
            return tr.step(x, tg) if args.eager else tr.replay()
    else:
        opt = train.build_optimizer(model, lr=lr)
        ema = None if args.no_ema else train.ModelEMA(model)

        def step():
            return train.train_step(model, opt, x, tg, ema)
    for _ in range(args.warmup):
        losses = step()
    torch.cuda.synchronize()
    sydist.barrier()
    ops.LAUNCHES = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        losses = step()
    e1.record()
    launches = ops.LAUNCHES
    host_ms = (time.perf_counter() - t0) * 1e3 / args.steps      # host time to ENQUEUE a step (launch-bound if ~ ms_per_step)
    torch.cuda.synchronize()
    sydist.barrier()
    ms = sydist.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    extra = {}
    if args.split and args.impl == "trainer" and args.eager:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        for _ in range(args.steps):
            tr.forward_backward(x, tg)
        ev[1].record()
        for _ in range(args.steps):
            tr.optimizer_step()
        ev[2].record()
        torch.cuda.synchronize()
        extra = {"fwd_bwd_ms": round(ev[0].elapsed_time(ev[1]) / args.steps, 3),
                 "optimizer_step_ms": round(ev[1].elapsed_time(ev[2]) / args.steps, 3)}
    if rank == 0:
        gf = bench.GFLOP_PER_PAIR.get(args.model, 0.0) * 3.0
        pairs = world * args.batch / (ms * 1e-3)
        line = {
            "metric": "frame-pairs/sec StreamYOLO-%s 600x960 fwd+bwd+step" % args.model, "value": round(pairs, 2), "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
            "scaling": "weak", "dtype": "bf16 activations / gradients, fp32 parameters", "data": "synthetic",
            "config": {"workload": "StreamYOLO-%s (random init) 600x960 frame pairs, training step, %d pairs/GPU" % (args.model, args.batch),
                       "parallelism": "dp%d, gradient mean over NCCL" % world, "ema": not args.no_ema, "impl": args.impl,
                       "overlap": not args.no_overlap},
            "tflops": round(pairs / world * gf / 1e3, 1), "loss": float(losses["total_loss"]),
            "host_enqueue_ms_per_step": round(host_ms, 3), "launches_per_step": launches // max(1, args.steps)}
        if args.impl == "trainer":
            line["allreduce"] = {"buckets": len(tr.sink.launched), "bytes": 4 * tr.fs.n_param, "world": world}
            line["config"]["cuda_graph_segments"] = segments
        line.update(extra)
        bench.emit(json.dumps(line))
    sydist.shutdown()


if __name__ == "__main__":
    main()
