"""Secondary measurements (not the headline): eval-mode forward throughput (BASELINE.json config 5 shape: StreamYOLO-l,
8 frame pairs per GPU, model.eval(), decoded [B, 11850, 13] output, NMS excluded) and on_pipe streaming latency
(one 600x960 frame per call with the feature buffer carried over, batch 1).
CUDA graphs + CUDA events; prints one JSON line per mode.   usage: python tools/bench_modes.py [model]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from streamyolo_b200 import synth

tag = sys.argv[1] if len(sys.argv) > 1 else "l"
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
model = bench.build_model(tag, dev).eval()
gf = bench.GFLOP_PER_PAIR[tag]


def timed(fn, steps=30, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


with torch.no_grad():
    # ---------------- eval forward, off_pipe, B = 8 pairs
    B = 8
    x = synth.synth_frames(B, 600, 960).to(dev)
    for _ in range(2):
        model(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        model(x)
    torch.cuda.current_stream().wait_stream(st)
    with torch.cuda.graph(g):
        out = model(x)
    ms = timed(g.replay)
    print(json.dumps({"mode": "eval off_pipe", "model": tag, "pairs_per_gpu": B, "ms_per_step": round(ms, 4),
                      "pairs_per_s": round(B / ms * 1e3, 1), "tflops": round(B / ms * gf, 1),
                      "out_shape": list(out.shape), "note": "BN folded into the conv epilogue, NMS excluded"}))
    # ---------------- on_pipe streaming, batch 1
    f0 = synth.synth_frames(1, 600, 960)[:, :3].contiguous().to(dev)
    o, buf = model(f0, mode="on_pipe")
    buf_static = tuple(b.clone() for b in buf)
    for _ in range(2):
        model(f0, buffer=buf_static, mode="on_pipe")
    torch.cuda.synchronize()
    g2 = torch.cuda.CUDAGraph()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        model(f0, buffer=buf_static, mode="on_pipe")
    torch.cuda.current_stream().wait_stream(st)
    with torch.cuda.graph(g2):
        o2, nb = model(f0, buffer=buf_static, mode="on_pipe")
        for d_, s_ in zip(buf_static, nb):
            d_.copy_(s_)                              # carry the feature buffer to the next frame
    ms2 = timed(g2.replay)
    print(json.dumps({"mode": "on_pipe", "model": tag, "batch": 1, "ms_per_frame": round(ms2, 4),
                      "fps": round(1e3 / ms2, 1), "budget_ms": 33.3,
                      "note": "single 600x960 frame + buffered previous-frame features -> decoded [1, 11850, 13]"}))
