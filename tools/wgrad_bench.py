"""Weight-gradient kernel timing (CUDA graph of 10 back-to-back launches, CUDA events).  usage: python tools/wgrad_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from streamyolo_b200 import ops
from streamyolo_b200.ops import View

SHAPES = [(16, 128, 128, 75, 120, 3, 1), (8, 256, 256, 75, 120, 3, 1), (16, 256, 256, 38, 60, 3, 1), (16, 512, 512, 19, 30, 3, 1),
          (16, 128, 128, 75, 120, 1, 1), (16, 64, 128, 300, 480, 3, 2), (16, 1024, 1024, 19, 30, 1, 1)]
for n, ci, co, h, w, k, s in SHAPES:
    ho, wo = ops.conv_out_hw(h, w, k, s)
    x = View(torch.randn((n, h, w, ci), device="cuda").to(torch.bfloat16))
    dy = View(torch.randn((n, ho, wo, co), device="cuda").to(torch.bfloat16))
    dw = torch.empty((co, ci, k, k), device="cuda")
    ws = ops.conv2d_wgrad(x, dy, k, s, dw)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g):
            for _ in range(10):
                ops.conv2d_wgrad(x, dy, k, s, dw, workspace=ws)
        g.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / 10)
    fl = 2.0 * n * ho * wo * co * ci * k * k
    print(f"wgrad {(n, ci, co, h, w, k, s)!s:34s} {best:7.1f} us  {fl / best / 1e6:7.0f} TFLOP/s  workspace {ws.numel() / 1e6:.1f} MB", flush=True)
