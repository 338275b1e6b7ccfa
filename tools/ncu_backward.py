"""The backward bricks of one BaseConv (BatchNorm+SiLU backward: reduce / finalize / apply; weight gradient + split-K
reduction) and the two HBM-bound forward glue kernels (Focus packing, prediction convs + decode), launched a few times -- the
target of `ncu --set full -k regex:<kernel> -s 2 -c 1`.    python tools/ncu_backward.py n cin cout h w k stride [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from streamyolo_b200 import ops
from streamyolo_b200.ops import View

n, ci, co, h, w, k, s = map(int, sys.argv[1:8])
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 3
dev = "cuda"
ho, wo = ops.conv_out_hw(h, w, k, s)
x = View(torch.randn((n, h, w, ci), device=dev).to(torch.bfloat16))
raw = View(torch.randn((n, ho, wo, co), device=dev).to(torch.bfloat16))
dy = View(torch.randn((n, ho, wo, co), device=dev).to(torch.bfloat16))
draw = View.empty(n, ho, wo, co, dev)
scale, shift = torch.rand((2, co), device=dev) + 0.5, torch.rand((2, co), device=dev) - 0.5
mean, invstd = torch.zeros((2, co), device=dev), torch.ones((2, co), device=dev)
dgamma, dbeta = torch.empty(co, device=dev), torch.empty(co, device=dev)
dw = torch.empty((co, ci, k, k), device=dev)
frames = torch.rand((8, 6, 600, 960), device=dev) * 255
xin = View.empty(16, 300, 480, 64, dev)
cf, rf = View(torch.randn((8, 75, 120, 256), device=dev).to(torch.bfloat16)), View(torch.randn((8, 75, 120, 256), device=dev).to(torch.bfloat16))
wr, br = torch.randn(4, 256, device=dev) * 0.05, torch.zeros(4, device=dev)
wo_, bo = torch.randn(1, 256, device=dev) * 0.05, torch.zeros(1, device=dev)
wc, bc = torch.randn(8, 256, device=dev) * 0.05, torch.zeros(8, device=dev)
out = torch.empty((8, 9000, 13), device=dev)
ws = None
for _ in range(reps):
    ops.bn_act_backward(raw, dy, draw, scale, shift, mean, invstd, n // 2, 1, dgamma, dbeta)
    ws = ops.conv2d_wgrad(x, draw, k, s, dw, workspace=ws)
    ops.focus_pack(frames, 2, xin)
    ops.head_pred_decode(cf, rf, wr, br, wo_, bo, wc, bc, 8, 0, 9000, out, None, sigmoid=True, decode=True)
torch.cuda.synchronize()
print("ok", float(dw.abs().mean()), float(draw.torch().float().abs().mean()))
