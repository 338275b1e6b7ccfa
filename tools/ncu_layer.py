"""One BaseConv of the benchmark workload in train mode (tcgen05 conv with statistics + in-kernel BatchNorm finalize, then the
normalise pass), launched a few times -- the target of `ncu --set full -k regex:conv_tc_kernel -s 2 -c 1` and of
compute-sanitizer runs.    python tools/ncu_layer.py n cin cout h w k stride [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from streamyolo_b200 import ops
from streamyolo_b200.ops import View

n, ci, co, h, w, k, s = map(int, sys.argv[1:8])
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 3
dev = "cuda"
x = View(torch.randn((n, h, w, ci), device=dev).to(torch.bfloat16))
wt = ops.pack_conv_weight(torch.randn((co, ci, k, k), device=dev) * 0.05)
ho, wo = ops.conv_out_hw(h, w, k, s)
raw, y = View.empty(n, ho, wo, co, dev), View.empty(n, ho, wo, co, dev)
gamma, beta = torch.ones(co, device=dev), torch.zeros(co, device=dev)
rm, rv = torch.zeros(co, device=dev), torch.ones(co, device=dev)
nbt = torch.zeros((), dtype=torch.long, device=dev)
partials = torch.empty((ops.conv_stat_rows(), 4 * co), device=dev)
ss = torch.empty((2, 2, co), device=dev)
sync = torch.zeros(4, dtype=torch.int32, device=dev)
for _ in range(reps):
    ops.conv2d(x, wt, raw, k, s, ops.SY_CONV_RAW, partials=partials, split_n=n // 2, bn=[(gamma, beta, rm, rv, nbt, 0)],
               momentum=0.03, eps=1e-3, scale_shift=ss, sync=sync)
    ops.bn_act_apply(raw, ss[0], ss[1], n // 2, 1, None, y)
torch.cuda.synchronize()
print("ok", float(y.torch().float().abs().mean()), sync.tolist())
