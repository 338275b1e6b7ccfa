"""Diagnostic probe of the tcgen05 conv kernel (run on the GPU box, one case per process so a
trap cannot poison later cases).  usage: python tools/tc_probe.py <case-index|all-list> [mode]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from streamyolo_b200 import ops
from streamyolo_b200.ops import View
from tests.test_gpu_ops import CONV_CASES, bf, rand_act, rand_w


def main():
    idx = int(sys.argv[1])
    n, ci, co, h, w, k, s = CONV_CASES[idx]
    print("case", idx, CONV_CASES[idx], flush=True)
    x, wt = rand_act(n, ci, h, w, 1), rand_w(co, ci, k, 2)
    ref = F.conv2d(x, wt, None, s, (k - 1) // 2)
    ho, wo = ops.conv_out_hw(h, w, k, s)
    y = View.empty(n, ho, wo, co, "cuda")
    y.buf.fill_(float("nan"))
    partials = torch.full((ops.conv_stat_rows(), 4 * co), float("nan"), device="cuda")
    ops.conv2d(ops.from_nchw(x), ops.pack_conv_weight(wt), y, k, s, ops.SY_CONV_RAW, impl="tc", partials=partials)
    torch.cuda.synchronize()
    got = y.nchw_float()
    nan = torch.isnan(got)
    print("nan fraction", nan.float().mean().item())
    g2 = torch.nan_to_num(got)
    rel = ((g2 - ref).norm() / ref.norm()).item()
    print("rel l2", rel)
    err = (g2 - ref).abs()
    rms = ref.pow(2).mean().sqrt().item()
    bad = err > (2 ** -8 * ref.abs() + 2 ** -8 * rms)
    print("bad fraction", bad.float().mean().item(), "max err", err.max().item(), "rms", rms)
    if bad.any():
        print("bad per image", bad.float().mean((1, 2, 3)).tolist())
        pc = bad.float().mean((0, 2, 3))
        print("bad per channel (first 32)", [round(v, 2) for v in pc[:32].tolist()])
        pr = bad.float().mean((0, 1, 3))
        print("bad per out row", [round(v, 2) for v in pr.tolist()])
        pcol = bad.float().mean((0, 1, 2))
        print("bad per out col", [round(v, 2) for v in pcol.tolist()])
        # correlation with candidates: is it a tap / k-ordering problem?
        sc = (g2 * ref).sum() / (ref * ref).sum()
        print("projection onto ref", sc.item())
    s1 = torch.nan_to_num(partials.view(-1, 2, 2, co)).sum(0)[0, 0]
    print("stats sum rel err", ((s1 - g2.sum((0, 2, 3))).abs().max() / (g2.abs().sum((0, 2, 3)).max() + 1e-9)).item())
    print("RESULT", "PASS" if (not bad.any() and not nan.any()) else "FAIL")


if __name__ == "__main__":
    main()
