"""Per-launch durations of one training-forward step with WARM caches: every product op is bracketed by CUDA events in
an eager (non-graph) run, so the GPU idles between launches but L2 holds what the previous kernels left there -- the
complement of the ncu launch list (cold caches, serialised).  Prints per-kind totals and the slowest layers.

usage: python tools/hot_trace.py [model] [pairs] [top]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from streamyolo_b200 import ops, synth
from streamyolo_b200.model import engine

tag = sys.argv[1] if len(sys.argv) > 1 else "l"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
TOP = int(sys.argv[3]) if len(sys.argv) > 3 else 40
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
model = bench.build_model(tag, dev)
engine.name_modules(model)
x = synth.synth_frames(B, 600, 960, seed=1234).to(dev)
fut, cur = synth.synth_labels(B, 600, 960, seed=1)
fut, cur = fut.to(dev), cur.to(dev)

REC = []
CUR = {"name": "?"}
orig_conv_bn_act = engine.conv_bn_act


def named_conv_bn_act(ctx, mods, *a, **k):
    CUR["name"] = "|".join(getattr(m, "_sy_name", "?") for m in mods)
    return orig_conv_bn_act(ctx, mods, *a, **k)


engine.conv_bn_act = named_conv_bn_act


def wrap(fname, kind, describe):
    orig = getattr(ops, fname)

    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(*a, **k)
        e1.record()
        REC.append((kind, CUR["name"], describe(*a, **k), e0, e1))
        return r
    setattr(ops, fname, f)


def d_conv(x, wpk, y, k, s, mode, *a, **kw):
    kh, kw_ = (k, k) if isinstance(k, int) else k
    return f"{x.n}x{x.h}x{x.w} {x.c}->{y.c} k{kh}x{kw_}s{s}", 2.0 * y.n * y.h * y.w * y.c * x.c * kh * kw_, \
        2.0 * (x.n * x.h * x.w * x.c + y.n * y.h * y.w * y.c)


def d_apply(x, sc, sh, split, act, res, y, *a, **kw):
    return f"{x.n}x{x.h}x{x.w}x{x.c}" + (" +res" if res is not None else ""), 0.0, \
        2.0 * x.n * x.h * x.w * x.c * (3 if res is not None else 2)


def d_other(*a, **kw):
    return "", 0.0, 0.0


wrap("conv2d", "conv", d_conv)
wrap("bn_act_apply", "apply", d_apply)
for name in ("focus_pack", "upsample_nearest", "spp_maxpool", "copy", "head_pred_decode", "tal_loss"):
    if hasattr(ops, name):
        wrap(name, name, d_other)

with torch.no_grad():
    for _ in range(2):
        model(x, (fut, cur))
    torch.cuda.synchronize()
    REC.clear()
    torch.cuda._sleep(int(60e6))      # ~30 ms blocker: the whole step is queued behind it, so the events are
    model(x, (fut, cur))              # time-stamped back to back on the device, not at host launch pace
    torch.cuda.synchronize()

rows = [(kind, name, desc[0], desc[1], desc[2], e0.elapsed_time(e1) * 1e3) for kind, name, desc, e0, e1 in REC]
tot = collections.defaultdict(float)
cnt = collections.Counter()
fl = collections.defaultdict(float)
by = collections.defaultdict(float)
for kind, name, d, f, b, us in rows:
    tot[kind] += us
    cnt[kind] += 1
    fl[kind] += f
    by[kind] += b
T = sum(tot.values())
print(f"warm per-op event timing, StreamYOLO-{tag}, {B} pairs: {len(rows)} ops, sum {T / 1e3:.3f} ms")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"  {k:18s} {v / 1e3:7.3f} ms  {cnt[k]:4d} launches  avg {v / cnt[k]:6.1f} us  {fl[k] / v / 1e6 if v else 0:7.1f} TFLOP/s  {by[k] / v / 1e3 if v else 0:7.0f} GB/s")
print(f"\n{'us':>8s} {'TFLOP/s':>8s} {'GB/s':>7s}  op")
for kind, name, d, f, b, us in sorted(rows, key=lambda r: -r[5])[:TOP]:
    print(f"{us:8.1f} {f / us / 1e6:8.1f} {b / us / 1e3:7.0f}  {kind:6s} {name[-44:]:44s} {d}")
if "--all" in sys.argv:
    print("\nin launch order")
    for kind, name, d, f, b, us in rows:
        print(f"{us:8.1f} {f / us / 1e6:8.1f} {b / us / 1e3:7.0f}  {kind:6s} {name[-44:]:44s} {d}")
