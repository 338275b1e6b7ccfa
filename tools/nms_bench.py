"""Timing of the device post-processing (sy_postprocess_nms): conf threshold + class-aware NMS for 8 images of 11 850 anchors,
for several candidate counts (the number of anchors above the confidence threshold).  usage: python tools/nms_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from streamyolo_b200.postprocess import postprocess
from test_postprocess import synth_pred

B, A = 8, 11850
pred = synth_pred(B, A, 8, 123).cuda()
score = (pred[..., 4] * pred[..., 5:].max(-1).values).flatten()
for frac in (0.02, 0.1, 0.3, 1.0):
    conf = float(torch.quantile(score[:100000], 1.0 - frac)) if frac < 1.0 else 0.0
    for _ in range(2):
        out = postprocess(pred, 8, conf, 0.65)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = postprocess(pred, 8, conf, 0.65)
    e1.record()
    torch.cuda.synchronize()
    ncand = int((score >= conf).sum()) // B
    kept = sum(0 if o is None else o.shape[0] for o in out) // B
    print(f"conf {conf:.4f}: ~{ncand} candidates / image, {kept} kept: {e0.elapsed_time(e1) / 5:.3f} ms for {B} images (incl. the wrapper's host sync)")
