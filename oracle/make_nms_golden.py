"""Generate tests/golden/nms_*.npz with torchvision (the reference's own NMS: yolox.utils.postprocess calls
torchvision.ops.batched_nms) -- TEST INFRASTRUCTURE.  Run in the build container:

    python oracle/make_nms_golden.py

Each fixture holds a synthetic eval-style prediction [A, 13], the thresholds, and the indices torchvision keeps (per-class
torchvision.ops.nms, merged in decreasing score order = the definition of batched_nms), so that the oracle's restatement
stays pinned even where torchvision is not installed."""
import os
import sys

import numpy as np
import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_postprocess import synth_pred  # noqa: E402

for seed, a, conf, thr in ((11, 2000, 0.05, 0.65), (12, 1500, 0.2, 0.45), (13, 3000, 0.01, 0.3)):
    p = synth_pred(1, a, 8, seed)[0]
    xyxy = torch.stack([p[:, 0] - p[:, 2] / 2, p[:, 1] - p[:, 3] / 2, p[:, 0] + p[:, 2] / 2, p[:, 1] + p[:, 3] / 2], 1)
    cconf, cls = torch.max(p[:, 5:], 1)
    score = p[:, 4] * cconf
    idx = (score >= conf).nonzero().flatten()
    keep = []
    for c in cls[idx].unique():
        m = idx[cls[idx] == c]
        keep.append(m[torchvision.ops.nms(xyxy[m], score[m], thr)])
    keep = torch.cat(keep)
    keep = keep[torch.argsort(score[keep], descending=True, stable=True)]
    path = os.path.join(ROOT, "tests", "golden", f"nms_{seed}.npz")
    np.savez_compressed(path, pred=p.numpy(), conf=np.float32(conf), thr=np.float32(thr), keep=keep.numpy(),
                        torchvision=np.array(torchvision.__version__))
    print(path, len(idx), "candidates ->", len(keep), "kept")
