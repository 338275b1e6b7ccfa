"""CPU restatement of the detection post-processing (TEST INFRASTRUCTURE ONLY: imported by tests/ and never by the
product).  It follows

  * [yolox 0.3.0] yolox/utils/boxes.py: postprocess (not vendored in /root/reference; call sites
    /root/reference/exps/evaluators/onex_stream_evaluator.py:148, sAP/streamyolo/streamyolo_det.py:62-83; restated from
    the published source): cxcywh -> xyxy, class_conf / class_pred = max over classes, conf mask on obj * class_conf,
    detections [x1, y1, x2, y2, obj, class_conf, class_pred], batched_nms, gather;
  * torchvision.ops.batched_nms / the nms CPU kernel (torchvision/csrc/ops/cpu/nms_kernel.cpp): scores sorted
    descending, greedy suppression with ovr = inter / (iarea + jarea - inter) > thr, per class.

Pinned by tests/test_postprocess.py against torchvision.ops.batched_nms itself (installed in this image)."""
import numpy as np
import torch


def nms_greedy(boxes: np.ndarray, scores: np.ndarray, classes: np.ndarray, thr: float, class_agnostic=False) -> np.ndarray:
    """Indices kept, in decreasing score order (ties: lower index first).  fp32 arithmetic like the kernel."""
    order = np.lexsort((np.arange(len(scores)), -scores.astype(np.float64)))
    b = boxes.astype(np.float32)
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = ((x2 - x1).astype(np.float32) * (y2 - y1).astype(np.float32)).astype(np.float32)
    removed = np.zeros(len(scores), bool)
    keep = []
    thr = np.float32(thr)
    for oi, i in enumerate(order):
        if removed[i]:
            continue
        keep.append(i)
        rest = order[oi + 1:]
        rest = rest[~removed[rest]]
        if not class_agnostic:
            rest = rest[classes[rest] == classes[i]]
        if len(rest) == 0:
            continue
        xx1, yy1 = np.maximum(x1[i], x1[rest]), np.maximum(y1[i], y1[rest])
        xx2, yy2 = np.minimum(x2[i], x2[rest]), np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), (xx2 - xx1).astype(np.float32))
        h = np.maximum(np.float32(0), (yy2 - yy1).astype(np.float32))
        inter = (w * h).astype(np.float32)
        ovr = inter / ((areas[i] + areas[rest]).astype(np.float32) - inter).astype(np.float32)
        removed[rest[ovr > thr]] = True
    return np.array(keep, np.int64)


def postprocess_oracle(prediction: torch.Tensor, num_classes: int, conf_thre=0.7, nms_thre=0.45, class_agnostic=False):
    pred = prediction.detach().float().cpu()
    out = []
    for p in pred:
        half_w, half_h = p[:, 2] / 2, p[:, 3] / 2
        xyxy = torch.stack([p[:, 0] - half_w, p[:, 1] - half_h, p[:, 0] + half_w, p[:, 1] + half_h], 1)
        class_conf, class_pred = torch.max(p[:, 5:5 + num_classes], 1)
        score = p[:, 4] * class_conf
        mask = score >= conf_thre
        idx = mask.nonzero().flatten()
        if idx.numel() == 0:
            out.append(None)
            continue
        keep = nms_greedy(xyxy[idx].numpy(), score[idx].numpy(), class_pred[idx].numpy(), nms_thre, class_agnostic)
        sel = idx[torch.from_numpy(keep)]
        out.append(torch.cat([xyxy[sel], p[sel, 4:5], class_conf[sel, None], class_pred[sel, None].float()], 1))
    return out
