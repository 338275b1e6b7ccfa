"""Drop-in for [yolox 0.3.0] ``yolox.utils.postprocess`` (imported by the reference's evaluators,
/root/reference/exps/evaluators/onex_stream_evaluator.py:14,148, and by sAP/streamyolo/streamyolo_det.py): same
signature, same return value (a list with one ``[n_i, 7]`` tensor or ``None`` per image, rows
``[x1, y1, x2, y2, obj_conf, class_conf, class_pred]`` in decreasing score order), computed by ``sy_postprocess_nms``
(one CTA per image, no per-image Python loop, no torchvision)."""
import torch

from . import ops


def postprocess(prediction, num_classes, conf_thre=0.7, nms_thre=0.45, class_agnostic=False):
    pred = prediction.detach().float().contiguous()
    det, count = ops.postprocess_nms(pred, num_classes, float(conf_thre), float(nms_thre), class_agnostic)
    counts = count.tolist()                       # the one host sync: the per-image row counts
    return [det[i, :n].to(prediction.dtype) if n > 0 else None for i, n in enumerate(counts)]
