"""PIPEHead: the still-image baseline head of the reference (mirror of /root/reference/exps/model/pipe_head.py, used by
cfgs/l_s50_still_dfp_flip.py) = TALHead without the trend-aware weighting: same towers / prediction convs / decode / SimOTA,
every foreground anchor weighs 1 in the IoU and L1 terms, and ``labels`` is ONE tensor [B, max_labels, 5] instead of the
(future, current) pair.  Runs on the same kernels: gamma = 0 makes the TAL weight 1 / (iou^0 + 1e-8) a constant, which the
normalisation w * sum(l) / sum(w * l) (tal_head.py:429-437) turns into 1 (to float roundoff)."""
import torch

from .tal_head import TALHead


class PIPEHead(TALHead):
    def __init__(self, num_classes, width=1.0, strides=[8, 16, 32], in_channels=[256, 512, 1024], act="silu", depthwise=False):
        super().__init__(num_classes, width, strides, in_channels, act, depthwise, gamma=0.0, ignore_thr=0.0, ignore_value=1.0)

    def forward(self, xin, labels=None, imgs=None):
        if labels is not None and torch.is_tensor(labels):
            labels = (labels, labels)
        return super().forward(xin, labels, imgs)
