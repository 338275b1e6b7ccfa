"""yolox==0.3.0 IOUloss restated (call site /root/reference/exps/model/tal_head.py:15,136).
Test infrastructure only."""
import torch
from torch import nn


class IOUloss(nn.Module):
    def __init__(self, reduction="none", loss_type="iou"):
        super().__init__()
        self.reduction, self.loss_type = reduction, loss_type

    def forward(self, pred, target):
        assert pred.shape[0] == target.shape[0]
        pred, target = pred.view(-1, 4), target.view(-1, 4)
        tl = torch.max(pred[:, :2] - pred[:, 2:] / 2, target[:, :2] - target[:, 2:] / 2)
        br = torch.min(pred[:, :2] + pred[:, 2:] / 2, target[:, :2] + target[:, 2:] / 2)
        area_p, area_g = torch.prod(pred[:, 2:], 1), torch.prod(target[:, 2:], 1)
        en = (tl < br).type(tl.type()).prod(dim=1)
        area_i = torch.prod(br - tl, 1) * en
        area_u = area_p + area_g - area_i
        iou = area_i / (area_u + 1e-16)
        if self.loss_type == "iou":
            loss = 1 - iou ** 2
        else:  # giou (unused by the reference configs)
            c_tl = torch.min(pred[:, :2] - pred[:, 2:] / 2, target[:, :2] - target[:, 2:] / 2)
            c_br = torch.max(pred[:, :2] + pred[:, 2:] / 2, target[:, :2] + target[:, 2:] / 2)
            area_c = torch.prod(c_br - c_tl, 1)
            loss = 1 - (iou - (area_c - area_u) / area_c.clamp(1e-16)).clamp(min=-1.0, max=1.0)
        if self.reduction == "mean":
            loss = loss.mean()
        elif self.reduction == "sum":
            loss = loss.sum()
        return loss
