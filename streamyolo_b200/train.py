"""One optimisation step on top of the training backward: the pieces of the reference's trainer loop that sit around
``loss.backward()`` (/root/reference/exps/train_utils/double_trainer.py:99-123, 171-175), as plain host code over
``model.backward.forward_backward`` and ``dist.allreduce_grads``.

  * ``build_optimizer``: [yolox 0.3.0] Exp.get_optimizer as configured by cfgs/*.py -- SGD, momentum 0.9, nesterov, three
    parameter groups (BatchNorm weights: no decay; conv / linear weights: weight decay 5e-4; biases: no decay).
  * ``ModelEMA``: [yolox 0.3.0] ModelEMA(model, 0.9998) with the decay ramp d * (1 - exp(-updates / 2000)).
  * ``train_step``: zero grads -> forward + backward (this GPU's shard) -> gradient mean over the ranks -> optimizer step
    -> EMA update.  Returns the loss dict.

STATUS (round 1): exercised on CPU with the kernels emulated (tests/test_cpu_backward.py); not yet run on a GPU."""
import copy
import math

import torch
from torch import nn

from . import dist as sydist
from .model import backward


def build_optimizer(model, lr, momentum=0.9, weight_decay=5e-4):
    pg0, pg1, pg2 = [], [], []          # BN weights | weights with decay | biases
    for _, m in model.named_modules():
        if hasattr(m, "bias") and isinstance(m.bias, nn.Parameter):
            pg2.append(m.bias)
        if isinstance(m, nn.BatchNorm2d):
            pg0.append(m.weight)
        elif hasattr(m, "weight") and isinstance(m.weight, nn.Parameter):
            pg1.append(m.weight)
    opt = torch.optim.SGD(pg0, lr=lr, momentum=momentum, nesterov=True)
    opt.add_param_group({"params": pg1, "weight_decay": weight_decay})
    opt.add_param_group({"params": pg2})
    return opt


class ModelEMA:
    def __init__(self, model, decay=0.9998, updates=0):
        self.ema = copy.deepcopy(model).eval()
        self.updates = updates
        self.decay = lambda x: decay * (1 - math.exp(-x / 2000))
        for p in self.ema.parameters():
            p.requires_grad_(False)

    @torch.no_grad()
    def update(self, model):
        self.updates += 1
        d = self.decay(self.updates)
        msd = model.state_dict()
        for k, v in self.ema.state_dict().items():
            if v.dtype.is_floating_point:
                v.mul_(d).add_((1.0 - d) * msd[k].detach())


def train_step(model, optimizer, x, targets, ema=None, grad_scale=1.0):
    for p in model.parameters():
        p.grad = None
    losses = backward.forward_backward(model, x, targets, grad_scale=grad_scale)
    sydist.allreduce_grads(model.parameters())
    if grad_scale != 1.0:
        for p in model.parameters():
            if p.grad is not None:
                p.grad.div_(grad_scale)
    optimizer.step()
    if ema is not None:
        ema.update(model)
    return losses
