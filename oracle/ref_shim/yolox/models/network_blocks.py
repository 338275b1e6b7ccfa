"""Restatement of the yolox==0.3.0 building blocks used by the reference
(call sites: /root/reference/exps/model/darknet.py:7, dfp_pafpn.py:10,
tal_head.py:16).  Sub-module attribute names fix the state_dict keys.
Test infrastructure only."""
import torch
from torch import nn

_ACTS = {"silu": lambda: nn.SiLU(inplace=True), "relu": lambda: nn.ReLU(inplace=True),
         "lrelu": lambda: nn.LeakyReLU(0.1, inplace=True)}


class BaseConv(nn.Module):
    def __init__(self, in_channels, out_channels, ksize, stride, groups=1, bias=False, act="silu"):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, ksize, stride, (ksize - 1) // 2,
                              groups=groups, bias=bias)
        self.bn = nn.BatchNorm2d(out_channels)
        self.act = _ACTS[act]()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))

    def fuseforward(self, x):
        return self.act(self.conv(x))


class DWConv(nn.Module):
    def __init__(self, in_channels, out_channels, ksize, stride=1, act="silu"):
        super().__init__()
        self.dconv = BaseConv(in_channels, in_channels, ksize, stride, groups=in_channels, act=act)
        self.pconv = BaseConv(in_channels, out_channels, 1, 1, act=act)

    def forward(self, x):
        return self.pconv(self.dconv(x))


class Bottleneck(nn.Module):
    def __init__(self, in_channels, out_channels, shortcut=True, expansion=0.5,
                 depthwise=False, act="silu"):
        super().__init__()
        mid = int(out_channels * expansion)
        second = DWConv if depthwise else BaseConv
        self.conv1 = BaseConv(in_channels, mid, 1, 1, act=act)
        self.conv2 = second(mid, out_channels, 3, 1, act=act)
        self.use_add = shortcut and in_channels == out_channels

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        return y + x if self.use_add else y


class ResLayer(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.layer1 = BaseConv(in_channels, in_channels // 2, 1, 1, act="lrelu")
        self.layer2 = BaseConv(in_channels // 2, in_channels, 3, 1, act="lrelu")

    def forward(self, x):
        return x + self.layer2(self.layer1(x))


class SPPBottleneck(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_sizes=(5, 9, 13), activation="silu"):
        super().__init__()
        mid = in_channels // 2
        self.conv1 = BaseConv(in_channels, mid, 1, 1, act=activation)
        self.m = nn.ModuleList([nn.MaxPool2d(k, 1, k // 2) for k in kernel_sizes])
        self.conv2 = BaseConv(mid * (len(kernel_sizes) + 1), out_channels, 1, 1, act=activation)

    def forward(self, x):
        x = self.conv1(x)
        return self.conv2(torch.cat([x] + [m(x) for m in self.m], 1))


class CSPLayer(nn.Module):
    def __init__(self, in_channels, out_channels, n=1, shortcut=True, expansion=0.5,
                 depthwise=False, act="silu"):
        super().__init__()
        mid = int(out_channels * expansion)
        self.conv1 = BaseConv(in_channels, mid, 1, 1, act=act)
        self.conv2 = BaseConv(in_channels, mid, 1, 1, act=act)
        self.conv3 = BaseConv(2 * mid, out_channels, 1, 1, act=act)
        self.m = nn.Sequential(*[Bottleneck(mid, mid, shortcut, 1.0, depthwise, act=act)
                                 for _ in range(n)])

    def forward(self, x):
        return self.conv3(torch.cat((self.m(self.conv1(x)), self.conv2(x)), 1))


class Focus(nn.Module):
    def __init__(self, in_channels, out_channels, ksize=1, stride=1, act="silu"):
        super().__init__()
        self.conv = BaseConv(in_channels * 4, out_channels, ksize, stride, act=act)

    def forward(self, x):
        tl, bl = x[..., ::2, ::2], x[..., 1::2, ::2]
        tr, br = x[..., ::2, 1::2], x[..., 1::2, 1::2]
        return self.conv(torch.cat((tl, bl, tr, br), 1))
