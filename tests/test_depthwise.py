"""Depthwise path (north_star: "3x3 depthwise ... as coalesced / vectorised HBM kernels"): [yolox] DWConv = depthwise k x k
BaseConv + pointwise 1x1 BaseConv, selected by depthwise=True in CSPDarknet / DFPPAFPN / TALHead
(/root/reference/exps/model/darknet.py:109, dfp_pafpn.py:31, tal_head.py:53).  Reference = the in-repo stand-in of the yolox
0.3.0 blocks (oracle/ref_shim, test infrastructure) with the SAME state_dict, fp32, on the CPU.

  * CPU (kernels emulated in torch, fp32 storage): DWConv, depthwise Bottleneck / CSPLayer modules train + eval to float
    roundoff; a whole depthwise YOLOX constructs with the reference's state_dict layout and runs train / eval forwards.
  * GPU: sy_dwconv2d against F.conv2d(groups=C) on the same bf16 operands (k = 1, 3, 5; stride 1, 2; odd sizes; RAW and
    FUSED with residual); the DWConv module (train-mode BatchNorm through the CUDA-core statistics kernels) against the
    reference module.
"""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))
import emul_ops  # noqa: E402
from streamyolo_b200 import synth  # noqa: E402
from streamyolo_b200.model import DFPPAFPN, TALHead, YOLOX  # noqa: E402
from streamyolo_b200.model import network_blocks as nb  # noqa: E402


def _ref_blocks():
    from yolox.models import network_blocks as rb          # oracle/ref_shim stand-in of yolox 0.3.0
    return rb


def _sync(ref, prod, seed=0):
    sd = synth.synth_state_dict({k: tuple(v.shape) for k, v in ref.state_dict().items()}, seed)
    ref.load_state_dict(sd, strict=True)
    prod.load_state_dict(sd, strict=True)                  # identical keys: the drop-in surface
    for m in list(ref.modules()) + list(prod.modules()):
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eps, m.momentum = 1e-3, 0.03


@pytest.mark.parametrize("kind", ["dwconv_s1", "dwconv_s2", "bottleneck", "csp"])
def test_depthwise_modules_match_reference_blocks(kind, monkeypatch):
    emul_ops.install(monkeypatch, exact=True)
    rb = _ref_blocks()
    if kind == "dwconv_s1":
        ref, prod = rb.DWConv(16, 24, 3, 1), nb.DWConv(16, 24, 3, 1)
    elif kind == "dwconv_s2":
        ref, prod = rb.DWConv(16, 32, 3, 2), nb.DWConv(16, 32, 3, 2)
    elif kind == "bottleneck":
        ref, prod = rb.Bottleneck(16, 16, True, 1.0, True), nb.Bottleneck(16, 16, True, 1.0, True)
    else:
        ref, prod = rb.CSPLayer(32, 32, 2, True, 0.5, True), nb.CSPLayer(32, 32, 2, True, 0.5, True)
    _sync(ref, prod)
    x = torch.randn(2, ref.state_dict()[next(iter(ref.state_dict()))].shape[0] if kind.startswith("dwconv") else (16 if kind == "bottleneck" else 32), 11, 13)
    for train in (True, False):
        ref.train(train)
        prod.train(train)
        with torch.no_grad():
            want, got = ref(x.clone()), prod(x.clone())
        assert torch.allclose(got.float(), want, rtol=1e-4, atol=1e-5), (kind, train, float((got.float() - want).abs().max()))
    for (k, a), b in zip(prod.state_dict().items(), ref.state_dict().values()):      # running statistics moved identically
        assert torch.allclose(a.float(), b.float(), rtol=1e-4, atol=1e-6), k


def test_depthwise_model_constructs_and_runs(monkeypatch):
    emul_ops.install(monkeypatch, exact=True)
    ch = [256, 512, 1024]
    m = YOLOX(DFPPAFPN(0.33, 0.125, in_channels=ch, depthwise=True), TALHead(8, 0.125, in_channels=ch, depthwise=True))
    keys = list(m.state_dict())
    assert "backbone.backbone.dark2.0.dconv.conv.weight" in keys and "backbone.jian1.pconv.bn.running_var" in keys
    assert "head.cls_convs.0.0.dconv.bn.weight" in keys and m.backbone.backbone.dark3[0].dconv.conv.groups == 16
    m.load_state_dict(synth.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}))
    m.head.use_l1 = True
    x = synth.synth_frames(2, 96, 128)
    tg = synth.synth_labels(2, 96, 128)
    m.train()
    with torch.no_grad():
        loss = m(x, tg)
    assert all(torch.isfinite(v) for v in loss.values())
    with pytest.raises(NotImplementedError):
        m(x, tg)                                           # gradients enabled: the backward does not cover depthwise layers
    m.eval()
    with torch.no_grad():
        out = m(x)
    assert out.shape == (2, 12 * 16 + 6 * 8 + 3 * 4, 13) and torch.isfinite(out).all()


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", [(2, 64, 75, 120, 3, 1), (2, 128, 38, 60, 3, 2), (3, 24, 19, 31, 5, 1), (2, 16, 15, 20, 1, 1),
                                  (1, 256, 150, 240, 3, 2), (2, 48, 37, 59, 5, 2)], ids=lambda c: "x".join(map(str, c)))
def test_dwconv_kernel_vs_torch(case):
    from streamyolo_b200 import ops
    from streamyolo_b200.ops import View
    from test_gpu_ops import check_close
    n, c, h, w, k, s = case
    g = torch.Generator().manual_seed(1)
    x = torch.randn((n, c, h, w), generator=g).to(torch.bfloat16).float().cuda()
    wt = (torch.randn((c, 1, k, k), generator=g) / k).to(torch.bfloat16).float().cuda()
    ref = F.conv2d(x, wt, None, s, (k - 1) // 2, groups=c)
    ho, wo = ref.shape[2], ref.shape[3]
    xv = ops.from_nchw(x)
    wpk = ops.pack_dw_weight(wt)
    assert torch.equal(wpk.float(), wt.reshape(c, k * k).t())
    y = View.empty(n, ho, wo, c, "cuda")
    y.buf.fill_(float("nan"))
    ops.conv2d(xv, wpk, y, k, s, ops.SY_CONV_RAW, impl="dw")
    torch.cuda.synchronize()
    check_close(y.nchw_float(), ref, f"dwconv raw {case}")
    scale = (torch.rand(c, generator=g) + 0.5).cuda()
    shift = (torch.rand(c, generator=g) - 0.5).cuda()
    res = torch.randn((n, c, ho, wo), generator=g).to(torch.bfloat16).float().cuda()
    y2 = View.empty(n, ho, wo, c, "cuda")
    ops.conv2d(xv, wpk, y2, k, s, ops.SY_CONV_FUSED, impl="dw", scale=scale, shift=shift, act=1, res=ops.from_nchw(res))
    torch.cuda.synchronize()
    want = F.silu(ref * scale[None, :, None, None] + shift[None, :, None, None]) + res
    check_close(y2.nchw_float(), want, f"dwconv fused {case}")


@pytest.mark.gpu
@pytest.mark.parametrize("stride", [1, 2])
def test_dwconv_module_vs_reference_block_gpu(stride):
    rb = _ref_blocks()
    ref, prod = rb.DWConv(64, 96, 3, stride), nb.DWConv(64, 96, 3, stride)
    _sync(ref, prod)
    with torch.no_grad():                                  # bf16-representable conv weights on both sides (the product stores them as bf16)
        for mod in (ref, prod):
            for m in mod.modules():
                if isinstance(m, torch.nn.Conv2d):
                    m.weight.copy_(m.weight.to(torch.bfloat16).float())
    prod.cuda()
    x = torch.randn(4, 64, 38, 60).to(torch.bfloat16).float()
    for train in (True, False):
        ref.train(train)
        prod.train(train)
        with torch.no_grad():
            want = ref(x.clone())
            got = prod(x.cuda()).float().cpu()
        rms = float(want.pow(2).mean().sqrt())
        err = (got - want).abs()
        # two bf16-stored layers deep against an fp32-storage reference: the depthwise output is rounded to bf16 (2^-9 per
        # element) before BatchNorm and the 64-term pointwise sum: relative l2 well below 1 %, no element off by more than a
        # few bf16 ulp of the tensor's scale
        assert float(err.pow(2).mean().sqrt()) <= 6e-3 * rms, (train, float(err.pow(2).mean().sqrt()), rms)
        assert bool((err <= 2.0 ** -5 * want.abs() + 2.0 ** -5 * rms).all()), (train, float(err.max()), rms)
    assert torch.allclose(prod.dconv.bn.running_mean.cpu(), ref.dconv.bn.running_mean, rtol=2e-3, atol=2e-4)
    assert torch.allclose(prod.pconv.bn.running_var.cpu(), ref.pconv.bn.running_var, rtol=5e-3, atol=5e-4)
