"""GPU: the product model (streamyolo_b200.model, CUDA kernels through the C ABI) against the CPU
oracle on identical synthetic weights / frames / labels, and against the reference-generated
golden fixtures.

Tolerances
  * oracle run with the product's storage rounding (q = bf16 round trip): feature maps
    ||a-b||_2 / ||b||_2 <= 1e-2 (bf16 rounding flips through ~100 layers), losses 1e-2 relative,
    BN running statistics 2e-3.
  * SimOTA/TAL given IDENTICAL fp32 head outputs: foreground set, matched GT ids bit exact,
    matched IoUs / loss values 1e-4 (north_star: integer indexing bit-exact).
  * golden fixtures (reference in fp32): losses within 5e-2 relative (bf16 activations vs fp32).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.make_golden import CASES  # noqa: E402
from oracle.streamyolo_oracle import OracleCfg, StreamYoloOracle, bf16_round, model_shapes  # noqa: E402
from streamyolo_b200 import ops, synth  # noqa: E402
from streamyolo_b200.model import DFPPAFPN, TALHead, YOLOX  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ORDER = ["total_loss", "iou_loss", "l1_loss", "conf_loss", "cls_loss", "num_fg"]


def build_product(depth, width, gamma=1.0, thr=0.5, val=1.5, momentum=0.03):
    ch = [256, 512, 1024]
    m = YOLOX(DFPPAFPN(depth, width, in_channels=ch),
              TALHead(8, width, in_channels=ch, gamma=gamma, ignore_thr=thr, ignore_value=val))
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.eps, mod.momentum = 1e-3, momentum
    m.head.initialize_biases(1e-2)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(synth.synth_state_dict(shapes), strict=True)
    m.head.use_l1 = True
    return m.cuda()


def build_oracle(depth, width, gamma=1.0, thr=0.5, val=1.5, momentum=0.03, q=bf16_round):
    cfg = OracleCfg(depth=depth, width=width, gamma=gamma, ignore_thr=thr, ignore_value=val, bn_momentum=momentum)
    return StreamYoloOracle(cfg, synth.synth_state_dict(model_shapes(depth, width)), q=q)


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.fixture(params=["tc", "simt"])
def impl(request, monkeypatch):
    monkeypatch.setenv("SY_CONV_IMPL", request.param)
    return request.param


@pytest.mark.parametrize("case", ["tiny_120x160", "tiny_empty_96x160"])
def test_train_forward_vs_oracle(case, impl):
    c = CASES[case]
    x = synth.synth_frames(c["B"], c["H"], c["W"])
    tg = synth.synth_labels(c["B"], c["H"], c["W"], empty_image=c["empty"])
    m = build_product(c["depth"], c["width"], c["gamma"], c["thr"], c["val"])
    m.train()
    feats = m.backbone(x.cuda())
    o = build_oracle(c["depth"], c["width"], c["gamma"], c["thr"], c["val"])
    o.trace = {}
    ofeats = o.backbone_off(x)
    for name, a, b in zip(("jian2", "jian1", "jian0"), feats, ofeats):
        r = rel(a, b)
        assert r < 1e-2, f"fused {name}: rel l2 {r}"
    # running statistics after the (two-group) step
    sd = m.state_dict()
    for k in ("backbone.backbone.stem.conv.bn.running_mean", "backbone.backbone.dark3.1.m.0.conv2.bn.running_var",
              "backbone.C3_n4.conv3.bn.running_var", "backbone.jian1.bn.running_mean"):
        assert torch.allclose(sd[k].cpu(), o.P[k], rtol=2e-3, atol=2e-4), k
    assert int(sd["backbone.backbone.stem.conv.bn.num_batches_tracked"]) == 2
    # full forward + loss
    m2 = build_product(c["depth"], c["width"], c["gamma"], c["thr"], c["val"])
    m2.train()
    loss = m2(x.cuda(), (tg[0].cuda(), tg[1].cuda()))
    torch.cuda.synchronize()
    o2 = build_oracle(c["depth"], c["width"], c["gamma"], c["thr"], c["val"])
    ref = o2.forward(x, tg)
    got = np.array([float(loss[k]) for k in ORDER])
    want = np.array([float(ref[k]) for k in ORDER])
    np.testing.assert_allclose(got, want, rtol=1e-2, atol=1e-3)
    gold = np.load(os.path.join(GOLD, case + ".npz"))["train_loss"]
    np.testing.assert_allclose(got, gold, rtol=5e-2, atol=5e-3)


def test_loss_kernels_bit_exact_assignment():
    """Feed the oracle's own fp32 head outputs to sy_tal_loss: integer results must be identical."""
    c = CASES["tiny_120x160"]
    x = synth.synth_frames(4, c["H"], c["W"])
    tg = synth.synth_labels(4, c["H"], c["W"], empty_image=2)
    o = build_oracle(c["depth"], c["width"], q=None)
    feats = o.backbone_off(x)
    outputs, origin, grid = o.flatten_decode(o.head_levels(feats), sigmoid=False)
    ref = o.losses(outputs, origin, grid, tg, return_aux=True)
    b, a, no = outputs.shape
    dev = "cuda"
    ws = torch.empty(ops.tal_loss_workspace_bytes(b, a, 120, 8), dtype=torch.uint8, device=dev)
    loss = torch.empty(6, device=dev)
    fg = torch.empty((b, a), dtype=torch.int32, device=dev)
    mt = torch.empty((b, a), dtype=torch.int32, device=dev)
    pi = torch.empty((b, a), device=dev)
    ops.tal_loss(outputs.cuda().contiguous(), origin.cuda().contiguous(), tg[0].cuda(), tg[1].cuda(), o.hw,
                 (8, 16, 32), 1.0, 0.5, 1.5, True, ws, loss, fg, mt, pi)
    torch.cuda.synchronize()
    aux = ref["aux"]
    assert torch.equal(fg.cpu().bool(), aux["fg"]), "foreground anchor set differs"
    assert torch.equal(mt.cpu().long(), aux["matched"]), "matched GT ids differ"
    assert torch.allclose(pi.cpu(), aux["pred_iou"], rtol=1e-5, atol=1e-6)
    got = loss.cpu().numpy()
    want = np.array([float(ref[k]) for k in ORDER])[[0, 1, 3, 4, 2, 5]]   # kernel order: total, iou, obj, cls, l1, num_fg
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)


def test_eval_and_on_pipe_vs_oracle(impl):
    c = CASES["tiny_120x160"]
    x = synth.synth_frames(c["B"], c["H"], c["W"])
    tg = synth.synth_labels(c["B"], c["H"], c["W"])
    xc = torch.cat([x[:, 0:3], x[:, 0:3]], 1)
    o = build_oracle(c["depth"], c["width"], momentum=1.0)
    o.forward(xc, tg)                       # calibration pass: running stats := batch stats
    o.training = False
    ref = o.forward(xc)
    m = build_product(c["depth"], c["width"])
    # drop-in property: the oracle's (== reference's) state_dict loads into the product
    m.load_state_dict({k: v.clone() for k, v in o.P.items()}, strict=True)
    m.eval()
    got = m(xc.cuda())
    torch.cuda.synchronize()
    assert list(map(tuple, m.head.hw)) == [tuple(h) for h in o.hw]
    assert got.shape == ref.shape
    r = rel(got[..., :4], ref[..., :4])
    assert r < 2e-2, f"eval boxes rel l2 {r}"
    assert (got[..., 4:].cpu() - ref[..., 4:]).abs().max().item() < 2e-2
    gold = np.load(os.path.join(GOLD, "tiny_120x160.npz"))
    sub = int(gold["eval_sub_step"])
    g = torch.from_numpy(gold["eval_sub"])
    assert rel(got[:, ::sub, :4], g[..., :4]) < 5e-2
    # on_pipe: star call == buffered call on the same frame; buffered call on a new frame vs oracle
    o1, buf = m(x[:1, 0:3].cuda(), mode="on_pipe")
    o1b, _ = m(x[:1, 0:3].cuda(), buffer=buf, mode="on_pipe")
    assert torch.equal(o1, o1b)
    o2, buf2 = m(x[1:2, 0:3].cuda(), buffer=buf, mode="on_pipe")
    r1, rbuf = o.forward(x[:1, 0:3], mode="on_pipe")
    r2, _ = o.forward(x[1:2, 0:3], buffer=rbuf, mode="on_pipe")
    torch.cuda.synchronize()
    assert rel(o1[..., :4], r1[..., :4]) < 2e-2 and rel(o2[..., :4], r2[..., :4]) < 2e-2
    for a, b in zip(buf, rbuf):
        assert rel(a, b) < 1e-2


def test_s_model_full_resolution_golden():
    """StreamYOLO-s, 600x960, B=2 (BASELINE.json config 1) against the reference-generated fixture."""
    c = CASES["s_600x960"]
    x = synth.synth_frames(c["B"], c["H"], c["W"])
    tg = synth.synth_labels(c["B"], c["H"], c["W"])
    m = build_product(c["depth"], c["width"])
    m.train()
    m.head.keep_assignment = True
    loss = m(x.cuda(), (tg[0].cuda(), tg[1].cuda()))
    torch.cuda.synchronize()
    got = np.array([float(loss[k]) for k in ORDER])
    gold = np.load(os.path.join(GOLD, "s_600x960.npz"))
    np.testing.assert_allclose(got, gold["train_loss"], rtol=5e-2, atol=5e-3)
    assert m.head.hw == [(75, 120), (38, 60), (19, 30)]
    # size-independent property at full size: every foreground anchor is a candidate of its matched GT
    asg = m.head.last_assignment
    fg = asg["fg_out"].bool()
    assert int(fg.sum()) >= 1 and (asg["matched_out"][fg] >= 0).all() and (asg["matched_out"][~fg] == -1).all()
    assert (asg["pred_iou_out"][fg] >= 0).all() and (asg["pred_iou_out"][fg] <= 1).all()
