"""CPU: the host-side routing of the training backward (streamyolo_b200/model/backward.py) with every kernel call
replaced by its torch emulation (tests/emul_ops.py): all parameter gradients against autograd through the oracle with the
same bf16 storage points (which tests/test_oracle_golden.py pins to the reference's loss.backward()).
The kernels themselves are tested on the GPU (tests/test_gpu_ops.py, tests/test_gpu_model.py)."""
import numpy as np
import pytest
import torch

from oracle.make_golden import CASES
from oracle.streamyolo_oracle import OracleCfg, StreamYoloOracle, bf16_round, model_shapes
from streamyolo_b200 import synth
from streamyolo_b200.model import DFPPAFPN, TALHead, YOLOX, backward

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emul_ops  # noqa: E402


def build_product(c):
    ch = [256, 512, 1024]
    m = YOLOX(DFPPAFPN(c["depth"], c["width"], in_channels=ch),
              TALHead(8, c["width"], in_channels=ch, gamma=c["gamma"], ignore_thr=c["thr"], ignore_value=c["val"]))
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.eps, mod.momentum = 1e-3, 0.03
    m.head.initialize_biases(1e-2)
    m.load_state_dict(synth.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}), strict=True)
    m.head.use_l1 = True
    return m.train()


def _run(c, monkeypatch, exact):
    emul_ops.install(monkeypatch, exact=exact)
    x = synth.synth_frames(c["B"], c["H"], c["W"])
    tg = synth.synth_labels(c["B"], c["H"], c["W"], empty_image=c["empty"])
    model = build_product(c)
    loss = backward.forward_backward(model, x, tg)
    cfg = OracleCfg(depth=c["depth"], width=c["width"], gamma=c["gamma"], ignore_thr=c["thr"], ignore_value=c["val"])
    o = StreamYoloOracle(cfg, synth.synth_state_dict(model_shapes(c["depth"], c["width"])), q=None if exact else bf16_round)
    for k, t in o.P.items():
        if t.dtype.is_floating_point and not k.endswith(("running_mean", "running_var")):
            t.requires_grad_(True)
    ref = o.forward(x, tg)
    ref["total_loss"].backward()
    params = dict(model.named_parameters())
    assert set(params) == {k for k, t in o.P.items() if t.grad is not None}
    report = []
    for k, p in params.items():
        assert p.grad is not None and torch.isfinite(p.grad).all(), f"no finite gradient reached {k}"
        g, r = p.grad.float().flatten(), o.P[k].grad.float().flatten()
        report.append((float((g - r).norm() / (r.norm() + 1e-12)), float(torch.dot(g, r) / (g.norm() * r.norm() + 1e-20)), k))
    report.sort(reverse=True)
    return float(loss["total_loss"]), float(ref["total_loss"].detach()), report, model, o


@pytest.mark.parametrize("name", ["tiny_120x160", "tiny_empty_96x160"])
def test_backward_routing_exact(name, monkeypatch):
    """fp32 storage everywhere (no bf16 rounding): the routing of the backward walk must reproduce autograd through the
    fp32 oracle -- i.e. the reference's loss.backward() -- for all 249 parameters to float roundoff."""
    got, want, report, model, o = _run(CASES[name], monkeypatch, exact=True)
    assert abs(got - want) < 1e-5 * abs(want)
    msg = "\n".join(f"{rel:8.3e} cos {cos:.5f} {k}" for rel, cos, k in report[:10])
    assert report[0][0] < 1e-3, "largest deviations:\n" + msg
    # the BatchNorm buffers follow the reference too (two statistic groups = two updates, current frames first)
    sd = model.state_dict()
    for k in ("backbone.backbone.stem.conv.bn.running_mean", "backbone.C3_n4.conv3.bn.running_var", "head.stems.1.bn.running_mean",
              "backbone.jian1.bn.running_var"):
        assert torch.allclose(sd[k], o.P[k].detach(), rtol=1e-4, atol=1e-6), k
    assert int(sd["backbone.backbone.dark3.0.bn.num_batches_tracked"]) == 2 and int(sd["backbone.jian0.bn.num_batches_tracked"]) == 2


def test_backward_routing_bf16_storage(monkeypatch):
    """Same walk with the product's bf16 storage points emulated: a random-init train-mode BatchNorm net amplifies the
    rounding noise, so this only checks that every gradient is finite, of the right magnitude and
    pointing the right way."""
    got, want, report, _, _ = _run(CASES["tiny_120x160"], monkeypatch, exact=False)
    assert abs(got - want) < 2e-2 * abs(want)
    cos = sorted(c_ for _, c_, _ in report)
    assert cos[len(cos) // 2] > 0.8, f"median cosine {cos[len(cos) // 2]:.3f}"
