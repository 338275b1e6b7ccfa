"""CPU: the oracle restatement vs fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  This is what pins oracle/streamyolo_oracle.py."""
import os

import numpy as np
import pytest
import torch

from oracle.streamyolo_oracle import OracleCfg, StreamYoloOracle, model_shapes, conv_gflop_per_pair
from oracle.make_golden import CASES
from streamyolo_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def stat3(t):
    t = t.detach().double()
    return np.array([t.mean().item(), t.abs().mean().item(), t.pow(2).mean().sqrt().item()])


@pytest.mark.parametrize("tag,dw", [("s", (0.33, 0.5)), ("m", (0.67, 0.75)), ("l", (1.0, 1.0))])
def test_state_dict_inventory(tag, dw):
    g = np.load(os.path.join(GOLD, "state_shapes.npz"))
    shapes = model_shapes(*dw)
    assert list(shapes) != [] and set(shapes) == set(g[tag + "_keys"].tolist())
    ref = dict(zip(g[tag + "_keys"].tolist(), g[tag + "_shapes"].tolist()))
    for k, s in shapes.items():
        assert "x".join(map(str, s)) == ref[k], k
    n = sum(int(np.prod(s)) for k, s in shapes.items()
            if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    assert n == int(g[tag + "_nparams"])


def test_flop_table():
    # 61.43 / 176.81 / 384.30 GFLOP per pair
    for (d, w), ref in {(0.33, 0.5): 61.43, (0.67, 0.75): 176.81, (1.0, 1.0): 384.30}.items():
        assert abs(conv_gflop_per_pair(d, w) - ref) / ref < 2e-3


def _oracle(c, momentum=0.03):
    cfg = OracleCfg(depth=c["depth"], width=c["width"], gamma=c["gamma"], ignore_thr=c["thr"],
                    ignore_value=c["val"], bn_momentum=momentum)
    st = synth.synth_state_dict(model_shapes(c["depth"], c["width"]))
    return StreamYoloOracle(cfg, st)


@pytest.mark.parametrize("name", ["tiny_120x160", "tiny_empty_96x160", "s_600x960"])
def test_oracle_matches_reference(name):
    c = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    x = synth.synth_frames(c["B"], c["H"], c["W"])
    tg = synth.synth_labels(c["B"], c["H"], c["W"], empty_image=c["empty"])
    o = _oracle(c)
    o.trace = {}
    feats = o.backbone_off(x)
    outputs, origin, grid = o.flatten_decode(o.head_levels(feats), sigmoid=False)
    res = o.losses(outputs, origin, grid, tg, return_aux=True)
    order = ["total_loss", "iou_loss", "l1_loss", "conf_loss", "cls_loss", "num_fg"]
    got = np.array([float(res[k]) for k in order])
    np.testing.assert_allclose(got, g["train_loss"], rtol=2e-4, atol=1e-5)
    # assignment: bit-exact integer results
    aux = res["aux"]
    bi, ai = aux["fg"].nonzero(as_tuple=True)
    assert np.array_equal(bi.numpy().astype(np.int32), g["fg_image"])
    assert np.array_equal(ai.numpy().astype(np.int32), g["fg_anchor"])
    assert np.array_equal(aux["matched"][bi, ai].numpy().astype(np.int32), g["fg_gt"])
    np.testing.assert_allclose(aux["pred_iou"][bi, ai].numpy(), g["fg_iou"], rtol=1e-4, atol=1e-6)
    # BN running statistics after one step
    for k, ref in zip(g["bn_keys"].tolist(), g["bn_stats_after_train"]):
        np.testing.assert_allclose(stat3(o.P[k]), ref, rtol=1e-4, atol=1e-6, err_msg=k)
    assert int(o.P["backbone.backbone.stem.conv.bn.num_batches_tracked"]) == g["nbt"][0] == 2
    assert int(o.P["backbone.jian2.bn.num_batches_tracked"]) == g["nbt"][1] == 2
    assert int(o.P["head.stems.0.bn.num_batches_tracked"]) == g["nbt"][2] == 1
    # per-BaseConv output statistics (last call = support frame for shared layers)
    for k, ref in zip(g["conv_keys"].tolist(), g["conv_stats_train"]):
        if k.startswith("backbone.jian"):
            continue  # oracle traces jian before the fuse; covered by the fused outputs
        np.testing.assert_allclose(stat3(o.trace[k]), ref, rtol=2e-4, atol=1e-6, err_msg=k)


@pytest.mark.parametrize("name", ["tiny_120x160", "s_600x960"])
def test_oracle_eval_and_on_pipe(name):
    c = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    x = synth.synth_frames(c["B"], c["H"], c["W"])
    tg = synth.synth_labels(c["B"], c["H"], c["W"], empty_image=c["empty"])
    o = _oracle(c, momentum=1.0)
    xc = torch.cat([x[:, 0:3], x[:, 0:3]], 1)
    o.forward(xc, tg)
    o.training = False
    ev = o.forward(xc)
    assert [list(h) for h in o.hw] == g["eval_hw"].tolist()
    sub = int(g["eval_sub_step"])
    np.testing.assert_allclose(ev[:, ::sub].numpy(), g["eval_sub"], rtol=2e-3, atol=2e-3)
    o1, buf = o.forward(x[:1, 0:3], mode="on_pipe")
    o2, buf2 = o.forward(x[1:2, 0:3], buffer=buf, mode="on_pipe")
    got = np.stack([stat3(o1), stat3(o2)] + [stat3(b) for b in buf2])
    np.testing.assert_allclose(got, g["on_pipe_stats"], rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(o2[:, ::sub].numpy(), g["on_pipe_sub2"], rtol=5e-3, atol=5e-3)


def test_on_pipe_star_equals_buffer_on_same_frame():
    # first call and a second call with the returned buffer on the SAME frame agree in eval mode
    c = CASES["tiny_120x160"]
    o = _oracle(c)
    o.training = False
    x = synth.synth_frames(1, 96, 160)[:, :3]
    a, buf = o.forward(x, mode="on_pipe")
    b, _ = o.forward(x, buffer=buf, mode="on_pipe")
    assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["tiny_120x160", "tiny_empty_96x160"])
def test_oracle_backward_matches_reference(name):
    """Autograd through the oracle (fp32) against loss.backward() of the unmodified reference: every parameter's
    gradient norm, and the head prediction-conv bias gradients element-wise (they are the per-channel sums of
    d loss / d raw head output, the quantity the loss-backward kernel produces).  This pins the oracle as the gradient
    reference for the backward path."""
    c = CASES[name]
    g = np.load(os.path.join(GOLD, "grad_" + name + ".npz"))
    o = _oracle(c)
    for k, t in o.P.items():
        if t.dtype.is_floating_point and not k.endswith(("running_mean", "running_var")):
            t.requires_grad_(True)
    x = synth.synth_frames(c["B"], c["H"], c["W"])
    tg = synth.synth_labels(c["B"], c["H"], c["W"], empty_image=c["empty"])
    loss = o.forward(x, tg)["total_loss"]
    assert abs(float(loss) - float(g["total_loss"])) <= 2e-4 * abs(float(g["total_loss"]))
    loss.backward()
    keys = g["grad_keys"].tolist()
    assert set(keys) == {k for k, t in o.P.items() if t.grad is not None}
    l2 = dict(zip(keys, g["grad_l2"].tolist()))
    worst = 0.0
    for k in keys:
        got = float(o.P[k].grad.norm())
        worst = max(worst, abs(got - l2[k]) / (l2[k] + 1e-6))
    assert worst < 5e-3, f"worst relative gradient-norm error {worst:.2e}"
    for f in g.files:
        if f.startswith("g:"):
            ref = torch.from_numpy(g[f])
            got = o.P[f[2:]].grad
            assert torch.allclose(got, ref, rtol=2e-3, atol=2e-5 * float(ref.abs().max()) + 1e-7), f
