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
    # the activation-gradient arena starts as NaN instead of uninitialised memory: a region the walk reads before any
    # consumer wrote it (first-write bookkeeping of backward.Tape) would poison the parameter gradients checked below
    monkeypatch.setattr(backward, "POISON", True)
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


# deeper variants of the same graph (2 and 3 bottlenecks per CSP stage, 6 and 9 in dark3/dark4: StreamYOLO-m / -l depth)
DEPTH_CASES = {
    "m_depth_96x128": dict(depth=0.67, width=0.125, H=96, W=128, B=2, gamma=1.0, thr=0.5, val=1.5, empty=-1),
    "l_depth_64x96": dict(depth=1.0, width=0.125, H=64, W=96, B=2, gamma=1.5, thr=0.4, val=1.7, empty=-1),
}


@pytest.mark.parametrize("name", list(DEPTH_CASES))
def test_backward_routing_exact_deeper_models(name, monkeypatch):
    got, want, report, _, _ = _run(DEPTH_CASES[name], monkeypatch, exact=True)
    assert abs(got - want) < 1e-5 * abs(want)
    msg = "\n".join(f"{rel:8.3e} cos {cos:.5f} {k}" for rel, cos, k in report[:10])
    # deeper tiny nets amplify float roundoff a little more (a routing mistake would be O(1), with cosine far from 1)
    assert report[0][0] < 1e-2 and min(c_ for _, c_, _ in report) > 0.9999, "largest deviations:\n" + msg


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


def test_train_steps_reduce_the_loss(monkeypatch):
    """The reference's optimiser setup (SGD nesterov, three parameter groups) on top of the backward walk: the parameter
    groups cover every parameter exactly once, and a few steps on a fixed batch bring the loss down."""
    from streamyolo_b200 import train
    c = CASES["tiny_120x160"]
    emul_ops.install(monkeypatch, exact=True)
    model = build_product(c)
    opt = train.build_optimizer(model, lr=2e-4)
    ids = [id(p) for g in opt.param_groups for p in g["params"]]
    assert sorted(ids) == sorted(id(p) for p in model.parameters()) and len(set(ids)) == len(ids)
    assert opt.param_groups[0].get("weight_decay", 0) == 0 and opt.param_groups[1]["weight_decay"] == 5e-4
    ema = train.ModelEMA(model)
    x = synth.synth_frames(c["B"], c["H"], c["W"])
    tg = synth.synth_labels(c["B"], c["H"], c["W"])
    losses = [float(train.train_step(model, opt, x, tg, ema)["total_loss"]) for _ in range(4)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    assert ema.updates == 4


def test_loss_backward_through_autograd(monkeypatch):
    """Opt-in drop-in for the reference trainer (double_trainer.py:105-123): with ``model.train_with_autograd`` the training
    forward returns a loss with a grad_fn; ``(scale * loss).backward()`` must hand autograd exactly the gradients of the
    explicit walk, times the scale (GradScaler semantics), and accumulate over two calls like autograd does."""
    c = CASES["tiny_120x160"]
    emul_ops.install(monkeypatch, exact=True)
    x = synth.synth_frames(c["B"], c["H"], c["W"])
    tg = synth.synth_labels(c["B"], c["H"], c["W"])
    ref_model = build_product(c)
    want_loss = backward.forward_backward(ref_model, x, tg)
    model = build_product(c)
    model.train_with_autograd = True
    out = model(x, tg)
    assert out["total_loss"].requires_grad and not out["iou_loss"].requires_grad
    assert float(out["total_loss"]) == float(want_loss["total_loss"])
    (out["total_loss"] * 64.0).backward()
    for (k, p), q in zip(model.named_parameters(), ref_model.parameters()):
        assert p.grad is not None, k
        assert torch.allclose(p.grad, 64.0 * q.grad, rtol=1e-5, atol=1e-6 * float(q.grad.abs().max()) + 1e-12), k
    with torch.no_grad():                                    # evaluation-style call inside training mode: the plain forward
        plain = model(x, tg)
    assert not plain["total_loss"].requires_grad


DDP_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch
import emul_ops, test_cpu_backward as T
from oracle.make_golden import CASES
from streamyolo_b200 import dist as d, synth
from streamyolo_b200.model import backward


class MP:
    def setattr(self, o, n, v): setattr(o, n, v)
    def setitem(self, dct, k, v): dct[k] = v


rank, local, world = d.init("gloo")
emul_ops.install(MP(), exact=True)
c = CASES["tiny_120x160"]
x = synth.synth_frames(4, c["H"], c["W"])
fut, cur = synth.synth_labels(4, c["H"], c["W"])


def grads(lo, hi):
    m = T.build_product(c)
    backward.forward_backward(m, x[lo:hi], (fut[lo:hi], cur[lo:hi]))
    return m


lo, hi = d.shard_pairs(4, world, rank)
mine = grads(lo, hi)                                    # this rank's shard
nb = d.allreduce_grads(mine.parameters(), bucket_bytes=64 << 10)
assert nb > 1, nb
a, b = grads(0, 2), grads(2, 4)                         # both shards in one process: the expected mean
for (k, p), pa, pb in zip(mine.named_parameters(), a.parameters(), b.parameters()):
    want = 0.5 * (pa.grad + pb.grad)
    assert torch.allclose(p.grad, want, rtol=1e-5, atol=1e-7 * float(want.abs().max()) + 1e-12), k
print("ok", rank)
"""


def test_ddp_gradient_allreduce_gloo_world2(tmp_path):
    """Row a19, host side: two processes (gloo), each runs the training backward on its shard of the global batch, then
    the bucketed all-reduce; every parameter gradient must equal the mean of the two shards' gradients (DDP semantics:
    BatchNorm statistics and the loss normaliser stay per rank)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "ddp_w.py"
    script.write_text(DDP_WORKER)
    port = 29900 + os.getpid() % 90
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="4")
        procs.append(subprocess.Popen([sys.executable, str(script), root], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-2000:] for o in outs)


@pytest.mark.parametrize("name", ["tiny_120x160", "tiny_empty_96x160"])
def test_forward_engine_routing_exact(name, monkeypatch):
    """The product's ordinary forward (engine.py: in-place concat slices, the two frames batched with grouped statistics,
    the DFP fusion with group-offset destinations, eval with folded BN, on_pipe with a buffer) with emulated kernels and
    fp32 storage, against the fp32 oracle: losses, eval outputs and streaming outputs to float roundoff."""
    c = CASES[name]
    emul_ops.install(monkeypatch, exact=True)
    x = synth.synth_frames(c["B"], c["H"], c["W"])
    tg = synth.synth_labels(c["B"], c["H"], c["W"], empty_image=c["empty"])
    cfg = OracleCfg(depth=c["depth"], width=c["width"], gamma=c["gamma"], ignore_thr=c["thr"], ignore_value=c["val"])
    o = StreamYoloOracle(cfg, synth.synth_state_dict(model_shapes(c["depth"], c["width"])), q=None)
    model = build_product(c)
    with torch.no_grad():                                    # the plain forward (engine.py), not the recording one
        loss = model(x, tg)
    ref = o.forward(x, tg)
    for k in ("total_loss", "iou_loss", "l1_loss", "conf_loss", "cls_loss", "num_fg"):
        assert abs(float(loss[k]) - float(ref[k])) <= 2e-5 * abs(float(ref[k])) + 1e-6, k
    sd = model.state_dict()
    for k in ("backbone.backbone.stem.conv.bn.running_var", "backbone.jian2.bn.running_mean", "head.stems.0.bn.running_var"):
        assert torch.allclose(sd[k], o.P[k], rtol=1e-4, atol=1e-6), k
    model.eval()
    o.training = False
    with torch.no_grad():
        ev, ev_ref = model(x), o.forward(x, mode="off_pipe")
        assert torch.allclose(ev, ev_ref, rtol=2e-4, atol=2e-4), float((ev - ev_ref).abs().max())
        o1, buf = model(x[:1, 0:3], mode="on_pipe")
        o2, _ = model(x[1:2, 0:3], buffer=buf, mode="on_pipe")
        r1, rbuf = o.forward(x[:1, 0:3], mode="on_pipe")
        r2, _ = o.forward(x[1:2, 0:3], buffer=rbuf, mode="on_pipe")
        assert torch.allclose(o1, r1, rtol=2e-4, atol=2e-4) and torch.allclose(o2, r2, rtol=2e-4, atol=2e-4)


def test_half_model_eval_forward(monkeypatch):
    """tools/eval.py --fp16 path: ``model.half()`` and half inputs go through the same kernels (parameters are repacked
    from whatever dtype they have); outputs stay close to the fp32-parameter model (fp16 parameter rounding only)."""
    c = CASES["tiny_120x160"]
    emul_ops.install(monkeypatch, exact=False)
    x = synth.synth_frames(c["B"], c["H"], c["W"])
    m32 = build_product(c).eval()
    m16 = build_product(c).eval().half()
    with torch.no_grad():
        a, b = m32(x), m16(x.half())
    assert b.dtype == torch.float32 and torch.isfinite(b).all()
    rel = float((a - b).norm() / a.norm())
    assert rel < 5e-2, rel


def test_decode_outputs_path(monkeypatch):
    """tools/eval.py:187-188: ``head.decode_in_inference = False`` + ``head.decode_outputs(outputs, dtype)`` must give what the
    in-kernel decode gives."""
    c = CASES["tiny_120x160"]
    emul_ops.install(monkeypatch, exact=True)
    x = synth.synth_frames(c["B"], c["H"], c["W"])
    m = build_product(c).eval()
    with torch.no_grad():
        want = m(x)
        m.head.decode_in_inference = False
        raw = m(x)
        got = m.head.decode_outputs(raw.clone(), dtype=raw.type())
    assert not torch.allclose(raw[..., :4], want[..., :4])
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)


def test_pipe_head_is_tal_head_without_trend_weights(monkeypatch):
    """exps/model/pipe_head.py (cfgs/l_s50_still_dfp_flip.py): one label tensor, no TAL weighting -- the oracle with
    gamma = 0 (constant weight, normalised to 1) on (labels, labels) is the reference's PIPEHead loss."""
    from streamyolo_b200.model import PIPEHead
    c = CASES["tiny_120x160"]
    emul_ops.install(monkeypatch, exact=True)
    x = synth.synth_frames(c["B"], c["H"], c["W"])
    fut, _ = synth.synth_labels(c["B"], c["H"], c["W"])
    ch = [256, 512, 1024]
    m = YOLOX(DFPPAFPN(c["depth"], c["width"], in_channels=ch), PIPEHead(8, c["width"], in_channels=ch))
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.eps, mod.momentum = 1e-3, 0.03
    m.load_state_dict(synth.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}), strict=True)
    m.head.use_l1 = True
    m.train()
    with torch.no_grad():
        loss = m(x, fut)
    cfg = OracleCfg(depth=c["depth"], width=c["width"], gamma=0.0, ignore_thr=0.0, ignore_value=1.0)
    o = StreamYoloOracle(cfg, synth.synth_state_dict(model_shapes(c["depth"], c["width"])), q=None)
    ref = o.forward(x, (fut, fut))
    for k in ("total_loss", "iou_loss", "l1_loss", "conf_loss", "cls_loss", "num_fg"):
        assert abs(float(loss[k]) - float(ref[k])) <= 2e-5 * abs(float(ref[k])) + 1e-6, k
    assert TALHead(80, 0.25).num_classes == 80          # any class count constructs (generic head kernel; tal_head.py:27)
    with pytest.raises(NotImplementedError):
        TALHead(1000)


def test_tape_gradient_regions_first_write_accumulate_and_zero_fill(monkeypatch):
    """The bookkeeping behind the memset-free gradient arena (backward.Tape): first write vs accumulate, reads of regions
    nobody wrote (zero-filled on the spot), partly written regions (completed with zeros, then accumulated), deferred
    shortcut gradients (handed to an exact-match consumer, materialised by any other access).  The network itself never
    takes the zero-fill / partial paths; this drives them directly on a NaN-poisoned arena."""
    from streamyolo_b200.ops import View
    emul_ops.install(monkeypatch, exact=True)
    monkeypatch.setattr(backward, "POISON", True)
    T = backward.Tape(torch.device("cpu"))
    act = View(torch.zeros((4, 3, 5, 16)))
    other = View(torch.zeros((4, 3, 5, 8)))
    T.rec(t="copy", src=act, dst=other)
    T.prepare_grads()
    g = T.g(act)
    assert torch.isnan(g.torch()).all()                          # poisoned, nothing written yet
    one = View(torch.ones((2, 3, 5, 4)))
    # first write to a sub-rectangle, then an accumulate on the same rectangle
    T.accumulate(one, act.imgs(0, 2).ch(0, 4))
    T.accumulate(one, act.imgs(0, 2).ch(0, 4))
    assert (g.torch()[0:2, :, :, 0:4] == 2).all() and torch.isnan(g.torch()[2:]).all()
    # a read of a region that is only partly written: the rest becomes zero, the written part is kept
    r = T.gread(act.imgs(0, 4).ch(0, 8))
    assert (r.torch()[0:2, :, :, 0:4] == 2).all() and (r.torch()[2:4] == 0).all() and (r.torch()[0:2, :, :, 4:8] == 0).all()
    assert torch.isnan(g.torch()[:, :, :, 8:]).all()
    # first() on a partly covered region completes it with zeros and asks for accumulation
    assert T.first(act.imgs(0, 4).ch(4, 8)) is False
    assert (g.torch()[:, :, :, 8:12] == 0).all() and torch.isnan(g.torch()[:, :, :, 12:]).all()
    # a fresh region: first() says "write"
    assert T.first(act.imgs(0, 4).ch(12, 4)) is True
    # deferred shortcut: exact-match consumer takes it, nothing is copied
    T2 = backward.Tape(torch.device("cpu"))
    T2.rec(t="copy", src=act, dst=other)
    T2.prepare_grads()
    src = View(torch.full((4, 3, 5, 8), 3.0))
    T2.defer(src, other)
    assert torch.isnan(T2.g(other).torch()).all()                 # still untouched
    assert T2.take_pending(other) is src and T2.first(other) is True
    # ... and any other access materialises it (copy), a second contribution then accumulates
    T2.defer(src, act.ch(0, 8))
    got = T2.gread(act.ch(0, 4))
    assert (got.torch() == 3).all() and (T2.g(act).torch()[:, :, :, 0:8] == 3).all()
    T2.defer(src, act.ch(0, 8))
    assert (T2.g(act).torch()[:, :, :, 0:8] == 6).all() and not any(T2.pending.values())
