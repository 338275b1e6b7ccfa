"""Is the deviation of the assembled backward (GPU) from autograd through the oracle rounding noise or a routing bug?

For every parameter, in walk order: rel(product, oracle) next to three noise floors --
  oracle(x * (1 + 1e-6)) vs oracle(x), oracle(x * (1 + 2^-9)) vs oracle(x) (half a bf16 ulp on every input: every stored value
  rounds differently somewhere), product(x * (1 + 2^-9)) vs product(x) (the product against itself).
A routing bug shows as product-vs-oracle far above product-vs-product; chaos shows all of them large together.

    python tests/tools/diag_bwd.py [case] [H W]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from oracle.make_golden import CASES
from oracle.streamyolo_oracle import OracleCfg, StreamYoloOracle, bf16_round, model_shapes
from streamyolo_b200 import synth, train
from streamyolo_b200.model import DFPPAFPN, TALHead, YOLOX, backward


def main():
    c = dict(CASES[sys.argv[1] if len(sys.argv) > 1 else "tiny_120x160"])
    if len(sys.argv) > 3:
        c["H"], c["W"] = int(sys.argv[2]), int(sys.argv[3])
    x = synth.synth_frames(c["B"], c["H"], c["W"])
    tg = synth.synth_labels(c["B"], c["H"], c["W"], empty_image=c["empty"])

    def product(xin):
        ch = [256, 512, 1024]
        m = YOLOX(DFPPAFPN(c["depth"], c["width"], in_channels=ch),
                  TALHead(8, c["width"], in_channels=ch, gamma=c["gamma"], ignore_thr=c["thr"], ignore_value=c["val"]))
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.eps, mod.momentum = 1e-3, 0.03
        m.head.initialize_biases(1e-2)
        m.load_state_dict(synth.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}), strict=True)
        m.head.use_l1 = True
        m.cuda().train()
        loss = backward.forward_backward(m, xin.cuda(), (tg[0].cuda(), tg[1].cuda()))
        torch.cuda.synchronize()
        return m, float(loss["total_loss"]), {k: p.grad.float().cpu() for k, p in m.named_parameters()}

    def oracle(xin):
        cfg = OracleCfg(depth=c["depth"], width=c["width"], gamma=c["gamma"], ignore_thr=c["thr"], ignore_value=c["val"])
        o = StreamYoloOracle(cfg, synth.synth_state_dict(model_shapes(c["depth"], c["width"])), q=bf16_round)
        for k, t in o.P.items():
            if t.dtype.is_floating_point and not k.endswith(("running_mean", "running_var")):
                t.requires_grad_(True)
        r = o.forward(xin, tg)
        r["total_loss"].backward()
        return float(r["total_loss"].detach()), {k: t.grad for k, t in o.P.items() if t.grad is not None}

    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-20))

    def cos(a, b):
        return float(torch.dot(a.flatten(), b.flatten()) / (a.norm() * b.norm() + 1e-20))

    m, pl, pg = product(x)
    _, pl2, pg2 = product(x * (1 + 2.0 ** -9))
    ol, og = oracle(x)
    _, og6 = oracle(x * (1 + 1e-6))
    _, og9 = oracle(x * (1 + 2.0 ** -9))
    print(f"{c['H']}x{c['W']} loss product {pl:.5f} product(nudged) {pl2:.5f} oracle {ol:.5f}")
    order = []
    for g in reversed(train.conv_groups_forward_order(m)):
        for mod in g:
            order += [mod.conv.weight, mod.bn.weight, mod.bn.bias]
    names = {id(p): k for k, p in m.named_parameters()}
    keys = [names[id(p)] for p in order] + [k for k in pg if "_preds" in k]
    print(f"{'parameter':52s} {'p-vs-o':>8s} {'cos':>7s} {'o 1e-6':>8s} {'o 2^-9':>8s} {'p 2^-9':>8s} {'|p|/|o|':>8s}")
    for k in keys:
        print(f"{k:52s} {rel(pg[k], og[k]):8.3f} {cos(pg[k], og[k]):7.3f} {rel(og6[k], og[k]):8.3f} {rel(og9[k], og[k]):8.3f} "
              f"{rel(pg2[k], pg[k]):8.3f} {float(pg[k].norm() / (og[k].norm() + 1e-20)):8.3f}")


if __name__ == "__main__":
    main()
