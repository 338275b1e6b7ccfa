"""The training step around ``loss.backward()``: what the reference's trainer loop does per iteration
(/root/reference/exps/train_utils/double_trainer.py:99-123, 171-175), on flat fp32 state.

  * ``build_optimizer`` / ``ModelEMA``: [yolox 0.3.0] Exp.get_optimizer (SGD, momentum 0.9, nesterov, three parameter groups:
    BatchNorm weights and biases without decay, conv / linear weights with weight decay 5e-4) and ModelEMA(model, 0.9998) as
    plain PyTorch -- the drop-in path, used when the reference's own Trainer drives ``model(inps, targets)`` /
    ``loss.backward()`` (the training forward returns a loss with a grad_fn, model/backward.py) -- and the bit-level
    reference of the fused kernel.
  * ``Trainer``: the B200-native step.  Parameters, gradients, momentum and the EMA copy live in FLAT fp32 buffers laid out in
    the order in which the backward walk finishes the gradients; every ``nn.Parameter`` / BatchNorm buffer of the model is a
    view into them (state_dict, checkpoints and ``model.parameters()`` are unchanged).
      - the weight-gradient / BatchNorm-gradient kernels write straight into the flat gradient buffer (``FlatSink``);
      - the buffer is cut into ~25 MB buckets; the moment the walk has enqueued the last gradient of a bucket, its NCCL
        all-reduce is launched on the communication stream and overlaps the rest of the walk (what DistributedDataParallel's
        reducer does for the reference, double_trainer.py:171; ``broadcast_buffers=False``: BatchNorm statistics stay local);
      - ONE launch (``sy_sgd_nesterov_ema_step``) then does unscale (1 / (world x loss scale)) + weight decay + momentum +
        nesterov + parameter update + EMA over the whole state;
      - conv operands are re-packed from the fp32 masters by ``sy_pack_conv_weight`` launches (engine.WEIGHT_EPOCH).
"""
import copy
import math

import torch
import torch.distributed as dist
from torch import nn

from . import dist as sydist
from . import ops
from .model import backward, engine


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
    """The step with stock PyTorch pieces (torch.optim.SGD, Python EMA, post-hoc bucketed all-reduce): the semantics
    reference of ``Trainer.step``."""
    for p in model.parameters():
        p.grad = None
    losses = backward.forward_backward(model, x, targets, grad_scale=grad_scale)
    sydist.allreduce_grads(model.parameters())
    if grad_scale != 1.0:
        for p in model.parameters():
            if p.grad is not None:
                p.grad.div_(grad_scale)
    optimizer.step()
    engine.WEIGHT_EPOCH += 1
    if ema is not None:
        ema.update(model)
    return losses


# ------------------------------------------------------------------------------------------------ flat state
def conv_groups_forward_order(model):
    """The BaseConv launch groups of one training forward in launch order (model/backward.py: pafpn_rec, dfp_rec,
    head_rec): a CSPLayer's conv1 | conv2 run as one GEMM, everything else alone."""
    net, head = model.backbone, model.head
    bb = net.backbone
    out = []

    def bc(m):
        out.append((m,))

    def csp(m):
        out.append((m.conv1, m.conv2))
        for blk in m.m:
            bc(blk.conv1)
            bc(blk.conv2)
        bc(m.conv3)

    bc(bb.stem.conv)
    bc(bb.dark2[0]); csp(bb.dark2[1])
    bc(bb.dark3[0]); csp(bb.dark3[1])
    bc(bb.dark4[0]); csp(bb.dark4[1])
    bc(bb.dark5[0]); bc(bb.dark5[1].conv1); bc(bb.dark5[1].conv2); csp(bb.dark5[2])
    bc(net.lateral_conv0); csp(net.C3_p4); bc(net.reduce_conv1); csp(net.C3_p3)
    bc(net.bu_conv2); csp(net.C3_n3); bc(net.bu_conv1); csp(net.C3_n4)
    bc(net.jian2); bc(net.jian1); bc(net.jian0)
    for k in range(len(head.stems)):
        bc(head.stems[k])
        out.append((head.cls_convs[k][0], head.reg_convs[k][0]))
        bc(head.cls_convs[k][1]); bc(head.reg_convs[k][1])
    return out


_ALIGN = 64     # floats: every segment of the flat buffers starts 256-byte aligned


class FlatState:
    """Flat fp32 buffers holding the model's parameters (walk order), their gradients, momentum, and the EMA copy; re-points
    the module's tensors into them."""

    def __init__(self, model, ema=True):
        dev = next(model.parameters()).device
        groups = list(reversed(conv_groups_forward_order(model)))        # the order the walk finishes them in
        head = model.head
        levels = list(reversed(range(len(head.stems))))
        seg_a, seg_b = [], []                                            # (tensor, slot) lists: no decay | decay
        for k in levels:                                                 # the head prediction convs finish first
            seg_b += [head.reg_preds[k].weight, head.obj_preds[k].weight, head.cls_preds[k].weight]
            seg_a += [head.reg_preds[k].bias, head.obj_preds[k].bias, head.cls_preds[k].bias]
        for g in groups:
            seg_b.append([m.conv.weight for m in g])                     # adjacent: one weight-gradient launch covers the group
            seg_a.append([m.bn.weight for m in g])
            seg_a.append([m.bn.bias for m in g])
        covered = set()
        self.offset = {}                                                 # id(tensor) -> (offset, numel)
        cur = 0

        def place(item):
            nonlocal cur
            cur = (cur + _ALIGN - 1) // _ALIGN * _ALIGN
            for t in (item if isinstance(item, list) else [item]):
                assert id(t) not in covered
                covered.add(id(t))
                self.offset[id(t)] = (cur, t.numel())
                cur += t.numel()

        for it in seg_a:
            place(it)
        cur = (cur + _ALIGN - 1) // _ALIGN * _ALIGN
        self.decay_begin = cur
        self.order_b = []                                                # parameters of the decayed class in walk order
        for it in seg_b:
            place(it)
            self.order_b += it if isinstance(it, list) else [it]
        cur = (cur + _ALIGN - 1) // _ALIGN * _ALIGN
        self.n_param = cur
        params = list(model.parameters())
        missing = [n for n, p in model.named_parameters() if id(p) not in covered]
        assert not missing, f"parameters outside the launch plan: {missing[:5]}"
        bufs = [b for b in model.buffers() if b.dtype.is_floating_point]
        for b in bufs:
            place(b)
        cur = (cur + _ALIGN - 1) // _ALIGN * _ALIGN
        self.n_total = cur
        self.state = torch.zeros(self.n_total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.n_param, dtype=torch.float32, device=dev)
        self.mom = torch.zeros(self.n_param, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for t in params + bufs:
                o, n = self.offset[id(t)]
                v = self.state[o:o + n].view(t.shape)
                v.copy_(t.detach().float())
                t.data = v                                              # the module tensor is now a view of the flat state
            for p in params:
                o, n = self.offset[id(p)]
                p.grad = self.grad[o:o + n].view(p.shape)
        self.ema = self.state.clone() if ema else None
        self.params = params

    def gview(self, p):
        o, n = self.offset[id(p)]
        return self.grad[o:o + n]

    def ema_state_dict(self, model):
        """state_dict of the EMA model ([yolox] ModelEMA.ema.state_dict()): float entries from the flat EMA copy, the rest
        (num_batches_tracked) as in the live model."""
        base = self.state.data_ptr()
        index = {base + 4 * o: (o, n) for (o, n) in self.offset.values()}
        out = {}
        for k, t in model.state_dict().items():
            hit = index.get(t.data_ptr()) if t.dtype == torch.float32 else None
            out[k] = self.ema[hit[0]:hit[0] + hit[1]].view(t.shape).clone() if hit else t.clone()
        return out


class FlatSink:
    """Gradient sink of the backward walk that writes into ``FlatState.grad`` and launches the all-reduce of a bucket as
    soon as all of its gradients have been enqueued."""

    def __init__(self, fs: FlatState, uses, bucket_bytes=25 << 20, overlap=True, world=None):
        self.fs, self.uses, self.overlap = fs, uses, overlap
        self.world = (dist.get_world_size() if dist.is_initialized() else 1) if world is None else world
        self.touched = set()
        # buckets over the decayed class in walk order; the small no-decay class is one last bucket
        self.buckets = []                   # [start, end, pending parameter count]
        self.bucket_of = {}
        start, size, members = fs.decay_begin, 0, []
        for p in fs.order_b:
            o, n = fs.offset[id(p)]
            if members and size + 4 * n > bucket_bytes:
                self._close(start, o, members)
                start, size, members = o, 0, []
            members.append(p)
            size += 4 * n
        self._close(start, fs.n_param, members)
        small = [p for p in fs.params if fs.offset[id(p)][0] < fs.decay_begin]
        self._close(0, fs.decay_begin, small)
        self.work = []
        self.launched = []
        self.on_bucket = None

    def _close(self, a, b, members):
        if not members:
            return
        idx = len(self.buckets)
        self.buckets.append([a, b, len(members)])
        for p in members:
            self.bucket_of[id(p)] = idx

    # ---- buffers for the kernels
    def _first(self, key):
        acc = key in self.touched
        self.touched.add(key)
        return acc

    def conv_weight(self, mods, cin, kh, kw, stem=False):
        o, _ = self.fs.offset[id(mods[0].conv.weight)]
        n = sum(m.conv.weight.numel() for m in mods)
        t = self.fs.grad[o:o + n]
        shape = tuple(mods[0].conv.weight.shape) if stem else (sum(m.conv.out_channels for m in mods), cin, kh, kw)
        return t.view(shape), self._first(("w", id(mods[0])))

    def bn(self, mods):
        c = sum(m.conv.out_channels for m in mods)
        og, _ = self.fs.offset[id(mods[0].bn.weight)]
        ob, _ = self.fs.offset[id(mods[0].bn.bias)]
        return self.fs.grad[og:og + c], self.fs.grad[ob:ob + c], self._first(("bn", id(mods[0])))

    def head(self, head, k):
        ws = [self.fs.gview(p).view(p.shape[0], p.shape[1]) for p in (head.reg_preds[k].weight, head.obj_preds[k].weight,
                                                                      head.cls_preds[k].weight)]
        bs = [self.fs.gview(p) for p in (head.reg_preds[k].bias, head.obj_preds[k].bias, head.cls_preds[k].bias)]
        return ws, bs, False

    # ---- completion tracking / communication
    def done(self, params):
        for p in params:
            key = id(p)
            left = self.pending.get(key)
            if left is None:
                continue
            left -= 1
            self.pending[key] = left
            if left == 0:
                b = self.buckets[self.bucket_of[key]]
                b[2] -= 1
                if b[2] == 0 and self.overlap:
                    self._launch(b)

    def begin(self, model):
        """per step: how many launches contribute to each parameter (a module recorded twice, e.g. DFP jian, finishes on its
        last launch)"""
        self.touched.clear()
        self.pending = {}
        for g in conv_groups_forward_order(model):
            n = self.uses.get(id(g[0]), 1)
            for m in g:
                for p in (m.conv.weight, m.bn.weight, m.bn.bias):
                    self.pending[id(p)] = n
        head = model.head
        for k in range(len(head.stems)):
            for p in (head.reg_preds[k].weight, head.obj_preds[k].weight, head.cls_preds[k].weight, head.reg_preds[k].bias,
                      head.obj_preds[k].bias, head.cls_preds[k].bias):
                self.pending[id(p)] = 1
        for i, b in enumerate(self.buckets):
            b[2] = sum(1 for k, v in self.bucket_of.items() if v == i)
        self.work, self.launched = [], []

    def _launch(self, b):
        self.launched.append((b[0], b[1]))
        if self.world > 1:
            if self.on_bucket is not None:
                self.on_bucket(b[0], b[1])          # graph capture: the trainer cuts the graph here and owns the collective
            else:
                self.work.append(dist.all_reduce(self.fs.grad[b[0]:b[1]], op=dist.ReduceOp.SUM, async_op=True))

    def finish(self):
        for b in self.buckets:
            if (b[0], b[1]) not in self.launched:
                self._launch(b)
        for w in self.work:
            w.wait()                        # the compute stream waits for the communication stream; no host sync


class Trainer:
    """B200-native training loop body for YOLOX(DFPPAFPN, TALHead): ``step(x, targets, lr)``."""

    def __init__(self, model, lr=0.01, momentum=0.9, weight_decay=5e-4, ema_decay=0.9998, use_ema=True,
                 bucket_bytes=25 << 20, overlap=True):
        assert model.training and model.head.use_l1
        self.model = model
        self.fs = FlatState(model, ema=use_ema)
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.ema_decay, self.updates = ema_decay, 0
        self.bucket_bytes, self.overlap = bucket_bytes, overlap
        self.sink = None
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        engine.WEIGHT_EPOCH += 1
        self._build_repack()
        self._repack()

    # ---- conv operands: every (parameter, layout) pair of the model re-packed by ONE launch after each update
    def _build_repack(self):
        dev = self.fs.state.device
        self.pack = ops.PackBatch(dev)
        self._packed_groups = []
        stem = self.model.backbone.backbone.stem.conv
        dt = ops.pack_conv_weight(stem.conv.weight[:1]).dtype     # bf16 (fp32 only under the CPU emulation with fp32 "storage")
        for g in conv_groups_forward_order(self.model):
            ws = [m.conv.weight for m in g]
            _, cin, kh, kw = ws[0].shape
            ot = sum(w.shape[0] for w in ws)
            if g[0] is stem:
                out = torch.empty((ot, kh, 64), dtype=dt, device=dev)
                self.pack.add(ws[0], out, 2)
                self._packed_groups.append((g, out, None))
                continue
            fwd = torch.empty((ot, kh * kw, cin), dtype=dt, device=dev)
            dg = torch.empty((cin, kh * kw, ot), dtype=dt, device=dev)
            o0 = 0
            for w in ws:
                self.pack.add(w, fwd[o0:o0 + w.shape[0]], 0)
                self.pack.add(w, dg, 1, out_pitch=ot, co_offset=o0)
                o0 += w.shape[0]
            self._packed_groups.append((g, fwd, dg))

    def _repack(self):
        """run the batched pack and hand the buffers to the engine's operand caches (keys of the current WEIGHT_EPOCH)"""
        self.pack.run()
        ep = engine.WEIGHT_EPOCH
        for g, fwd, dg in self._packed_groups:
            ws = [m.conv.weight for m in g]
            if dg is None:                                  # stem
                g[0]._pk, g[0]._pk_key = fwd, (ws[0]._version, ws[0].data_ptr(), ws[0].device, ep)
                continue
            if len(g) == 1:
                g[0]._pk, g[0]._pk_key = fwd, (ws[0]._version, ws[0].data_ptr(), ws[0].device, ep)
            else:
                g[0]._pk2 = fwd
                g[0]._pk2_key = (ws[0]._version, ws[1]._version, ws[0].data_ptr(), ws[1].data_ptr(), ws[0].device, ep)
            g[0]._pkd, g[0]._pkd_key = dg, tuple((w._version, w.data_ptr()) for w in ws) + (ep,)

    def forward_backward(self, x, targets, loss_scale=1.0):
        T, loss = backward._record(self.model, x, targets)
        if self.sink is None:
            self.sink = FlatSink(self.fs, dict(T.uses), self.bucket_bytes, self.overlap, self.world)
        self.sink.uses = dict(T.uses)
        self.sink.begin(self.model)
        with torch.no_grad():
            backward._walk(T, self.model.head, loss_scale, self.sink)
        return loss

    def _hyper_values(self, lr, loss_scale):
        d = self.ema_decay * (1 - math.exp(-self.updates / 2000)) if self.fs.ema is not None else 0.0
        return [self.lr if lr is None else lr, self.momentum, self.weight_decay, 1.0 / (self.world * loss_scale), d, 1.0 - d]

    def optimizer_step(self, lr=None, loss_scale=1.0, found_inf=None, hyper=None):
        if hyper is None:
            self.updates += 1
        h = self._hyper_values(lr, loss_scale)
        ops.sgd_nesterov_ema_step(self.fs.state, self.fs.grad, self.fs.mom, self.fs.ema, self.fs.n_param, self.fs.decay_begin,
                                  h[0], h[1], h[2], inv_scale=h[3], nesterov=True, ema_decay=h[4], found_inf=found_inf,
                                  hyper=hyper)
        engine.WEIGHT_EPOCH += 1            # the conv operands follow the new masters: one batched re-pack launch
        self._repack()

    def step(self, x, targets, lr=None, loss_scale=1.0):
        loss = self.forward_backward(x, targets, loss_scale)
        self.optimizer_step(lr, loss_scale)
        return backward._loss_dict(loss)

    # ---- the whole step as CUDA graph(s) (the eager step is bound by the host's launch rate: ~800-1300 launches)
    def capture(self, x, targets, loss_scale=1.0):
        """Capture forward + backward + weight re-pack + optimiser step on static input buffers (``x`` / ``targets`` become the
        graph's inputs: copy new data into them, then ``replay(lr)``).  One process: ONE graph.  Several ranks: the walk is cut
        into one graph segment per gradient bucket; between two segments the host enqueues that bucket's NCCL all-reduce on the
        communication stream, where it overlaps the following segments (the collectives themselves stay outside the
        captured graphs); a last segment holds the optimiser step."""
        dev = self.fs.state.device
        self._hyper = torch.zeros(8, dtype=torch.float32, device=dev)
        self._hyper_host = torch.zeros(8, dtype=torch.float32).pin_memory()
        self._loss_scale = loss_scale
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up on a side stream (allocator, lazy module attributes)
            self._set_hyper(None)
            self.forward_backward(x, targets, loss_scale)
            self.optimizer_step(hyper=self._hyper)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        pool = torch.cuda.graph_pool_handle()
        self._plan = []                                    # [(graph, bucket range or None)]
        cur = [None]

        def begin():
            g = torch.cuda.CUDAGraph()
            g.capture_begin(pool=pool)
            cur[0] = g

        def cut(a, b):
            cur[0].capture_end()
            self._plan.append((cur[0], (a, b)))
            begin()

        self.sink.on_bucket = cut if self.world > 1 else None
        with torch.cuda.stream(side):
            begin()
            loss = self.forward_backward(x, targets, loss_scale)
            if self.world > 1:                             # the optimiser step waits for the collectives: its own segment
                cur[0].capture_end()
                self._plan.append((cur[0], None))
                begin()
            self.optimizer_step(hyper=self._hyper)
            cur[0].capture_end()
            self._plan.append((cur[0], None))
        self.sink.on_bucket = None
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph_loss = loss
        return len(self._plan)

    def _set_hyper(self, lr):
        self.updates += 1
        self._hyper_host[:6] = torch.tensor(self._hyper_values(lr, self._loss_scale))
        self._hyper.copy_(self._hyper_host, non_blocking=True)

    def replay(self, lr=None):
        self._set_hyper(lr)
        works = []
        for i, (g, bucket) in enumerate(self._plan):
            if i == len(self._plan) - 1:
                for w in works:
                    w.wait()
            g.replay()
            if bucket is not None:
                works.append(dist.all_reduce(self.fs.grad[bucket[0]:bucket[1]], op=dist.ReduceOp.SUM, async_op=True))
        return backward._loss_dict(self._graph_loss)

    def ema_state_dict(self):
        return self.fs.ema_state_dict(self.model)
