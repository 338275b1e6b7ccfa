"""Layer executor: walks the module tree with NHWC bf16 views and launches the CUDA ops.

Train mode (``model.training``): every BaseConv = tcgen05 conv writing the raw bf16 result + per-tile
statistic partials  ->  bn_finalize (batch statistics, running-stat update)  ->  bn_act_apply
(normalise + SiLU + optional residual, written straight into its consumer's concat slice).
The two frames of a pair are batched through the shared-weight backbone as 2B images with
*grouped* statistics (group 0 = current frames, group 1 = support frames), which reproduces the
reference's two sequential passes (/root/reference/exps/model/dfp_pafpn.py:120,145) exactly,
including the order of the two running-statistic updates.

Eval mode: BatchNorm is folded into a per-channel scale/shift applied in the conv epilogue
together with SiLU and the residual (what yolox ``fuse_model`` + ``fuseforward`` achieve).
"""
import os

import torch

from .. import ops
from ..ops import View


# normalise pass inside the conv launch (1 launch / BaseConv).  For every layer it was measured SLOWER on B200 (11.7 vs
# 8.2 ms/step: one CTA per SM cannot keep enough bytes in flight on the big tensors), so SY_FUSE_APPLY=1 is a debug
# switch; SY_FUSE_APPLY_MAX_MB fuses only the layers whose raw output is at most that many MB (L2 resident, launch-bound)
FUSE_APPLY = os.environ.get("SY_FUSE_APPLY", "0") != "0"
FUSE_APPLY_MAX_BYTES = float(os.environ.get("SY_FUSE_APPLY_MAX_MB", "0")) * 1e6
WEIGHT_EPOCH = 0  # bumped by whoever updates parameters through raw pointers (train.Trainer's fused optimiser kernel does
                  # not touch torch's version counters): part of every packed-operand cache key
TRACE = None      # debugging: set to a dict to capture every BaseConv's stored output by module name


def name_modules(model):
    for n, m in model.named_modules():
        m._sy_name = n


def _trace(m, y):
    if TRACE is not None:
        TRACE[getattr(m, "_sy_name", str(id(m)))] = y.torch().permute(0, 3, 1, 2).float().cpu()


class Ctx:
    """Per-forward execution context."""

    def __init__(self, train, n, split, device):
        self.train = train
        self.n = n              # images in the batched tensor
        self.split = split      # first image of statistics group 1 (== n: single group)
        self.groups = 2 if split < n else 1
        self.device = device
        self.impl = os.environ.get("SY_CONV_IMPL", "tc")


def _packed(m):
    w = m.conv.weight
    key = (w._version, w.data_ptr(), w.device, WEIGHT_EPOCH)
    if getattr(m, "_pk_key", None) != key:
        m._pk = ops.pack_conv_weight(w)
        m._pk_key = key
    return m._pk


def _packed_dw(m):
    """depthwise BaseConv: bf16 [k*k][C]"""
    w = m.conv.weight
    key = (w._version, w.data_ptr(), w.device, WEIGHT_EPOCH)
    if getattr(m, "_pk_key", None) != key:
        m._pk = ops.pack_dw_weight(w)
        m._pk_key = key
    return m._pk


def _packed_dgrad(mods):
    """Data-gradient operand (flipped taps, transposed channels) of one BaseConv or of a conv1 | conv2 pair."""
    ws = [m.conv.weight for m in mods]
    key = tuple((w._version, w.data_ptr()) for w in ws) + (WEIGHT_EPOCH,)
    m0 = mods[0]
    if getattr(m0, "_pkd_key", None) != key:
        m0._pkd = ops.pack_conv_weight_dgrad(*ws)
        m0._pkd_key = key
    return m0._pkd


def _folded(m):
    """Eval: scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale (fp32)."""
    bn = m.bn
    key = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
           getattr(m, "_stats_epoch", 0), bn.weight.data_ptr(), bn.eps, WEIGHT_EPOCH)
    if getattr(m, "_fold_key", None) != key:
        with torch.no_grad():
            scale = bn.weight.float() * torch.rsqrt(bn.running_var.float() + bn.eps)
            shift = bn.bias.float() - bn.running_mean.float() * scale
        m._fold = (scale.contiguous(), shift.contiguous())
        m._fold_key = key
    return m._fold


SYNC_SLOTS = 1024
_SYNC_POOLS = {}      # (device, stream) -> [int32 tensor of 4 * SYNC_SLOTS counters, next slot, scope depth]


def _sync_pool(device):
    key = (str(device), torch.cuda.current_stream().cuda_stream if torch.device(device).type == "cuda" else 0)
    st = _SYNC_POOLS.get(key)
    if st is None:
        st = [torch.zeros(4 * SYNC_SLOTS, dtype=torch.int32, device=device), 0, 0]
        _SYNC_POOLS[key] = st
    return st


class forward_scope:
    """Outermost scope of one forward (YOLOX / DFPPAFPN / TALHead / a stand-alone block / the recording forward): rewinds the
    grid-barrier counter pool of the current stream, so that the first train-mode conv of the forward zeroes it (ONE memset
    per forward, captured into the CUDA graph with it).  Every launch then takes its own counter slot: a launch that was
    aborted, or a module used by two forwards, can no longer leave a stale count for the next launch (the kernels also
    leave their slot at zero when they complete).  Launches on ONE stream only: two concurrent train-mode convs would each
    need every SM for their grid barrier (include/streamyolo_sm100.h)."""

    def __init__(self, device):
        self.device = device
        self.st = _sync_pool(device)

    def __enter__(self):
        if self.st[2] == 0:
            self.st[1] = 0
            _arm_raw_window(self.device)
        self.st[2] += 1
        return self

    def __exit__(self, *a):
        self.st[2] -= 1
        return False


def _sync(m, device):
    """Four zeroed counters for one train-mode conv launch (two grid barriers + exit ticket), from the stream's pool."""
    st = _sync_pool(device)
    if st[1] == 0:
        st[0].zero_()
    i = st[1]
    st[1] = (i + 1) % SYNC_SLOTS
    return st[0][4 * i:4 * i + 4]


def _bn_seg(m, c_begin=0):
    bn = m.bn
    return (bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked, c_begin)


# raw conv outputs of the plain (non-recording) train-mode forward: one arena per (device, stream), marked persisting in L2
RAW_ARENA_MB = float(os.environ.get("SY_RAW_ARENA_MB", "40"))
_RAW_ARENAS = {}


def _raw_view(ctx, n, h, w, c):
    """The raw bf16 conv output of a train-mode BaseConv: dead as soon as its normalise pass has run, so every layer whose
    tensor fits aliases ONE arena that the launching stream treats as a persisting-L2 window (sy_l2_persist_window)."""
    nbytes = n * h * w * c * 2
    if RAW_ARENA_MB <= 0 or torch.device(ctx.device).type != "cuda" or nbytes > RAW_ARENA_MB * 1e6:
        return View.empty(n, h, w, c, ctx.device)
    key = str(ctx.device)
    st = _RAW_ARENAS.get(key)
    if st is None:
        if torch.cuda.is_current_stream_capturing():       # the device limit / stream attribute cannot be set during capture
            return View.empty(n, h, w, c, ctx.device)
        arena = torch.empty(int(RAW_ARENA_MB * 1e6) // 256 * 256, dtype=torch.uint8, device=ctx.device)
        st = _RAW_ARENAS[key] = [arena, ops.l2_persist_window(arena)]
    return View(st[0][:nbytes].view(torch.bfloat16).view(n, h, w, c))


def _arm_raw_window(device):
    """(re)apply the persisting window on the CURRENT stream (a forward may run on another stream than the one the arena was
    created on).  Not during stream capture: capture on ``graph_capture_stream()``, which carries the window already."""
    st = _RAW_ARENAS.get(str(device))
    if st is not None and st[1] > 0 and not torch.cuda.is_current_stream_capturing():
        ops.l2_persist_window(st[0])


_CAPTURE_STREAMS = {}


def graph_capture_stream(device):
    """A side stream to capture CUDA graphs of the forward on (``torch.cuda.graph(g, stream=...)``): the persisting-L2 window of
    the raw arena is set on it BEFORE the capture starts, so the captured kernel nodes inherit it."""
    key = str(device)
    if key not in _CAPTURE_STREAMS:
        _CAPTURE_STREAMS[key] = torch.cuda.Stream(device=device)
    st = _CAPTURE_STREAMS[key]
    with torch.cuda.stream(st):
        _arm_raw_window(device)
    return st


def _dbg_skip_apply(nbytes):
    """timing experiments only (tools/ab_step.py): SY_DBG_SKIP_APPLY="lo:hi" (MB) drops the normalise pass of the layers whose
    raw output size lies in [lo, hi) -- the results are garbage, the step time shows what those launches really cost"""
    e = os.environ.get("SY_DBG_SKIP_APPLY")
    if not e:
        return False
    lo, hi = (float(v) for v in e.split(":"))
    return lo * 1e6 <= nbytes < hi * 1e6


def conv_bn_act(ctx: Ctx, mods, x: View, wpk, k, s, y: View, res: View = None, act=1, y_goff1=0, res_goff1=0, impl=None):
    """Train mode: conv -> batch statistics -> BatchNorm (running-stat update) -> act (+res) into ``y``.
    Tensor-core path = 2 launches: the conv writes the raw bf16 result, accumulates the statistics and
    (grid barrier + parallel reduce in its tail) publishes scale/shift; then the normalise pass.  ``mods``: one BaseConv, or two whose
    outputs are concatenated along channels (CSPLayer conv1 | conv2)."""
    kh, kw = (k, k) if isinstance(k, int) else k
    ho = (x.h + 2 * ((kh - 1) // 2) - kh) // s + 1
    wo = (x.w + 2 * ((kw - 1) // 2) - kw) // s + 1
    cout = sum(m.conv.out_channels for m in mods)
    raw = _raw_view(ctx, x.n, ho, wo, cout)
    bn0 = mods[0].bn
    mom = float(0.1 if bn0.momentum is None else bn0.momentum)
    n = x.n
    split = ctx.split if ctx.groups == 2 else 0
    for m in mods:
        m._stats_epoch = getattr(m, "_stats_epoch", 0) + 1
    impl = impl or ctx.impl
    if impl == "tc":
        partials = torch.empty((ops.conv_stat_rows(), 4 * cout), dtype=torch.float32, device=ctx.device)
        segs, c0 = [], 0
        for m in mods:
            segs.append(_bn_seg(m, c0))
            c0 += m.conv.out_channels
        ss = torch.empty((2, 2, cout), dtype=torch.float32, device=ctx.device)
        if FUSE_APPLY or n * ho * wo * cout * 2 <= FUSE_APPLY_MAX_BYTES:
            ops.conv2d(x, wpk, raw, k, s, ops.SY_CONV_RAW, impl="tc", partials=partials, split_n=split, bn=segs,
                       momentum=mom, eps=float(bn0.eps), scale_shift=ss, sync=_sync(mods[0], ctx.device), act=act,
                       apply_y=y, apply_res=res, y_goff1=y_goff1, res_goff1=res_goff1)
        else:
            ops.conv2d(x, wpk, raw, k, s, ops.SY_CONV_RAW, impl="tc", partials=partials, split_n=split, bn=segs,
                       momentum=mom, eps=float(bn0.eps), scale_shift=ss, sync=_sync(mods[0], ctx.device))
            if not _dbg_skip_apply(n * ho * wo * cout * 2):
                ops.bn_act_apply(raw, ss[0].data_ptr(), ss[1].data_ptr(), split if split else n, act, res, y, y_goff1,
                                 res_goff1)
        return
    # CUDA-core path (cross-check of the tensor-core kernel; depthwise convs): conv, separate statistics pass, separate
    # finalize per module, apply
    ops.conv2d(x, wpk, raw, k, s, ops.SY_CONV_RAW, impl=impl)
    sc = torch.empty((2, 2, cout), dtype=torch.float32, device=ctx.device)
    sp = split if split else n
    c0 = 0
    for m in mods:
        c = m.conv.out_channels
        P = ops.stats_num_partials(n, ho * wo)
        partials = torch.empty((P, 2, c), dtype=torch.float32, device=ctx.device)
        ops.channel_stats(raw.ch(c0, c), partials)
        tmp = torch.empty((2, 2, c), dtype=torch.float32, device=ctx.device)
        bn = m.bn
        ops.bn_finalize(partials, (P // n) * sp if split else 0, 2 if split else 1, sp * ho * wo,
                        bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                        mom, float(bn.eps), tmp[0], tmp[1])
        sc[:, :, c0:c0 + c] = tmp
        c0 += c
    if y_goff1 == 0 and res_goff1 == 0:
        ops.bn_act_apply(raw, sc[0], sc[1], sp, act, res, y)
    else:   # group-1 images go to a shifted destination (DFP fusion): one call per group
        nb = sp
        ops.bn_act_apply(raw.imgs(0, nb), sc[0, 0], sc[1, 0], nb, act,
                         res.imgs(0, nb) if res is not None else None, y.imgs(0, nb))
        y1 = View(y.buf, y.c0, y.c, y.n0, nb).shifted(y_goff1 + nb * y.img_elems())
        r1 = View(res.buf, res.c0, res.c, res.n0, nb).shifted(res_goff1 + nb * res.img_elems()) if res is not None else None
        ops.bn_act_apply(raw.imgs(nb, nb), sc[0, 1], sc[1, 1], nb, act, r1, y1)


def base_conv(ctx: Ctx, m, x: View, y: View = None, res: View = None) -> View:
    """[yolox] BaseConv: act(bn(conv(x))) (+ res).  ``y`` may be a slice of a concat buffer.  Also takes a [yolox] DWConv
    (depthwise BaseConv then pointwise BaseConv) wherever the reference's ``Conv = DWConv if depthwise else BaseConv`` puts one."""
    if hasattr(m, "dconv"):
        return base_conv(ctx, m.pconv, base_conv(ctx, m.dconv, x), y, res)
    k, s = m.ksize, m.stride
    ho, wo = ops.conv_out_hw(x.h, x.w, k, s)
    cout = m.conv.out_channels
    if y is None:
        y = View.empty(x.n, ho, wo, cout, ctx.device)
    dw = m.conv.groups > 1
    wpk = _packed_dw(m) if dw else _packed(m)
    impl = "dw" if dw else ctx.impl
    act = 1 if m.act_name == "silu" else 0
    if not ctx.train:
        scale, shift = _folded(m)
        ops.conv2d(x, wpk, y, k, s, ops.SY_CONV_FUSED, impl=impl, scale=scale, shift=shift, act=act, res=res)
    else:
        conv_bn_act(ctx, (m,), x, wpk, k, s, y, res, act, impl=impl)
    _trace(m, y)
    return y


def _packed_pair(m1, m2):
    """conv1 | conv2 of a CSPLayer as one [2*hidden][1][Cin] GEMM operand."""
    w1, w2 = m1.conv.weight, m2.conv.weight
    key = (w1._version, w2._version, w1.data_ptr(), w2.data_ptr(), w1.device, WEIGHT_EPOCH)
    if getattr(m1, "_pk2_key", None) != key:
        m1._pk2 = ops.pack_conv_weight(w1, w2)
        m1._pk2_key = key
    return m1._pk2


def _folded_pair(m1, m2):
    a, b = _folded(m1), _folded(m2)
    key = (id(a[0]), id(b[0]))
    if getattr(m1, "_fold2_key", None) != key:
        m1._fold2 = (torch.cat([a[0], b[0]]).contiguous(), torch.cat([a[1], b[1]]).contiguous())
        m1._fold2_key = key
    return m1._fold2


def conv_pair(ctx: Ctx, m1, m2, x: View) -> View:
    """Two BaseConvs with the same geometry reading the same input as ONE launch: [.., c1 + c2] output, one BatchNorm
    parameter segment per module (CSPLayer conv1 | conv2; the first cls / reg tower convs of a head level)."""
    if hasattr(m1, "dconv") or hasattr(m2, "dconv"):          # depthwise variants: two ordinary launches into one buffer
        c1, c2 = m1.pconv.conv.out_channels, m2.pconv.conv.out_channels
        u = View.empty(x.n, x.h, x.w, c1 + c2, ctx.device)
        base_conv(ctx, m1, x, u.ch(0, c1))
        base_conv(ctx, m2, x, u.ch(c1, c2))
        return u
    c1, c2 = m1.conv.out_channels, m2.conv.out_channels
    k, s = m1.ksize, m1.stride
    assert (m2.ksize, m2.stride, m2.conv.in_channels) == (k, s, m1.conv.in_channels)
    ho, wo = ops.conv_out_hw(x.h, x.w, k, s)
    u = View.empty(x.n, ho, wo, c1 + c2, ctx.device)
    wpk = _packed_pair(m1, m2)
    if not ctx.train:
        scale, shift = _folded_pair(m1, m2)
        ops.conv2d(x, wpk, u, k, s, ops.SY_CONV_FUSED, impl=ctx.impl, scale=scale, shift=shift, act=1)
    else:
        conv_bn_act(ctx, (m1, m2), x, wpk, k, s, u)
    _trace(m1, u.ch(0, c1))
    _trace(m2, u.ch(c1, c2))
    return u


def csp_layer(ctx: Ctx, m, x: View, out: View = None) -> View:
    """[yolox] CSPLayer: conv3(cat(m(conv1 x), conv2 x)).  conv1 and conv2 read the same input, so they
    run as ONE GEMM with 2*hidden output channels written straight into the concat buffer; the
    bottleneck chain then updates the first half in place.  No concat copy ever happens."""
    hid = m.conv1.conv.out_channels
    u = View.empty(x.n, x.h, x.w, 2 * hid, ctx.device)
    a = u.ch(0, hid)
    wpk = _packed_pair(m.conv1, m.conv2)
    if not ctx.train:
        scale, shift = _folded_pair(m.conv1, m.conv2)
        ops.conv2d(x, wpk, u, 1, 1, ops.SY_CONV_FUSED, impl=ctx.impl, scale=scale, shift=shift, act=1)
    else:
        conv_bn_act(ctx, (m.conv1, m.conv2), x, wpk, 1, 1, u)
    _trace(m.conv1, a)
    _trace(m.conv2, u.ch(hid, hid))
    for blk in m.m:
        t = base_conv(ctx, blk.conv1, a)
        base_conv(ctx, blk.conv2, t, a, res=a if blk.use_add else None)
    return base_conv(ctx, m.conv3, u, out)


def _packed_stem(bc):
    w = bc.conv.weight
    key = (w._version, w.data_ptr(), w.device, WEIGHT_EPOCH)
    if getattr(bc, "_pk_key", None) != key:
        bc._pk = ops.pack_stem_weight(w)
        bc._pk_key = key
    return bc._pk


def focus_stem(ctx: Ctx, m, x, frames) -> View:
    """[yolox] Focus + BaseConv straight from the NCHW float frame-pair batch: space-to-depth + W-gather into
    a 64-channel NHWC tensor, then the tensor-core kernel runs the 3x3 stem as a 3x1 conv (3 K blocks)."""
    b, ch, h, w = x.shape
    bc = m.conv
    cout = bc.conv.out_channels
    n = frames * b
    xin = View.empty(n, h // 2, w // 2, 64, ctx.device)
    ops.focus_pack(x, frames, xin)
    wpk = _packed_stem(bc)
    y = View.empty(n, h // 2, w // 2, cout, ctx.device)
    if not ctx.train:
        scale, shift = _folded(bc)
        ops.conv2d(xin, wpk, y, ops.STEM_K, 1, ops.SY_CONV_FUSED, impl=ctx.impl, scale=scale, shift=shift, act=1)
    else:
        conv_bn_act(ctx, (bc,), xin, wpk, ops.STEM_K, 1, y)
    _trace(bc, y)
    return y


def spp_bottleneck(ctx: Ctx, m, x: View) -> View:
    hid = m.conv1.conv.out_channels
    s = View.empty(x.n, x.h, x.w, 4 * hid, ctx.device)
    base_conv(ctx, m.conv1, x, s.ch(0, hid))
    ops.spp_maxpool(s.ch(0, hid), s.ch(hid, hid), s.ch(2 * hid, hid), s.ch(3 * hid, hid))
    return base_conv(ctx, m.conv2, s)


def pafpn_frames(ctx: Ctx, net, x, frames):
    """CSPDarknet + PAFPN for ``frames`` x B images (/root/reference/exps/model/darknet.py:167-179,
    dfp_pafpn.py:120-140).  Returns the un-fused (pan_out2, pan_out1, pan_out0) views."""
    bb = net.backbone
    dev = ctx.device
    c3 = net.C3_p3.conv3.conv.out_channels
    c4 = net.C3_p4.conv3.conv.out_channels
    t = focus_stem(ctx, bb.stem, x, frames)
    t = base_conv(ctx, bb.dark2[0], t)
    t = csp_layer(ctx, bb.dark2[1], t)
    t = base_conv(ctx, bb.dark3[0], t)
    n, h8, w8 = t.n, t.h, t.w
    f1 = View.empty(n, h8, w8, 2 * c3, dev)              # cat(up(fpn_out1), dark3)
    x2 = csp_layer(ctx, bb.dark3[1], t, f1.ch(c3, c3))
    t = base_conv(ctx, bb.dark4[0], x2)
    h16, w16 = t.h, t.w
    f0 = View.empty(n, h16, w16, 2 * c4, dev)            # cat(up(fpn_out0), dark4)
    x1 = csp_layer(ctx, bb.dark4[1], t, f0.ch(c4, c4))
    t = base_conv(ctx, bb.dark5[0], x1)
    h32, w32 = t.h, t.w
    t = spp_bottleneck(ctx, bb.dark5[1], t)
    x0 = csp_layer(ctx, bb.dark5[2], t)
    z0 = View.empty(n, h32, w32, 2 * c4, dev)            # cat(bu_conv1, fpn_out0)
    fpn0 = base_conv(ctx, net.lateral_conv0, x0, z0.ch(c4, c4))
    ops.upsample_nearest(fpn0, f0.ch(0, c4))
    fo0 = csp_layer(ctx, net.C3_p4, f0)
    z1 = View.empty(n, h16, w16, 2 * c3, dev)            # cat(bu_conv2, fpn_out1)
    fpn1 = base_conv(ctx, net.reduce_conv1, fo0, z1.ch(c3, c3))
    ops.upsample_nearest(fpn1, f1.ch(0, c3))
    pan2 = csp_layer(ctx, net.C3_p3, f1)
    base_conv(ctx, net.bu_conv2, pan2, z1.ch(0, c3))
    pan1 = csp_layer(ctx, net.C3_n3, z1)
    base_conv(ctx, net.bu_conv1, pan1, z0.ch(0, c4))
    pan0 = csp_layer(ctx, net.C3_n4, z0)
    return pan2, pan1, pan0


def dfp_fuse(ctx: Ctx, net, cur, sup):
    """Dual-Flow Perception fusion (/root/reference/exps/model/dfp_pafpn.py:168-170, 211-221):
    out = cat(jian(cur), jian(sup)) + cur, a single bf16 rounding after the residual add.
    ``cur`` / ``sup`` are per-level views with the same image count."""
    outs = []
    for m, c, s in zip((net.jian2, net.jian1, net.jian0), cur, sup):
        nb = c.n
        if hasattr(m, "dconv"):                               # depthwise=True: jian is a DWConv (dfp_pafpn.py:83-105)
            half = m.pconv.conv.out_channels
            out = View.empty(nb, c.h, c.w, 2 * half, ctx.device)
            sub = Ctx(ctx.train, nb, nb, ctx.device)          # two calls = two BatchNorm batches, like the reference
            base_conv(sub, m, c, out.ch(0, half), res=c.ch(0, half))
            base_conv(sub, m, s, out.ch(half, half), res=c.ch(half, half))
            outs.append(out)
            continue
        half = m.conv.out_channels
        out = View.empty(nb, c.h, c.w, 2 * half, ctx.device)
        wpk = _packed(m)
        if not ctx.train:
            scale, shift = _folded(m)
            ops.conv2d(c, wpk, out.ch(0, half), 1, 1, ops.SY_CONV_FUSED, impl=ctx.impl, scale=scale, shift=shift,
                       act=1, res=c.ch(0, half))
            ops.conv2d(s, wpk, out.ch(half, half), 1, 1, ops.SY_CONV_FUSED, impl=ctx.impl, scale=scale,
                       shift=shift, act=1, res=c.ch(half, half))
        else:
            # the reference runs jian(cur) then jian(sup): two BN batches, two running-stat updates.
            # Batched here (grouped statistics) when cur/sup are the two halves of one buffer.
            same = (c.buf is s.buf) and s.n0 == c.n0 + nb and c.c0 == s.c0 and c.c == s.c
            if same:
                both = View(c.buf, c.c0, c.c, c.n0, 2 * nb)
                sub = Ctx(True, 2 * nb, nb, ctx.device)
                # group 1 (support frames, images nb..2nb-1) lands in channels [half, 2*half) of image n - nb
                yv = View(out.buf, 0, half, 0, 2 * nb)
                rv = View(c.buf, c.c0, half, c.n0, 2 * nb)
                conv_bn_act(sub, (m,), both, wpk, 1, 1, yv, rv, 1,
                            y_goff1=half - nb * yv.img_elems(), res_goff1=half - nb * rv.img_elems())
            else:
                sub = Ctx(True, nb, nb, ctx.device)
                for src, dst, r in ((c, out.ch(0, half), c.ch(0, half)), (s, out.ch(half, half), c.ch(half, half))):
                    conv_bn_act(sub, (m,), src, wpk, 1, 1, dst, r, 1)
        outs.append(out)
    return tuple(outs)


def as_view(t) -> View:
    """Accept a View or an NCHW-shaped torch tensor (zero-copy when it is channels-last bf16)."""
    if isinstance(t, View):
        return t
    p = t.permute(0, 2, 3, 1)
    if t.dtype == torch.bfloat16 and p.is_contiguous():
        return View(p)
    return View(p.contiguous().to(torch.bfloat16))


def as_nchw(v: View):
    """NCHW-shaped (channels-last memory) tensor over a view, as the reference API returns."""
    return v.torch().permute(0, 3, 1, 2)
