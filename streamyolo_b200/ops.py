"""ctypes binding of libstreamyolo_sm100.so (include/streamyolo_sm100.h) + NHWC view helper.

The library is the product; there is NO fallback: if it is missing or the device is not an
sm_100 part, every op raises RuntimeError.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libstreamyolo_sm100.so")

SY_CONV_RAW, SY_CONV_FUSED = 0, 1


class SyTensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("pitch", C.c_int64)]


class SyBnSegment(C.Structure):
    _fields_ = [("gamma", C.c_void_p), ("beta", C.c_void_p), ("running_mean", C.c_void_p),
                ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p), ("c_begin", C.c_int32)]


class SyConvDesc(C.Structure):
    _fields_ = [("x", SyTensor), ("y", SyTensor), ("w", C.c_void_p), ("kh", C.c_int32), ("kw", C.c_int32),
                ("stride", C.c_int32), ("mode", C.c_int32), ("act", C.c_int32), ("scale", C.c_void_p),
                ("shift", C.c_void_p), ("res", SyTensor), ("split_n", C.c_int32), ("stat_partials", C.c_void_p),
                ("n_partials", C.c_int32), ("rows_written", C.POINTER(C.c_int32)), ("bn", SyBnSegment * 2),
                ("momentum", C.c_float), ("eps", C.c_float), ("scale_shift", C.c_void_p), ("mean_invstd", C.c_void_p),
                ("sync", C.c_void_p),
                ("apply_y", SyTensor), ("apply_res", SyTensor), ("apply_y_group1_offset", C.c_int64),
                ("apply_res_group1_offset", C.c_int64),
                ("debug_timeline", C.c_void_p),
                ("debug_timeline_events", C.c_int32), ("debug_flags", C.c_int32), ("debug_f32", C.c_void_p)]


class SyHeadPredDesc(C.Structure):
    _fields_ = [("cls_feat", SyTensor), ("reg_feat", SyTensor),
                ("w_reg", C.c_void_p), ("b_reg", C.c_void_p), ("w_obj", C.c_void_p), ("b_obj", C.c_void_p),
                ("w_cls", C.c_void_p), ("b_cls", C.c_void_p),
                ("num_classes", C.c_int32), ("stride", C.c_int32), ("anchor_offset", C.c_int32),
                ("a_total", C.c_int32), ("sigmoid", C.c_int32), ("decode", C.c_int32),
                ("out", C.c_void_p), ("origin", C.c_void_p)]


class SyTalLossDesc(C.Structure):
    _fields_ = [("b", C.c_int32), ("a_total", C.c_int32), ("max_labels", C.c_int32), ("num_classes", C.c_int32),
                ("n_levels", C.c_int32), ("level_h", C.c_int32 * 4), ("level_w", C.c_int32 * 4),
                ("level_stride", C.c_int32 * 4),
                ("outputs", C.c_void_p), ("origin", C.c_void_p), ("labels_fut", C.c_void_p),
                ("labels_cur", C.c_void_p), ("gamma", C.c_float), ("ignore_thr", C.c_float),
                ("ignore_value", C.c_float), ("use_l1", C.c_int32), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t), ("loss_out", C.c_void_p), ("fg_out", C.c_void_p),
                ("matched_out", C.c_void_p), ("pred_iou_out", C.c_void_p)]


class SyConvWgradDesc(C.Structure):
    _fields_ = [("x", SyTensor), ("dy", SyTensor), ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32),
                ("dw", C.c_void_p), ("accumulate", C.c_int32), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t)]


class SyNmsDesc(C.Structure):
    _fields_ = [("pred", C.c_void_p), ("b", C.c_int32), ("a_total", C.c_int32), ("num_classes", C.c_int32),
                ("max_det", C.c_int32), ("conf_thre", C.c_float), ("nms_thre", C.c_float), ("class_agnostic", C.c_int32),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("det_out", C.c_void_p),
                ("count_out", C.c_void_p)]


class SyBnActBwdDesc(C.Structure):
    _fields_ = [("raw", SyTensor), ("dy", SyTensor), ("draw", SyTensor), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("mean", C.c_void_p), ("invstd", C.c_void_p), ("split_n", C.c_int32), ("act", C.c_int32),
                ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("accumulate", C.c_int32), ("partials", C.c_void_p),
                ("n_partials", C.c_int32), ("coef", C.c_void_p)]


class SyHeadPredBwdDesc(C.Structure):
    _fields_ = [("grad_raw", C.c_void_p), ("cls_feat", SyTensor), ("reg_feat", SyTensor), ("d_cls_feat", SyTensor),
                ("d_reg_feat", SyTensor), ("w_reg", C.c_void_p), ("w_obj", C.c_void_p), ("w_cls", C.c_void_p),
                ("num_classes", C.c_int32), ("a_total", C.c_int32), ("anchor_offset", C.c_int32),
                ("dw_reg", C.c_void_p), ("dw_obj", C.c_void_p), ("dw_cls", C.c_void_p), ("db_reg", C.c_void_p),
                ("db_obj", C.c_void_p), ("db_cls", C.c_void_p), ("accumulate", C.c_int32), ("partials", C.c_void_p),
                ("n_partials", C.c_int32)]


class SySgdEmaDesc(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("momentum_buf", C.c_void_p), ("ema", C.c_void_p),
                ("n_param", C.c_int64), ("n_total", C.c_int64), ("decay_begin", C.c_int64),
                ("lr", C.c_float), ("momentum", C.c_float), ("weight_decay", C.c_float), ("inv_scale", C.c_float),
                ("nesterov", C.c_int32), ("ema_decay", C.c_float), ("ema_one_minus_decay", C.c_float),
                ("found_inf", C.c_void_p), ("hyper", C.c_void_p)]


class SyPackItem(C.Structure):
    _fields_ = [("w", C.c_void_p), ("out", C.c_void_p), ("cout", C.c_int32), ("cin", C.c_int32), ("kh", C.c_int32),
                ("taps", C.c_int32), ("mode", C.c_int32), ("co_offset", C.c_int32), ("out_pitch", C.c_int64),
                ("begin", C.c_int64)]


class SyConvPlan(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("mode", "bn", "m_tiles", "n_tiles", "rounds", "kblocks", "patch_h", "patch_w")]


class SyTalLossBwdDesc(C.Structure):
    _fields_ = [("outputs", C.c_void_p), ("origin", C.c_void_p), ("labels_fut", C.c_void_p),
                ("b", C.c_int32), ("a_total", C.c_int32), ("max_labels", C.c_int32), ("num_classes", C.c_int32),
                ("n_levels", C.c_int32), ("level_h", C.c_int32 * 4), ("level_w", C.c_int32 * 4),
                ("level_stride", C.c_int32 * 4), ("gamma", C.c_float), ("use_l1", C.c_int32),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("grad_scale", C.c_float),
                ("grad_outputs", C.c_void_p), ("grad_origin", C.c_void_p), ("grad_raw", C.c_void_p)]


# every symbol include/streamyolo_sm100.h declares: (restype, argtypes)
_SIG = {
    "sy_last_error_string": (C.c_char_p, []),
    "sy_version": (C.c_int, []),
    "sy_check_device": (C.c_int, []),
    "sy_conv_stat_rows": (C.c_int, []),
    "sy_l2_persist_window": (C.c_int, [C.c_void_p, C.c_size_t, C.c_float, C.POINTER(C.c_size_t), C.c_void_p]),
    "sy_conv2d_tc": (C.c_int, [C.POINTER(SyConvDesc), C.c_void_p]),
    "sy_conv2d_plan": (C.c_int, [C.c_int32] * 8 + [C.POINTER(SyConvPlan)]),
    "sy_conv2d_simt": (C.c_int, [C.POINTER(SyConvDesc), C.c_void_p]),
    "sy_dwconv2d": (C.c_int, [C.POINTER(SyConvDesc), C.c_void_p]),
    "sy_focus_pack": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, SyTensor,
                                C.c_void_p]),
    "sy_stats_num_partials": (C.c_int, [C.c_int32, C.c_int32]),
    "sy_channel_stats": (C.c_int, [SyTensor, C.c_void_p, C.c_int32, C.c_void_p]),
    "sy_bn_finalize": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "sy_bn_act_apply": (C.c_int, [SyTensor, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, SyTensor, SyTensor,
                                  C.c_int64, C.c_int64, C.c_void_p]),
    "sy_upsample_nearest": (C.c_int, [SyTensor, SyTensor, C.c_void_p]),
    "sy_spp_maxpool": (C.c_int, [SyTensor, SyTensor, SyTensor, SyTensor, C.c_void_p]),
    "sy_copy": (C.c_int, [SyTensor, SyTensor, C.c_void_p]),
    "sy_head_pred_decode": (C.c_int, [C.POINTER(SyHeadPredDesc), C.c_void_p]),
    "sy_tal_loss_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "sy_tal_loss": (C.c_int, [C.POINTER(SyTalLossDesc), C.c_void_p]),
    "sy_tal_loss_backward": (C.c_int, [C.POINTER(SyTalLossBwdDesc), C.c_void_p]),
    "sy_add": (C.c_int, [SyTensor, SyTensor, C.c_void_p]),
    "sy_spp_maxpool_backward_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "sy_spp_maxpool_backward": (C.c_int, [SyTensor, SyTensor, SyTensor, SyTensor, SyTensor, C.c_void_p, C.c_size_t,
                                          C.c_void_p]),
    "sy_dilate2": (C.c_int, [SyTensor, SyTensor, C.c_void_p]),
    "sy_upsample_nearest_backward": (C.c_int, [SyTensor, SyTensor, C.c_void_p]),
    "sy_head_pred_bwd_rows": (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    "sy_head_pred_backward": (C.c_int, [C.POINTER(SyHeadPredBwdDesc), C.c_void_p]),
    "sy_bn_act_bwd_rows": (C.c_int, [C.c_int32, C.c_int32]),
    "sy_bn_act_backward": (C.c_int, [C.POINTER(SyBnActBwdDesc), C.c_void_p]),
    "sy_postprocess_nms_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "sy_postprocess_nms": (C.c_int, [C.POINTER(SyNmsDesc), C.c_void_p]),
    "sy_conv2d_wgrad_workspace_bytes": (C.c_size_t, [C.POINTER(SyConvWgradDesc)]),
    "sy_conv2d_wgrad_tc": (C.c_int, [C.POINTER(SyConvWgradDesc), C.c_void_p]),
    "sy_pack_conv_weight": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_int64, C.c_int32, C.c_void_p]),
    "sy_pack_item_tiles": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "sy_pack_conv_weights_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "sy_sgd_nesterov_ema_step": (C.c_int, [C.POINTER(SySgdEmaDesc), C.c_void_p]),
    "sy_resize_bilinear": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_void_p]),
    "sy_scale_labels": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIG)

_lib = None
_device_ok = False
LAUNCHES = 0      # kernels launched through this binding (bench.py reports it as gpu_launches)


def load_library():
    """dlopen the in-tree library (no GPU needed) and type every entry point."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m streamyolo_b200.build` "
                               "(there is no fallback path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIG.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def lib():
    """Library handle for compute calls: also insists on a CUDA sm_100 device."""
    global _device_ok
    l = load_library()
    if not _device_ok:
        if not torch.cuda.is_available():
            raise RuntimeError("streamyolo_b200 needs a CUDA sm_100 (B200) device; there is no CPU path")
        rc = l.sy_check_device()
        if rc != 0:
            raise RuntimeError("streamyolo_b200: " + l.sy_last_error_string().decode())
        _device_ok = True
    return l


def _check(rc, kernels=1):
    global LAUNCHES
    LAUNCHES += kernels
    if rc != 0:
        raise RuntimeError(f"libstreamyolo_sm100 error {rc}: " + load_library().sy_last_error_string().decode())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


NULL_T = SyTensor(None, 0, 0, 0, 0, 0)


class View:
    """Channel-slice / image-slice view of an NHWC bf16 buffer ``buf[N,H,W,Ctot]``."""
    __slots__ = ("buf", "n0", "n", "c0", "c", "off")

    def __init__(self, buf, c0=0, c=None, n0=0, n=None):
        assert buf.dtype == torch.bfloat16 and buf.dim() == 4 and buf.is_contiguous()
        self.buf, self.c0, self.n0, self.off = buf, c0, n0, 0
        self.c = buf.shape[3] - c0 if c is None else c
        self.n = buf.shape[0] - n0 if n is None else n

    @staticmethod
    def empty(n, h, w, c, device):
        return View(torch.empty((n, h, w, c), dtype=torch.bfloat16, device=device))

    @property
    def h(self):
        return self.buf.shape[1]

    @property
    def w(self):
        return self.buf.shape[2]

    def ch(self, c0, c):
        return View(self.buf, self.c0 + c0, c, self.n0, self.n)

    def imgs(self, n0, n):
        return View(self.buf, self.c0, self.c, self.n0 + n0, n)

    def img_elems(self):
        b = self.buf
        return b.shape[1] * b.shape[2] * b.shape[3]

    def shifted(self, elems):
        """Same shape, base address moved by ``elems`` elements (used for group-offset destinations)."""
        v = View(self.buf, self.c0, self.c, self.n0, self.n)
        v.off = getattr(self, "off", 0) + elems
        return v

    def st(self):
        b = self.buf
        ptr = b.data_ptr() + 2 * (self.n0 * b.shape[1] * b.shape[2] * b.shape[3] + self.c0 + self.off)
        return SyTensor(ptr, self.n, b.shape[1], b.shape[2], self.c, b.shape[3])

    def torch(self):
        """NHWC torch view (for tests)."""
        return self.buf[self.n0:self.n0 + self.n, :, :, self.c0:self.c0 + self.c]

    def nchw_float(self):
        return self.torch().permute(0, 3, 1, 2).float()


def from_nchw(x):
    """float NCHW torch tensor -> bf16 NHWC View (test helper)."""
    return View(x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))


def _w32(w):
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    return w


def pack_conv_weight(*ws):
    """OIHW fp32 parameter(s) -> bf16 [sum O][kh*kw][I] contiguous (GEMM B operand, K-major), packed on the device
    (sy_pack_conv_weight).  Several weights with the same [I, kh, kw] (CSPLayer conv1 | conv2) land in one operand."""
    _, i, kh, kw = ws[0].shape
    out = torch.empty((sum(w.shape[0] for w in ws), kh * kw, i), dtype=torch.bfloat16, device=ws[0].device)
    o0 = 0
    for w in ws:
        w = _w32(w)
        assert tuple(w.shape[1:]) == (i, kh, kw)
        _check(lib().sy_pack_conv_weight(w.data_ptr(), w.shape[0], i, kh, kw, 0, out.data_ptr() + 2 * o0 * kh * kw * i, 0, 0,
                                         _stream()))
        o0 += w.shape[0]
    return out


def conv_out_hw(h, w, k, s):
    p = (k - 1) // 2
    return (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1


def conv_stat_rows():
    return load_library().sy_conv_stat_rows()


def conv2d(x: View, wpk, y: View, k, s, mode, impl="tc", scale=None, shift=None, act=1, res: View = None,
           partials=None, split_n=0, timeline=None, debug_flags=0, bn=None, momentum=0.03, eps=1e-3, scale_shift=None,
           sync=None, apply_y: View = None, apply_res: View = None, y_goff1=0, res_goff1=0, mean_invstd=None, debug_f32=None):
    """``k`` is an int (square) or (kh, kw).  With ``partials`` (RAW mode, tensor-core path) returns the number
    of per-CTA statistic rows the launch writes."""
    d = SyConvDesc()
    d.x, d.y = x.st(), y.st()
    d.w = wpk.data_ptr()
    d.kh, d.kw = (k, k) if isinstance(k, int) else k
    d.stride, d.mode, d.act = s, mode, act
    d.scale = scale.data_ptr() if scale is not None else None
    d.shift = shift.data_ptr() if shift is not None else None
    d.res = res.st() if res is not None else NULL_T
    d.split_n = split_n
    rows = C.c_int32(0)
    if partials is not None:
        d.stat_partials, d.n_partials = partials.data_ptr(), partials.shape[0]
        d.rows_written = C.pointer(rows)
    if bn:
        for i, (g, b_, rm, rv, nbt, c0) in enumerate(bn):
            seg = d.bn[i]
            seg.gamma, seg.beta = g.data_ptr(), b_.data_ptr()
            seg.running_mean = rm.data_ptr() if rm is not None else None
            seg.running_var = rv.data_ptr() if rv is not None else None
            seg.num_batches_tracked = nbt.data_ptr() if nbt is not None else None
            seg.c_begin = c0
        d.momentum, d.eps = momentum, eps
        d.scale_shift, d.sync = scale_shift.data_ptr(), sync.data_ptr()
        d.mean_invstd = mean_invstd.data_ptr() if mean_invstd is not None else None
        if apply_y is not None:
            d.apply_y = apply_y.st()
            d.apply_res = apply_res.st() if apply_res is not None else NULL_T
            d.apply_y_group1_offset, d.apply_res_group1_offset = y_goff1, res_goff1
    d.debug_flags = debug_flags
    d.debug_f32 = debug_f32.data_ptr() if debug_f32 is not None else None
    if timeline is not None:
        d.debug_timeline, d.debug_timeline_events = timeline.data_ptr(), timeline.numel() // 2
    fn = {"tc": lib().sy_conv2d_tc, "simt": lib().sy_conv2d_simt, "dw": lib().sy_dwconv2d}[impl]
    _check(fn(C.byref(d), _stream()))
    return rows.value


def focus_pack(x, frames, y: View):
    assert x.dtype == torch.float32 and x.is_contiguous()
    b, ch, h, w = x.shape
    _check(lib().sy_focus_pack(x.data_ptr(), b, ch, h, w, frames, y.st(), _stream()))


STEM_K = (3, 1)      # the stem runs as a 3x1 conv over the W-gathered 64-channel focus tensor


def pack_dw_weight(w):
    """depthwise [C, 1, k, k] fp32 parameter -> bf16 [k*k][C] (sy_pack_conv_weight mode 0 on the [1, C, k, k] view)."""
    c, one, kh, kw = w.shape
    assert one == 1
    return pack_conv_weight(_w32(w).view(1, c, kh, kw)).view(kh * kw, c)


def pack_stem_weight(w):
    """[O,12,3,3] float -> bf16 [O][3 (row)][64 = 3 taps x (12 focus + 4 zero) + 16 zero] (sy_pack_conv_weight, mode 2)."""
    w = _w32(w)
    o, i, kh, kw = w.shape
    out = torch.empty((o, kh, 64), dtype=torch.bfloat16, device=w.device)
    _check(lib().sy_pack_conv_weight(w.data_ptr(), o, i, kh, kw, 2, out.data_ptr(), 0, 0, _stream()))
    return out


def stats_num_partials(n, hw):
    return load_library().sy_stats_num_partials(n, hw)


def channel_stats(x: View, partials):
    _check(lib().sy_channel_stats(x.st(), partials.data_ptr(), partials.shape[0], _stream()))


def bn_finalize(partials, p_split, groups, count, gamma, beta, rmean, rvar, nbt, momentum, eps, scale, shift):
    c = gamma.numel()
    _check(lib().sy_bn_finalize(partials.data_ptr(), partials.shape[0], p_split, groups, count, c,
                                gamma.data_ptr(), beta.data_ptr(),
                                rmean.data_ptr() if rmean is not None else None,
                                rvar.data_ptr() if rvar is not None else None,
                                nbt.data_ptr() if nbt is not None else None,
                                momentum, eps, scale.data_ptr(), shift.data_ptr(), _stream()))


def bn_act_apply(x: View, scale_ptr, shift_ptr, split_n, act, res, y: View, y_goff1=0, res_goff1=0):
    """``scale_ptr`` / ``shift_ptr``: fp32 [groups][C] tensors (or their raw device addresses)"""
    if torch.is_tensor(scale_ptr):
        scale_ptr, shift_ptr = scale_ptr.data_ptr(), shift_ptr.data_ptr()
    _check(lib().sy_bn_act_apply(x.st(), scale_ptr, shift_ptr, split_n, act,
                                 res.st() if res is not None else NULL_T, y.st(), y_goff1, res_goff1, _stream()))


def upsample_nearest(x: View, y: View):
    _check(lib().sy_upsample_nearest(x.st(), y.st(), _stream()))


def spp_maxpool(x: View, y5: View, y9: View, y13: View):
    _check(lib().sy_spp_maxpool(x.st(), y5.st(), y9.st(), y13.st(), _stream()))


def copy(x: View, y: View):
    _check(lib().sy_copy(x.st(), y.st(), _stream()))


def head_pred_decode(cls_feat: View, reg_feat: View, w_reg, b_reg, w_obj, b_obj, w_cls, b_cls, stride,
                     anchor_offset, a_total, out, origin, sigmoid, decode):
    d = SyHeadPredDesc()
    d.cls_feat, d.reg_feat = cls_feat.st(), reg_feat.st()
    d.w_reg, d.b_reg, d.w_obj, d.b_obj = w_reg.data_ptr(), b_reg.data_ptr(), w_obj.data_ptr(), b_obj.data_ptr()
    d.w_cls, d.b_cls = w_cls.data_ptr(), b_cls.data_ptr()
    d.num_classes = w_cls.shape[0]
    d.stride, d.anchor_offset, d.a_total = stride, anchor_offset, a_total
    d.sigmoid, d.decode = int(sigmoid), int(decode)
    d.out = out.data_ptr()
    d.origin = origin.data_ptr() if origin is not None else None
    _check(lib().sy_head_pred_decode(C.byref(d), _stream()))


def tal_loss_workspace_bytes(b, a_total, max_labels, num_classes):
    return load_library().sy_tal_loss_workspace_bytes(b, a_total, max_labels, num_classes)


def tal_loss(outputs, origin, labels_fut, labels_cur, hw, strides, gamma, ignore_thr, ignore_value, use_l1,
             workspace, loss_out, fg_out=None, matched_out=None, pred_iou_out=None):
    d = SyTalLossDesc()
    b, a, no = outputs.shape
    d.b, d.a_total, d.max_labels, d.num_classes = b, a, labels_fut.shape[1], no - 5
    d.n_levels = len(hw)
    for i, ((h, w), s) in enumerate(zip(hw, strides)):
        d.level_h[i], d.level_w[i], d.level_stride[i] = h, w, s
    d.outputs = outputs.data_ptr()
    d.origin = origin.data_ptr() if origin is not None else None
    d.labels_fut, d.labels_cur = labels_fut.data_ptr(), labels_cur.data_ptr()
    d.gamma, d.ignore_thr, d.ignore_value, d.use_l1 = gamma, ignore_thr, ignore_value, int(use_l1)
    d.workspace, d.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    d.loss_out = loss_out.data_ptr()
    d.fg_out = fg_out.data_ptr() if fg_out is not None else None
    d.matched_out = matched_out.data_ptr() if matched_out is not None else None
    d.pred_iou_out = pred_iou_out.data_ptr() if pred_iou_out is not None else None
    _check(lib().sy_tal_loss(C.byref(d), _stream()), kernels=6)


def tal_loss_backward(outputs, origin, labels_fut, hw, strides, gamma, use_l1, workspace, grad_scale=1.0,
                      grad_outputs=None, grad_origin=None, grad_raw=None):
    """Gradient of total_loss w.r.t. the head outputs; run after tal_loss() on the same workspace."""
    d = SyTalLossBwdDesc()
    b, a, no = outputs.shape
    d.outputs = outputs.data_ptr()
    d.origin = origin.data_ptr() if origin is not None else None
    d.labels_fut = labels_fut.data_ptr()
    d.b, d.a_total, d.max_labels, d.num_classes = b, a, labels_fut.shape[1], no - 5
    d.n_levels = len(hw)
    for i, ((h, w), s) in enumerate(zip(hw, strides)):
        d.level_h[i], d.level_w[i], d.level_stride[i] = h, w, s
    d.gamma, d.use_l1 = gamma, int(use_l1)
    d.workspace, d.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    d.grad_scale = grad_scale
    d.grad_outputs = grad_outputs.data_ptr() if grad_outputs is not None else None
    d.grad_origin = grad_origin.data_ptr() if grad_origin is not None else None
    d.grad_raw = grad_raw.data_ptr() if grad_raw is not None else None
    _check(lib().sy_tal_loss_backward(C.byref(d), _stream()), kernels=1)


def pack_conv_weight_dgrad(*ws):
    """Weights for the data gradient of a stride-1 conv: dx = conv(dy, w') with w'[ci][kh-1-r][kw-1-s][co] = w[co][ci][r][s]
    (same padding), i.e. the forward tensor-core kernel on the flipped, channel-transposed filter; packed on the device
    (sy_pack_conv_weight, mode 1).  Several weights (CSPLayer conv1 | conv2) are concatenated along co."""
    _, i, kh, kw = ws[0].shape
    ot = sum(w.shape[0] for w in ws)
    out = torch.empty((i, kh * kw, ot), dtype=torch.bfloat16, device=ws[0].device)
    o0 = 0
    for w in ws:
        w = _w32(w)
        _check(lib().sy_pack_conv_weight(w.data_ptr(), w.shape[0], i, kh, kw, 1, out.data_ptr(), ot, o0, _stream()))
        o0 += w.shape[0]
    return out


def conv2d_wgrad(x: View, dy: View, k, s, dw, accumulate=False, workspace=None):
    """dw[cout, cin, kh, kw] (fp32) (+)= weight gradient of the conv that maps x to (the shape of) dy."""
    kh, kw = (k, k) if isinstance(k, int) else k
    assert dw.dtype == torch.float32 and dw.is_contiguous() and tuple(dw.shape) == (dy.c, x.c, kh, kw)
    d = SyConvWgradDesc()
    d.x, d.dy = x.st(), dy.st()
    d.kh, d.kw, d.stride = kh, kw, s
    d.dw, d.accumulate = dw.data_ptr(), int(accumulate)
    need = lib().sy_conv2d_wgrad_workspace_bytes(C.byref(d))
    if need == 0:
        raise RuntimeError("conv2d_wgrad: " + (lib().sy_last_error_string() or b"").decode())
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dw.device)
    d.workspace, d.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    _check(lib().sy_conv2d_wgrad_tc(C.byref(d), _stream()), kernels=2)
    return workspace


def postprocess_nms(pred, num_classes, conf_thre, nms_thre, class_agnostic=False, max_det=None):
    """-> (det [B, max_det, 7] fp32, count [B] int32) on the device; rows [x1, y1, x2, y2, obj, class_conf, class_pred]."""
    assert pred.dtype == torch.float32 and pred.is_contiguous() and pred.shape[2] == 5 + num_classes
    b, a, _ = pred.shape
    max_det = a if max_det is None else max_det
    det = torch.empty((b, max_det, 7), dtype=torch.float32, device=pred.device)
    count = torch.empty((b,), dtype=torch.int32, device=pred.device)
    ws = torch.empty(load_library().sy_postprocess_nms_workspace_bytes(b, a), dtype=torch.uint8, device=pred.device)
    d = SyNmsDesc()
    d.pred, d.b, d.a_total, d.num_classes, d.max_det = pred.data_ptr(), b, a, num_classes, max_det
    d.conf_thre, d.nms_thre, d.class_agnostic = conf_thre, nms_thre, int(class_agnostic)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    d.det_out, d.count_out = det.data_ptr(), count.data_ptr()
    _check(lib().sy_postprocess_nms(C.byref(d), _stream()), kernels=1)
    return det, count


def bn_act_backward(raw: View, dy: View, draw: View, scale, shift, mean, invstd, split_n, act, dgamma, dbeta,
                    accumulate=False):
    """BatchNorm(train) + SiLU backward of one BaseConv; scale/shift/mean/invstd: fp32 [2, C] (group-major)."""
    rows = load_library().sy_bn_act_bwd_rows(raw.n, raw.h * raw.w)
    partials = torch.empty((rows, 2 * raw.c), dtype=torch.float32, device=dgamma.device)
    coef = torch.empty((8 * raw.c,), dtype=torch.float32, device=dgamma.device)
    d = SyBnActBwdDesc()
    d.raw, d.dy, d.draw = raw.st(), dy.st(), draw.st()
    d.scale, d.shift, d.mean, d.invstd = scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr()
    d.split_n, d.act = split_n, int(act)
    d.dgamma, d.dbeta, d.accumulate = dgamma.data_ptr(), dbeta.data_ptr(), int(accumulate)
    d.partials, d.n_partials, d.coef = partials.data_ptr(), rows, coef.data_ptr()
    _check(lib().sy_bn_act_backward(C.byref(d), _stream()), kernels=3)


def dilate2(g: View, D: View):
    _check(lib().sy_dilate2(g.st(), D.st(), _stream()))


def conv2d_dgrad_stride2(dy: View, w, dx: View, k=3):
    """Data gradient of a stride-2 conv: zero-insert dy to dx's spatial size, then the stride-1 forward kernel on the
    flipped, channel-transposed filter (``w`` = the forward OIHW weights)."""
    D = View.empty(dx.n, dx.h, dx.w, dy.c, dy.buf.device)
    dilate2(dy, D)
    conv2d(D, pack_conv_weight_dgrad(w), dx, k, 1, SY_CONV_RAW)


def upsample_nearest_backward(dy: View, dx: View):
    _check(lib().sy_upsample_nearest_backward(dy.st(), dx.st(), _stream()))


def head_pred_backward(grad_raw, cls_feat: View, reg_feat: View, d_cls_feat: View, d_reg_feat: View, w_reg, w_obj, w_cls,
                       a_total, anchor_offset, dw_reg, dw_obj, dw_cls, db_reg, db_obj, db_cls, accumulate=False):
    nc = w_cls.shape[0]
    rows = load_library().sy_head_pred_bwd_rows(cls_feat.n, cls_feat.h, cls_feat.w)
    partials = torch.empty((rows, (5 + nc) * (cls_feat.c + 1)), dtype=torch.float32, device=grad_raw.device)
    d = SyHeadPredBwdDesc()
    d.grad_raw = grad_raw.data_ptr()
    d.cls_feat, d.reg_feat, d.d_cls_feat, d.d_reg_feat = cls_feat.st(), reg_feat.st(), d_cls_feat.st(), d_reg_feat.st()
    d.w_reg, d.w_obj, d.w_cls = w_reg.data_ptr(), w_obj.data_ptr(), w_cls.data_ptr()
    d.num_classes, d.a_total, d.anchor_offset = nc, a_total, anchor_offset
    d.dw_reg, d.dw_obj, d.dw_cls = dw_reg.data_ptr(), dw_obj.data_ptr(), dw_cls.data_ptr()
    d.db_reg, d.db_obj, d.db_cls = db_reg.data_ptr(), db_obj.data_ptr(), db_cls.data_ptr()
    d.accumulate, d.partials, d.n_partials = int(accumulate), partials.data_ptr(), rows
    _check(lib().sy_head_pred_backward(C.byref(d), _stream()), kernels=3)


def add_(x: View, y: View):
    """y += x"""
    _check(lib().sy_add(x.st(), y.st(), _stream()))


def spp_maxpool_backward(x: View, d5: View, d9: View, d13: View, dx: View):
    ws = torch.empty(load_library().sy_spp_maxpool_backward_workspace_bytes(x.n, x.h, x.w, x.c), dtype=torch.uint8,
                     device=x.buf.device)
    _check(lib().sy_spp_maxpool_backward(x.st(), d5.st(), d9.st(), d13.st(), dx.st(), ws.data_ptr(), ws.numel(), _stream()),
           kernels=2)


def conv2d_plan(n, h, w, cin, cout, k, s):
    """Tiling decisions of the tensor-core conv for a layer shape (host-only: works without a GPU)."""
    kh, kw = (k, k) if isinstance(k, int) else k
    p = SyConvPlan()
    rc = load_library().sy_conv2d_plan(n, h, w, cin, cout, kh, kw, s, C.byref(p))
    if rc != 0:
        raise RuntimeError("conv2d_plan: " + (load_library().sy_last_error_string() or b"").decode())
    return {f: getattr(p, f) for f, _ in SyConvPlan._fields_}


def sgd_nesterov_ema_step(param, grad, momentum_buf, ema, n_param, decay_begin, lr, momentum=0.9, weight_decay=5e-4,
                          inv_scale=1.0, nesterov=True, ema_decay=0.0, found_inf=None, hyper=None):
    """One fused optimiser step over flat fp32 state (sy_sgd_nesterov_ema_step); ``ema`` may be None.  ``hyper``: device
    fp32 [lr, momentum, weight_decay, inv_scale, ema_decay, 1 - ema_decay] replacing the scalars (CUDA-graph replays)."""
    d = SySgdEmaDesc()
    d.param, d.grad, d.momentum_buf = param.data_ptr(), grad.data_ptr(), momentum_buf.data_ptr()
    d.ema = ema.data_ptr() if ema is not None else None
    d.n_param, d.n_total, d.decay_begin = n_param, param.numel(), decay_begin
    d.lr, d.momentum, d.weight_decay, d.inv_scale, d.nesterov = lr, momentum, weight_decay, inv_scale, int(nesterov)
    d.ema_decay, d.ema_one_minus_decay = ema_decay, 1.0 - ema_decay
    d.found_inf = found_inf.data_ptr() if found_inf is not None else None
    d.hyper = hyper.data_ptr() if hyper is not None else None
    _check(lib().sy_sgd_nesterov_ema_step(C.byref(d), _stream()))


def resize_bilinear(x, size):
    """F.interpolate(x, size=size, mode="bilinear", align_corners=False) of an NCHW fp32 batch on the device."""
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    b, c, hi, wi = x.shape
    y = torch.empty((b, c, size[0], size[1]), dtype=torch.float32, device=x.device)
    _check(lib().sy_resize_bilinear(x.data_ptr(), b * c, hi, wi, y.data_ptr(), size[0], size[1], _stream()))
    return y


def scale_labels_(labels, sx, sy):
    """labels[..., 1::2] *= sx; labels[..., 2::2] *= sy (in place; [.., cols] fp32 contiguous)."""
    assert labels.dtype == torch.float32 and labels.is_contiguous()
    cols = labels.shape[-1]
    _check(lib().sy_scale_labels(labels.data_ptr(), labels.numel() // cols, cols, sx, sy, _stream()))
    return labels


def l2_persist_window(t, hit_ratio=1.0):
    """Persisting-L2 window over tensor ``t`` (None: clear) for the kernels launched on the current stream; returns the
    number of bytes the device granted."""
    got = C.c_size_t(0)
    if t is None:
        _check(lib().sy_l2_persist_window(None, 0, 0.0, C.byref(got), _stream()), kernels=0)
        return 0
    _check(lib().sy_l2_persist_window(t.data_ptr(), t.numel() * t.element_size(), hit_ratio, C.byref(got), _stream()), kernels=0)
    return got.value


class PackBatch:
    """Every conv operand of a model re-packed in ONE launch (sy_pack_conv_weights_batch).  ``add`` the (fp32 parameter,
    bf16 destination, layout) pairs once -- the tensors must keep their addresses -- then ``run()`` after every update."""

    def __init__(self, device):
        self.device, self.items, self.total, self.table = device, [], 0, None

    def add(self, w, out, mode, out_pitch=0, co_offset=0):
        assert w.dtype == torch.float32 and w.is_contiguous() and out.dtype == torch.bfloat16
        o, i, kh, kw = w.shape
        it = SyPackItem(w.data_ptr(), out.data_ptr(), o, i, kh, kh * kw, mode, co_offset, out_pitch, self.total)
        self.items.append(it)
        self.total += load_library().sy_pack_item_tiles(o, i, mode)      # work tiles (see include/streamyolo_sm100.h)

    def run(self):
        if self.table is None:
            arr = (SyPackItem * len(self.items))(*self.items)
            self.table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
        _check(lib().sy_pack_conv_weights_batch(self.table.data_ptr(), len(self.items), self.total, _stream()))
