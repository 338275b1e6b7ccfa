"""GPU: the cta_group::2 ("pair") variant of the tensor-core convolution (conv_tc.cu, PAIR): two CTAs of one cluster compute a
256-pixel x 256-channel tile with one tcgen05.mma.cta_group::2 stream.  SY_CONV_PAIR=1 forces it on every eligible layer
(linear or halo tiles, any tile width), SY_CONV_PAIR=0 switches it off; the two must agree with each other and with F.conv2d on
  * the fp32 accumulators (debug store), the stored bf16 values, the per-channel statistics and the BatchNorm finalize,
  * pair tiles whose second half lies past the end of the tensor, channel counts that leave the peer's weight half empty,
  * more pair tiles than resident pairs (several rounds: barrier phases wrap), repeated launches (grid barrier counters)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from streamyolo_b200 import ops  # noqa: E402
from streamyolo_b200.ops import View  # noqa: E402
from test_gpu_ops import DEV, check_close, rand_act, rand_w  # noqa: E402

CASES = [
    # (n, cin, cout, h, w, k, s), tile width BN, A mode
    ((4, 256, 256, 38, 60, 3, 1), 256, "off"),       # 72 M tiles -> 36 pair tiles
    ((16, 256, 256, 38, 60, 3, 1), 256, "off"),      # 285 M tiles -> 143 pair tiles: two rounds of 74 pairs
    ((3, 256, 512, 19, 30, 3, 1), 256, "off"),       # 14 M tiles, two N tiles
    ((3, 128, 512, 21, 30, 3, 1), 256, "off"),       # 15 M tiles (odd): the last pair's second half is past the end
    ((2, 512, 1024, 38, 60, 3, 2), 256, "off"),      # stride 2
    ((1, 2048, 1024, 19, 30, 1, 1), 256, "off"),     # 1x1, 32 K blocks
    ((2, 256, 384, 19, 30, 3, 1), 256, "off"),       # Cout = 384: the second N tile's upper weight half is empty
    ((1, 256, 256, 9, 7, 3, 1), 256, "off"),         # 63 pixels: one pair tile, the peer CTA has no pixel at all
    ((2, 64, 128, 150, 240, 3, 2), 128, "off"),      # linear tiles, BN = 128 (each CTA stages 64 weight rows)
    ((4, 128, 128, 38, 60, 1, 1), 128, "off"),
    ((2, 64, 64, 75, 120, 1, 1), 64, "off"),         # BN = 64: 32 weight rows per CTA
    ((2, 192, 96, 19, 30, 3, 1), 128, "off"),        # Cout = 96 < BN: the peer's weight half is partly out of bounds
    ((4, 128, 128, 75, 120, 3, 1), 128, "halo"),     # halo tiles (16 x 8 patches): a pair = two patches, two halos
    ((2, 64, 64, 150, 240, 3, 1), 64, "halo"),
    ((3, 128, 128, 19, 30, 3, 1), 128, "halo"),      # ragged patches on a small map, odd patch count per image
    ((2, 256, 256, 38, 60, 3, 1), 256, "halo"),
    ((2, 96, 96, 19, 30, 3, 1), 128, "halo"),        # channel counts that are not multiples of 64
]


def n_m_tiles(case, amode):
    n, ci, co, h, w, k, s = case
    ho, wo = ops.conv_out_hw(h, w, k, s)
    return n * -(-ho // 16) * -(-wo // 8) if amode == "halo" else -(-n * ho * wo // 128)


def run(case, bn, amode, pair, monkeypatch, reps=1):
    n, ci, co, h, w, k, s = case
    monkeypatch.setenv("SY_CONV_PAIR", "1" if pair else "0")
    monkeypatch.setenv("SY_CONV_A", amode)
    monkeypatch.setenv("SY_CONV_BN", str(bn))
    x, wt = rand_act(n, ci, h, w, 71), rand_w(co, ci, k, 72)
    ho, wo = ops.conv_out_hw(h, w, k, s)
    g = torch.Generator().manual_seed(73)
    gamma, beta = (torch.rand(co, generator=g) + 0.5).to(DEV), (torch.rand(co, generator=g) - 0.5).to(DEV)
    rm, rv = torch.zeros(co, device=DEV), torch.ones(co, device=DEV)
    nbt = torch.zeros((), dtype=torch.long, device=DEV)
    y = View.empty(n, ho, wo, co, DEV)
    acc = torch.full((n * ho * wo, co), float("nan"), device=DEV)
    partials = torch.full((ops.conv_stat_rows(), 4 * co), float("nan"), device=DEV)
    ss = torch.empty((2, 2, co), device=DEV)
    mi = torch.empty((2, 2, co), device=DEV)
    sync = torch.zeros(4, dtype=torch.int32, device=DEV)
    split = n // 2
    xv, wp = ops.from_nchw(x), ops.pack_conv_weight(wt)
    for _ in range(reps):
        y.buf.fill_(float("nan"))
        rows = ops.conv2d(xv, wp, y, k, s, ops.SY_CONV_RAW, partials=partials, split_n=split, bn=[(gamma, beta, rm, rv, nbt, 0)],
                          momentum=0.03, eps=1e-3, scale_shift=ss, mean_invstd=mi, sync=sync, debug_f32=acc)
    torch.cuda.synchronize()
    assert sync.tolist() == [0, 0, 0, 0]
    ref = F.conv2d(x, wt, None, s, (k - 1) // 2)
    return dict(rows=rows, y=y.nchw_float(), acc=acc.view(n, ho, wo, co).permute(0, 3, 1, 2), ss=ss.clone(), mi=mi.clone(),
                rm=rm, rv=rv, nbt=int(nbt), ref=ref, groups=2 if 0 < split < n else 1)


@pytest.mark.parametrize("case,bn,amode", CASES, ids=lambda c: "x".join(map(str, c)) if isinstance(c, tuple) else str(c))
def test_pair_conv_matches_reference_and_single_cta(case, bn, amode, monkeypatch):
    n, ci, co, h, w, k, s = case
    ho, wo = ops.conv_out_hw(h, w, k, s)
    a = run(case, bn, amode, True, monkeypatch, reps=3)
    b = run(case, bn, amode, False, monkeypatch, reps=3)
    pair_tiles = -(-n_m_tiles(case, amode) // 2) * -(-co // bn)
    assert a["rows"] == 2 * min(74, pair_tiles), f"pair mode did not run: {a['rows']} statistic rows for {pair_tiles} pair tiles"
    check_close(a["y"], a["ref"], f"pair conv {case}")
    # fp32 accumulators: same K order, same operands -> within fp32 summation noise of the single-CTA kernel and of cuDNN
    rms = a["ref"].pow(2).mean().sqrt().item()
    assert torch.isfinite(a["acc"]).all()
    assert (a["acc"] - a["ref"]).abs().max().item() <= 1e-3 * rms * 8
    assert (a["acc"] - b["acc"]).abs().max().item() <= 1e-4 * rms
    assert (a["y"] != b["y"]).float().mean().item() < 1e-3
    # statistics, BatchNorm finalize (scale/shift, mean/invstd, running statistics after three launches)
    for key in ("ss", "mi"):
        assert torch.allclose(a[key][:, :a["groups"]], b[key][:, :a["groups"]], rtol=1e-4, atol=1e-5), key
    assert torch.allclose(a["rm"], b["rm"], rtol=1e-4, atol=1e-6) and torch.allclose(a["rv"], b["rv"], rtol=1e-4, atol=1e-6)
    assert a["nbt"] == b["nbt"] == 3 * a["groups"]
    # mean / invstd against the stored values
    st = a["y"]
    split = n // 2
    groups = [(0, split), (split, n)] if a["groups"] == 2 else [(0, n)]
    for gi, (i0, i1) in enumerate(groups):
        part = st[i0:i1]
        mean, var = part.mean((0, 2, 3)), part.var((0, 2, 3), unbiased=False)
        assert torch.allclose(a["mi"][0, gi], mean, rtol=1e-3, atol=1e-3 * rms)
        assert torch.allclose(a["mi"][1, gi], (var + 1e-3).rsqrt(), rtol=2e-3)


def test_pair_conv_fused_residual(monkeypatch):
    """FUSED epilogue (eval mode: folded BN scale/shift, SiLU, residual) in pair mode."""
    monkeypatch.setenv("SY_CONV_PAIR", "1")
    monkeypatch.setenv("SY_CONV_A", "off")
    monkeypatch.setenv("SY_CONV_BN", "256")
    n, ci, co, h, w = 3, 256, 256, 19, 30
    x, wt = rand_act(n, ci, h, w, 81), rand_w(co, ci, 3, 82)
    g = torch.Generator().manual_seed(83)
    scale, shift = (torch.rand(co, generator=g) + 0.5).to(DEV), (torch.rand(co, generator=g) - 0.5).to(DEV)
    resid = rand_act(n, co, h, w, 84)
    ref = F.silu(F.conv2d(x, wt, None, 1, 1) * scale[None, :, None, None] + shift[None, :, None, None]) + resid
    yv = ops.from_nchw(resid)
    ops.conv2d(ops.from_nchw(x), ops.pack_conv_weight(wt), yv, 3, 1, ops.SY_CONV_FUSED, scale=scale, shift=shift, act=1, res=yv)
    torch.cuda.synchronize()
    check_close(yv.nchw_float(), ref, "pair conv fused")


def test_pair_conv_in_graph_with_neighbours(monkeypatch):
    """Pair launches (clusters + programmatic dependent launch) between ordinary launches inside one CUDA graph, replayed."""
    monkeypatch.setenv("SY_CONV_A", "off")
    n, c, h, w = 4, 256, 38, 60
    x, w1, w3 = rand_act(n, c, h, w, 91), rand_w(c, c, 1, 92), rand_w(c, c, 3, 93)
    xv, p1, p3 = ops.from_nchw(x), ops.pack_conv_weight(w1), ops.pack_conv_weight(w3)
    t1, t2, t3 = (View.empty(n, h, w, c, DEV) for _ in range(3))

    def chain():
        ops.conv2d(xv, p1, t1, 1, 1, ops.SY_CONV_RAW)          # 1x1: epilogue-bound, never paired
        ops.conv2d(t1, p3, t2, 3, 1, ops.SY_CONV_RAW)          # 3x3 256 -> 256: paired by the heuristic
        ops.conv2d(t2, p1, t3, 1, 1, ops.SY_CONV_RAW)

    monkeypatch.setenv("SY_CONV_PAIR", "0")
    chain()
    torch.cuda.synchronize()
    want = t3.torch().clone()
    monkeypatch.delenv("SY_CONV_PAIR")
    st = torch.cuda.Stream()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        chain()
        torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=st):
            chain()
        for _ in range(3):
            t3.buf.zero_()
            gr.replay()
        torch.cuda.synchronize()
    assert (t3.torch().float() - want.float()).abs().max().item() <= 2.0 ** -6 * want.float().abs().max().item()
    assert (t3.torch() != want).float().mean().item() < 1e-2
