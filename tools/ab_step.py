"""A/B the whole training-forward step under different kernel switches in ONE process: for every setting the step is
re-captured as a CUDA graph (the switches are read at launch / capture time) and replayed.
Prints one line per setting: ms/step, pairs/s, loss (must agree across settings up to bf16 noise).

usage: python tools/ab_step.py [model] [pairs]    settings are listed in SETTINGS below"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from streamyolo_b200 import ops, synth
from streamyolo_b200.model import engine

tag = sys.argv[1] if len(sys.argv) > 1 else "l"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
model = bench.build_model(tag, dev)
x = synth.synth_frames(B, 600, 960, seed=1234).to(dev)
fut, cur = synth.synth_labels(B, 600, 960, seed=1)
fut, cur = fut.to(dev), cur.to(dev)

SETTINGS = [
    # (label, environment switches, fuse-apply threshold MB)
    ("base", {}, 0),
    ("team epilogue off", {"SY_CONV_TEAM": "0"}, 0),
    ("pair mode off (no cta_group::2)", {"SY_CONV_PAIR": "0"}, 0),
    ("pair mode forced on every BN=256 linear layer", {"SY_CONV_PAIR": "1"}, 0),
    ("fuse apply <= 10 MB", {}, 10),
    ("fuse apply <= 20 MB", {}, 20),
    ("fuse apply <= 40 MB", {}, 40),
    ("fuse apply <= 10 MB +pair", {"SY_PAIR_APPLY": "1"}, 10),
    ("fuse apply <= 20 MB +pair", {"SY_PAIR_APPLY": "1"}, 20),
    ("fuse apply <= 40 MB +pair", {"SY_PAIR_APPLY": "1"}, 40),
    ("head pred 1 px/thread", {"SY_HEAD_PT": "1"}, 0),
    ("head pred 4 px/thread", {"SY_HEAD_PT": "4"}, 0),
    ("apply: raw loads streaming (ld.cs)", {"SY_APPLY_HINTS": "1"}, 0),
    ("apply: stores .cg", {"SY_APPLY_HINTS": "2"}, 0),
    ("apply: ld.cs + st.cg", {"SY_APPLY_HINTS": "3"}, 0),
    ("raw arena 20 MB", {"SY_RAW_ARENA_MB": "20"}, 0),
    ("raw arena 60 MB", {"SY_RAW_ARENA_MB": "60"}, 0),
    ("raw arena 80 MB", {"SY_RAW_ARENA_MB": "80"}, 0),
    ("bn128 rule 300,4", {"SY_BN128_RULE": "300,4"}, 0),
    ("bn128 rule 300,8", {"SY_BN128_RULE": "300,8"}, 0),
    ("bn128 rule 300,16", {"SY_BN128_RULE": "300,16"}, 0),
    ("bn128 rule 600,8", {"SY_BN128_RULE": "600,8"}, 0),
    ("bn128 rule 1200,4", {"SY_BN128_RULE": "1200,4"}, 0),
    ("epi cycles 1500,1900,1900", {"SY_EPI_CYCLES": "1500,1900,1900"}, 0),
    ("epi cycles 1200,1900,1900", {"SY_EPI_CYCLES": "1200,1900,1900"}, 0),
    ("epi cycles 1900,1400,1900", {"SY_EPI_CYCLES": "1900,1400,1900"}, 0),
    ("epi cycles 1900,1900,1450", {"SY_EPI_CYCLES": "1900,1900,1450"}, 0),
    ("epi cycles 1900,1900,1000", {"SY_EPI_CYCLES": "1900,1900,1000"}, 0),
    ("epi cycles 1900,1900,700", {"SY_EPI_CYCLES": "1900,1900,700"}, 0),
    ("epi cycles 1900,1900,400", {"SY_EPI_CYCLES": "1900,1900,400"}, 0),
    ("epi cycles 1900,1900,100", {"SY_EPI_CYCLES": "1900,1900,100"}, 0),
    ("raw arena off (no L2 window)", {"SY_RAW_ARENA_MB": "0"}, 0),
    ("halo off", {"SY_CONV_A": "off"}, 0),
    ("halo forced", {"SY_CONV_A": "halo"}, 0),
    ("apply carveout default", {"SY_APPLY_CARVEOUT": "-1"}, 0),
    ("apply carveout 0", {"SY_APPLY_CARVEOUT": "0"}, 0),
    ("no tmem prefetch", {"SY_CONV_DEBUG": "32"}, 0),
    ("base again", {}, 0),
    ("one staging tile", {"SY_STAGE_TILES": "1"}, 0),
    ("two staging tiles", {"SY_STAGE_TILES": "2"}, 0),
    ("apply v1", {"SY_APPLY": "v1"}, 0),
    ("apply cap 3/SM", {"SY_APPLY_CAP": "3"}, 0),
    ("patch tiles", {"SY_CONV_TILES": "patch"}, 0),
    ("no pdl", {"SY_PDL": "0"}, 0),
    # marginal cost of whole kernel classes inside the graph (results are garbage, only the time counts)
    ("skip every normalise pass", {"SY_DBG_SKIP_APPLY": "0:1000000"}, 0),
    ("skip normalise <= 8 MB", {"SY_DBG_SKIP_APPLY": "0:8"}, 0),
    ("skip normalise 8-40 MB", {"SY_DBG_SKIP_APPLY": "8:40"}, 0),
    ("skip normalise > 40 MB", {"SY_DBG_SKIP_APPLY": "40:1000000"}, 0),
    ("conv without MMAs", {"SY_CONV_DEBUG": "1"}, 0),
    ("conv without TMA loads", {"SY_CONV_DEBUG": "2"}, 0),
    ("conv without MMAs and loads", {"SY_CONV_DEBUG": "3"}, 0),
    ("no applies, conv w/o MMAs+loads", {"SY_CONV_DEBUG": "3", "SY_DBG_SKIP_APPLY": "0:1000000"}, 0),
]
if len(sys.argv) > 3:
    SETTINGS = [s for s in SETTINGS if any(k in s[0] for k in sys.argv[3].split(","))]
SWITCHES = ("SY_EPI_CYCLES", "SY_BN128_RULE", "SY_APPLY_HINTS", "SY_HEAD_PT", "SY_PAIR_APPLY", "SY_RAW_ARENA_MB", "SY_CONV_TEAM", "SY_CONV_PAIR", "SY_DBG_SKIP_APPLY", "SY_CONV_TILES", "SY_PDL", "SY_APPLY", "SY_APPLY_CAP", "SY_STAGE_TILES", "SY_APPLY_CARVEOUT", "SY_CONV_DEBUG", "SY_CONV_A")


def measure(label, env, fuse_mb, steps=20, warmup=4):
    for k in SWITCHES:
        os.environ.pop(k, None)
    os.environ.update(env)
    engine.FUSE_APPLY_MAX_BYTES = fuse_mb * 1e6
    mb = float(os.environ.get("SY_RAW_ARENA_MB", "40"))
    if mb != engine.RAW_ARENA_MB:
        engine._RAW_ARENAS.clear()          # the arena is allocated once at its configured size: re-create it
        engine._CAPTURE_STREAMS.clear()
    engine.RAW_ARENA_MB = mb
    with torch.no_grad():
        ops.LAUNCHES = 0
        out = model(x, (fut, cur))
        launches = ops.LAUNCHES
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            model(x, (fut, cur))
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=engine.graph_capture_stream(dev)):
            o = model(x, (fut, cur))
            loss = o["total_loss"]
        for _ in range(warmup):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    print(f"{label:30s} {ms:8.3f} ms/step {B / ms * 1e3:8.1f} pairs/s  loss {float(loss):.5f}  launches {launches}", flush=True)
    del g


print("raw arena MB", engine.RAW_ARENA_MB, flush=True)
for s in SETTINGS:
    try:
        measure(*s)
    except Exception as e:  # keep going: a failing switch must not hide the others
        print(f"{s[0]:30s} FAILED: {type(e).__name__}: {str(e)[:300]}", flush=True)
        torch.cuda.synchronize()

print("persisting-L2 window granted (bytes):", {k: v[1] for k, v in engine._RAW_ARENAS.items()})
