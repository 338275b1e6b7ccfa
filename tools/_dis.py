import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
os.environ["SY_CONV_A"] = "off"
import conv_dissect as cd
for shape in [(8, 256, 256, 75, 120, 3, 1), (16, 256, 256, 38, 60, 3, 1), (16, 128, 128, 75, 120, 3, 1)]:
    for tiles in ("patch", "linear"):
        os.environ["SY_CONV_TILES"] = tiles
        t = [cd.run(*shape, f) for f in (0, 1, 2, 3)]
        print(f"{str(shape):32s} {tiles:6s} full {t[0]:6.1f} | no-MMA {t[1]:6.1f} | no-TMA {t[2]:6.1f} | neither {t[3]:6.1f}", flush=True)
