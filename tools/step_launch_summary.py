"""Summarise an ncu launch list of bench.py --no-graph (metrics gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum): the kernels of the LAST forward+loss step (between the last two k_final launches) per kernel name with
their DRAM traffic; optionally writes that step's rows as a small CSV.
    python tools/step_launch_summary.py gpurun_out/launches_step.csv [trimmed.csv]"""
import collections
import csv
import re
import sys


def main(path, trimmed=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    by = collections.OrderedDict()
    for r in rows:
        by.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"])})[r["Metric Name"]] = (
            float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    ids = list(by)
    L = [by[i] for i in ids]
    kf = [i for i, l in enumerate(L) if "k_final" in l["name"]]
    a, b = kf[-2] + 1, kf[-1] + 1

    def us(v, u):
        return v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v

    def mb(v, u):
        return v / 1e6 if u == "byte" else v / 1e3 if u == "Kbyte" else v if u == "Mbyte" else v * 1e3

    tot = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0])
    for l in L[a:b]:
        t = tot[l["name"]]
        t[0] += us(*l["gpu__time_duration.sum"]); t[1] += 1
        t[2] += mb(*l["dram__bytes_read.sum"]); t[3] += mb(*l["dram__bytes_write.sum"])
    T = sum(t[0] for t in tot.values())
    print(f"# ncu launch list of ONE forward+loss step (StreamYOLO-l, 8 pairs, eager, cold caches, serialised): {b - a} launches, sum {T / 1e3:.3f} ms")
    print("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none ... bench.py --steps 1 "
          "--warmup 3 --no-graph --no-train --no-extras --no-cpu-baseline")
    print("#  ms/step  share    n   avg us  DRAM rd MB  DRAM wr MB  kernel   (conv_tc_kernel<BN, timeline, A mode 1 linear / 2 halo, pair>)")
    cr = cw = cn = ct = 0
    for k, t in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print(f"{t[0] / 1e3:10.3f} {100 * t[0] / T:5.1f}% {t[1]:4d} {t[0] / t[1]:8.1f} {t[2]:11.1f} {t[3]:11.1f}  {k[:70]}")
        if "conv_tc" in k:
            cr += t[2]; cw += t[3]; cn += t[1]; ct += t[0]
    print(f"# conv family: {cn} launches, {ct / 1e3:.3f} ms cold, DRAM read {cr:.0f} MB + written {cw:.0f} MB = {(cr + cw) / max(cn, 1):.1f} MB "
          "per launch; algorithmic FLOPs 3074.4 GFLOP per step")
    print(f"# whole step DRAM traffic: {sum(t[2] + t[3] for t in tot.values()) / 1e3:.2f} GB (algorithmic minimum 8 x 1.10 GB = 8.8 GB)")
    if trimmed:
        keep = set(ids[a:b])
        with open(trimmed, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=["ID", "Kernel Name", "Block Size", "Grid Size", "Metric Name", "Metric Unit", "Metric Value"])
            w.writeheader()
            for r in rows:
                if r["ID"] in keep:
                    w.writerow({k: (re.sub(r"\(CUtensorMap.*|\(const.*|\(sy::.*", "", r[k]) if k == "Kernel Name" else r[k]) for k in w.fieldnames})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
