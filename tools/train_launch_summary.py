"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of tools/bench_train.py --eager: the kernels of the
LAST training step (between the last two sgd_ema_kernel launches), per kernel name.
    python tools/train_launch_summary.py gpurun_out/launches_train.csv"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [re.sub(r"\(.*", "", r["Kernel Name"]) for r in rows]
    idx = [i for i, n in enumerate(names) if "sgd_ema" in n]
    a, b = idx[-2] + 1, idx[-1] + 1
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r, n in zip(rows[a:b], names[a:b]):
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        tot[n] += v
        cnt[n] += 1
    T = sum(tot.values())
    print(f"step launches {b - a} sum {T / 1e3:.2f} ms (cold cache, serialised)")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:45]:
        print(f"{v / 1e3:8.3f} ms {100 * v / T:5.1f}% n={cnt[k]:4d} avg={v / cnt[k]:8.1f}us  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1])
