"""Summarise an ncu --metrics gpu__time_duration.sum CSV: per-kernel totals and shares."""
import collections
import csv
import re
import sys


def main(path, steps=1.0):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"{'ms/step':>9s} {'share':>6s} {'n/step':>7s} {'avg us':>8s}  kernel")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:30]:
        print(f"{v / 1e3 / steps:9.3f} {100 * v / T:5.1f}% {cnt[k] / steps:7.1f} {v / cnt[k]:8.1f}  {k[:100]}")
    print(f"total {T / 1e3 / steps:.3f} ms/step over {sum(cnt.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
