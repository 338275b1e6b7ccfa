"""Top stalled SASS instructions of an ncu report (source page).  usage: ncu_hot.py report.ncu-rep [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(lines[start:]))
tot = sum(float(r["# Samples"] or 0) for r in rows)
print("total samples", tot, "instructions", len(rows))
for i, r in enumerate(rows):
    r["_i"] = i
top = sorted(rows, key=lambda r: -float(r["# Samples"] or 0))[:n]
for r in sorted(top, key=lambda r: r["_i"]):
    print(f'{r["_i"]:5d} {100*float(r["# Samples"] or 0)/tot:5.1f}%  {r["Source"][:110]}')
