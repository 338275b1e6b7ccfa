"""Run the GPU checks in isolated subprocesses (a device trap in one cannot mask the others) and
collect logs under gpurun_out/diag/.   usage: python tools/gpu_diag.py [groups...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "diag")
os.makedirs(OUT, exist_ok=True)
PY = sys.executable


def run(name, cmd, timeout=600):
    t0 = time.time()
    log = os.path.join(OUT, name + ".log")
    with open(log, "w") as fh:
        try:
            r = subprocess.run(cmd, cwd=ROOT, stdout=fh, stderr=subprocess.STDOUT, timeout=timeout)
            rc = r.returncode
        except subprocess.TimeoutExpired:
            rc = -999
    tail = open(log).read().strip().splitlines()[-3:]
    print(f"[{name}] rc={rc} {time.time() - t0:.1f}s :: " + " | ".join(tail), flush=True)
    return rc


GROUPS = {
    "ops_simt": lambda: run("ops_simt", [PY, "-m", "pytest", "tests/test_gpu_ops.py", "-q", "-k", "not tc", "-x"]),
    "tc_probe": lambda: [run(f"tc_probe_{i}", [PY, "tools/tc_probe.py", str(i)], 180) for i in range(12)],
    "ops_tc": lambda: run("ops_tc", [PY, "-m", "pytest", "tests/test_gpu_ops.py", "-q", "-k", "tc"]),
    "model_simt": lambda: run("model_simt", [PY, "-m", "pytest", "tests/test_gpu_model.py", "-q", "-k", "simt or loss_kernels"]),
    "model_tc": lambda: run("model_tc", [PY, "-m", "pytest", "tests/test_gpu_model.py", "-q", "-k", "not simt"]),
}

if __name__ == "__main__":
    sel = sys.argv[1:] or list(GROUPS)
    for g in sel:
        GROUPS[g]()
