#!/usr/bin/env python
"""Measures the TF32 denominator ONCE and writes profiles/tf32_peak.json (committed, then read by bench.py).

Method (same as the driver's MEASURED_PEAKS.json bf16 entry, with allow_tf32): torch.matmul fp32 8192^3 through cuBLAS TF32,
best of 10 single launches (burst) and back to back for 4 s (sustained); nvidia-smi SM clocks sampled under load."""
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = True
    n = 8192
    a, b = torch.randn(n, n, device=dev), torch.randn(n, n, device=dev)
    for _ in range(5):
        a @ b
    torch.cuda.synchronize()
    fl = 2.0 * n ** 3
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = max(best, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    clocks, stop = [], threading.Event()

    def sample():
        while not stop.is_set():
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                                 capture_output=True, text=True).stdout.strip()
            if out:
                clocks.append([float(x) for x in out.split(",")])
            stop.wait(0.2)

    th = threading.Thread(target=sample, daemon=True)
    th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps, t0 = 0, time.time()
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            a @ b
        reps += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    sustained = fl * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    sm = sorted(c[0] for c in clocks)
    out = {"tf32_tflops_burst": round(best, 1), "tf32_tflops_sustained": round(sustained, 1), "tf32_nominal_tflops": 1125.0,
           "how": "torch.matmul fp32 8192^3, allow_tf32 (cuBLAS): best of 10 (burst), back to back for 4 s (sustained)",
           "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "sm_mhz_median_under_load": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(c[1] for c in clocks) if clocks else None,
           "power_w_max": max(c[2] for c in clocks) if clocks else None, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "tf32_peak.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
