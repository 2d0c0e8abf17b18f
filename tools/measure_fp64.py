"""One-off measurement of the fp64 roofline denominators on the GPU box:
cuBLAS DGEMM (torch.matmul fp64) at square and at the config-2 skinny shapes, the DFMA / DMMA microbenchmarks,
host core count.  Writes gpurun_out/fp64_peaks.json; the numbers are copied into profiles/ by hand."""
import json
import os
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent


def time_mm(a, b, reps=8):
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    out = {"gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count()}
    dev = "cuda"
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        ms = time_mm(a, b)
        out[f"cublas_dgemm_{n}_tflops"] = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
        del a, b
    A = torch.randn(10000, 100, dtype=torch.float64, device=dev)
    X = torch.randn(100, 4096, dtype=torch.float64, device=dev)
    W = torch.randn(10000, 4096, dtype=torch.float64, device=dev)
    t_fwd = time_mm(A, X)
    t_adj = time_mm(A.t(), W)
    out["c2_cublas_fwd_ms"] = t_fwd
    out["c2_cublas_adj_ms"] = t_adj
    out["c2_cublas_fwd_tflops"] = 2.0 * 10000 * 100 * 4096 / (t_fwd * 1e-3) / 1e12
    out["c2_cublas_adj_tflops"] = 2.0 * 10000 * 100 * 4096 / (t_adj * 1e-3) / 1e12
    out["c2_unfused_cublas_evals_per_s"] = 4096 / ((t_fwd + t_adj) * 1e-3)
    exe = ROOT / "tools" / "fp64_peaks"
    if exe.exists():
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        try:
            out["micro"] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:  # noqa: BLE001
            out["micro_error"] = f"{e}: {r.stdout[-300:]} {r.stderr[-300:]}"
    smi = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit", "--format=csv,noheader"],
                         capture_output=True, text=True)
    out["nvidia_smi"] = smi.stdout.strip()
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "fp64_peaks.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
