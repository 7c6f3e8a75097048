"""Measure the FP64 dense-GEMM (cuBLAS DGEMM via torch.matmul) throughput on this GPU.

MEASURED_PEAKS.json (driver-written) holds only HBM GB/s and bf16 TF/s; the PLS Langevin step is an FP64
tensor-core (DMMA) path, so its roofline denominator is measured here the same way the driver measures bf16:
torch.matmul fp64 N^3, best of 10 (burst) and back to back for ~4 s (sustained), CUDA events.
Writes profiles/fp64_peak.json when --out is given.
"""
import argparse, json, time
import torch

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--out", type=str, default=None)
    a = ap.parse_args()
    n = a.n
    dev = torch.device("cuda:0")
    x = torch.randn(n, n, dtype=torch.float64, device=dev)
    y = torch.randn(n, n, dtype=torch.float64, device=dev)
    z = torch.empty(n, n, dtype=torch.float64, device=dev)
    for _ in range(3):
        torch.matmul(x, y, out=z)
    torch.cuda.synchronize()
    flops = 2.0 * n ** 3
    best = 1e30
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(x, y, out=z); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = flops / best * 1e-9
    # sustained
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = max(1, int(4000.0 / best))
    e0.record()
    for _ in range(reps):
        torch.matmul(x, y, out=z)
    e1.record(); e1.synchronize()
    sustained = flops * reps / e0.elapsed_time(e1) * 1e-9
    res = {"fp64_tflops": round(burst, 2), "fp64_tflops_sustained": round(sustained, 2), "n": n,
           "gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "how": f"torch.matmul fp64 {n}^3 (2*N^3): best of 10 (burst) and {reps} back to back (sustained), CUDA events",
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)

if __name__ == "__main__":
    main()
