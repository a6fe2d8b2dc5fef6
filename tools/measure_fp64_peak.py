"""Measure the FP64 GEMM peak of this GPU (cuBLAS DGEMM 8192^3 through torch.matmul) -- the roofline
denominator for the Gram kernels, since MEASURED_PEAKS.json has no FP64 entry.
Same protocol as the driver's bf16 entry: best of 10 (burst) and back to back for 4 s (sustained).
Writes one JSON line to stdout."""
import json
import time

import torch

N = 8192
a = torch.randn(N, N, dtype=torch.float64, device="cuda")
b = torch.randn(N, N, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
flops = 2.0 * N ** 3
best = 1e30
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = flops / (best * 1e-3) / 1e12
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
reps = 0
e0.record()
while True:
    for _ in range(5):
        torch.matmul(a, b, out=c)
    reps += 5
    torch.cuda.synchronize()
    if time.time() - t0 > 4.0:
        break
e1.record()
torch.cuda.synchronize()
sustained = flops * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
print(json.dumps({"fp64_gemm_tflops": round(burst, 2), "fp64_gemm_tflops_sustained": round(sustained, 2),
                  "how": "torch.matmul float64 8192^3 (cuBLAS DGEMM), best of 10 and 4 s back to back",
                  "gpu": torch.cuda.get_device_name(0)}))
