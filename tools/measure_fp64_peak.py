"""Measures the FP64 GEMM peak of this GPU with cuBLAS DGEMM (torch.matmul, float64,
8192^3): best of 10 (burst) and back-to-back for ~3 s (sustained).  This is the
denominator for the 128x128 Double `.*.` roofline ("builder measures").
Writes gpurun_out/fp64_peak.json; copy it to profiles/r01_fp64_peak.json."""
import json
import time

import torch

n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
flops = 2.0 * n ** 3
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = max(5, int(3000.0 / best))
e0.record()
for _ in range(iters):
    torch.matmul(a, b, out=c)
e1.record()
torch.cuda.synchronize()
sust = e0.elapsed_time(e1) / iters
out = {"fp64_tflops": flops / best / 1e9, "fp64_tflops_sustained": flops / sust / 1e9, "n": n, "iters": iters,
       "how": "torch.matmul float64 8192^3 (cuBLAS DGEMM), CUDA events, best of 10 / back-to-back",
       "gpu": torch.cuda.get_device_name(0), "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
print(json.dumps(out))
with open("gpurun_out/fp64_peak.json", "w") as f:
    json.dump(out, f, indent=1)
