# cuBLAS DGEMM 8192^3 via torch (library FP64 GEMM rate, for the FP64 roofline denominator)
import json, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
torch.matmul(a, b); torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"probe": "cublas_dgemm_8192", "tflops": 2 * n**3 / best * 1e-9, "ms": best}))
# skinny shape of the hot path: (1M x 64) @ (64 x 64)
y = torch.randn(1_000_000, 64, dtype=torch.float64, device="cuda"); w = torch.randn(64, 64, dtype=torch.float64, device="cuda")
torch.matmul(y, w); torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(y, w); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"probe": "cublas_dgemm_1Mx64x64", "tflops": 2 * 1e6 * 64 * 64 / best * 1e-9, "ms": best, "gbs": 2 * 512e6 / best * 1e-6}))
