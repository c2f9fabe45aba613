"""cuBLAS FP64 GEMM peak on this box (the FP64 roofline denominator; MEASURED_PEAKS.json has none)."""
import json, torch
assert torch.cuda.is_available()
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(3): torch.matmul(a, b)
    torch.cuda.synchronize(); best = 1e9
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_tflops"] = 2 * n**3 / best * 1e-9
# batched small: L=32, [20000x60]x[60x60]
a = torch.randn(32, 20000, 64, dtype=torch.float64, device="cuda"); b = torch.randn(32, 64, 64, dtype=torch.float64, device="cuda")
for _ in range(3): torch.matmul(a, b)
torch.cuda.synchronize(); best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
res["bmm_32x20000x64x64_tflops"] = 2 * 32 * 20000 * 64 * 64 / best * 1e-9
print(json.dumps(res))
