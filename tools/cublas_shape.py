"""Measurement only: cuBLAS bf16 on the headline GEMM shape (1M x 768) . (768 x 512) -- the ceiling the
hand-written candidate pass is compared with (it never runs on the product path)."""
import json, sys, torch
dev = torch.device("cuda", 0)
out = {}
for (m, k, n) in [(1_000_000, 768, 512), (8192, 8192, 8192), (1_000_000, 768, 256)]:
    a = torch.randn((m, k), device=dev, dtype=torch.bfloat16)
    b = torch.randn((n, k), device=dev, dtype=torch.bfloat16)
    c = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
    for _ in range(5):
        torch.matmul(a, b.t(), out=c)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            torch.matmul(a, b.t(), out=c)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 5)
    out[f"{m}x{k}x{n}"] = {"ms": best, "tflops": 2.0 * m * k * n / best / 1e9}
    del a, b, c
print(json.dumps(out))
