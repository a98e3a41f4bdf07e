"""Dev: extra peaks for DESIGN.md (SURVEY 8d asks for the TF32 GEMM peak next to the bf16 one)."""
import torch, time
dev = "cuda"
def bench(fn, flops, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return flops / best / 1e9
n = 8192
x = torch.randn(n, n, device=dev); y = torch.randn(n, n, device=dev)
torch.backends.cuda.matmul.allow_tf32 = True
print("TF32 matmul 8192^3: %.0f TFLOP/s" % bench(lambda: x @ y, 2 * n ** 3))
torch.backends.cuda.matmul.allow_tf32 = False
print("FP32 (no TF32) matmul 8192^3: %.1f TFLOP/s" % bench(lambda: x @ y, 2 * n ** 3, 3))
xb = x.bfloat16(); yb = y.bfloat16()
print("BF16 matmul 8192^3: %.0f TFLOP/s" % bench(lambda: xb @ yb, 2 * n ** 3))
xd = torch.randn(4096, 4096, device=dev, dtype=torch.float64); yd = torch.randn(4096, 4096, device=dev, dtype=torch.float64)
print("FP64 matmul 4096^3: %.1f TFLOP/s" % bench(lambda: xd @ yd, 2 * 4096 ** 3, 3))
a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev); b = torch.empty_like(a)
print("copy 2 GiB (read+write): %.0f GB/s" % (1e3 * bench(lambda: b.copy_(a), 2 * a.numel() * 2)))
