"""cuBLAS DGEMM throughput on this box: the practical FP64 ceiling (development tool)."""
import torch, time
torch.backends.cuda.matmul.allow_tf32 = False
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ b
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"cuBLAS dgemm {n}^3: best {best:.3f} ms  {2*n**3/best*1e-9:.2f} TFLOP/s")
