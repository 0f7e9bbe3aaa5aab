"""How fast are the library FP64 building blocks on this GPU?  (sizing the reduced solve, K5)"""
import torch, time
torch.backends.cuda.preferred_linalg_library("cusolver")
dev = "cuda"
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
for n in (3009, 6000, 12060, 30009):
    A = torch.randn(n, n // 4 + 8, dtype=torch.float64, device=dev)
    S = A @ A.T + n * torch.eye(n, dtype=torch.float64, device=dev)
    ms = t(lambda: torch.linalg.cholesky_ex(S))
    B = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms_gemm = t(lambda: torch.mm(B, B), reps=2)
    print(f"n={n:6d}  potrf {ms:9.2f} ms = {n**3/3/ms/1e9:6.2f} TFLOP/s   dgemm {ms_gemm:9.2f} ms = {2*n**3/ms_gemm/1e9:6.2f} TFLOP/s", flush=True)
    del A, S, B
