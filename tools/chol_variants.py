"""cuSOLVER potrf / potrs variants at the reduced-system sizes (which fill mode, which triangular solve)."""
import torch
dev = "cuda"
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
for lib in ("cusolver", "magma"):
    try:
        torch.backends.cuda.preferred_linalg_library(lib)
    except Exception as e:
        print(lib, "unavailable", e); continue
    for n in (3009, 12060):
        A = torch.randn(n, n // 4 + 8, dtype=torch.float64, device=dev)
        S = A @ A.T + n * torch.eye(n, dtype=torch.float64, device=dev)
        b = torch.randn(n, 1, dtype=torch.float64, device=dev)
        lo = t(lambda: torch.linalg.cholesky_ex(S, upper=False))
        up = t(lambda: torch.linalg.cholesky_ex(S, upper=True))
        L = torch.linalg.cholesky(S)
        ps = t(lambda: torch.cholesky_solve(b, L))
        st = t(lambda: torch.linalg.solve_triangular(L.mT, torch.linalg.solve_triangular(L, b, upper=False), upper=True))
        Sc = S.clone()
        print(f"{lib:9s} n={n:6d} potrf lower {lo:7.3f} ms  upper {up:7.3f} ms  potrs {ps:7.3f} ms  2x trsm {st:7.3f} ms", flush=True)
