"""Does splitting a pinned H2D copy over several streams (copy engines) raise the rate on this box?"""
import torch, time
n = 29_119_744 // 8
h = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device="cuda")
for k in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(k)]
    best = 1e9
    for rep in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step = (n + k - 1) // k
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{k} stream(s): {n * 8 / best / 1e9:6.1f} GB/s  ({best * 1e3:.3f} ms)")
