"""Dense Cholesky of the reduced solve: csrc/dense.cu (own) vs cusolverDnDpotrf on this GPU, one JSON line per size.
usage: chol_bench.py [n ...]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from robot_camera_calibration_b200 import _lib as L

lib = L.load()
sizes = [int(x) for x in sys.argv[1:]] or [3009, 6030, 12060, 30009]
for n in sizes:
    A = torch.randn(n, n // 4 + 8, dtype=torch.float64, device="cuda")
    S = A @ A.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
    del A
    ld = n + 2 + (n & 1)
    out = {"n": n}
    for name, cus in (("cusolver", 1), ("own", 0)):
        best, info = 1e30, C.c_int32()
        for rep in range(3):
            buf = torch.zeros((n + 1, ld), dtype=torch.float64, device="cuda")
            buf[:n, :n] = S
            ms = C.c_double()
            rc = lib.rcc_dense_potrf(0, C.c_void_p(buf.data_ptr()), n, ld, 0 if cus else 1, cus, C.byref(info), C.byref(ms))
            assert rc == 0 and info.value == 0, (name, n, rc, info.value, rep)
            best = min(best, ms.value)
            if rep == 0:
                Lf = torch.triu(buf[:n, :n]).T
                if name == "cusolver":
                    Lref = Lf.clone()
                else:
                    out["rel_diff_vs_cusolver"] = (torch.linalg.norm(Lf - Lref) / torch.linalg.norm(Lref)).item()
                del Lf
            del buf
        out[name + "_ms"] = best
        out[name + "_tflops"] = n ** 3 / 3 / best / 1e9
        if name == "own":
            # back-substitution on the factor just computed: hand-written kernel vs cublasDtrsv
            buf = torch.zeros((n + 1, ld), dtype=torch.float64, device="cuda")
            buf[:n, :n] = S
            lib.rcc_dense_potrf(0, C.c_void_p(buf.data_ptr()), n, ld, 1, 0, C.byref(info), C.byref(C.c_double()))
            r = torch.randn(n, dtype=torch.float64, device="cuda")
            for tname, cub in (("trsv_own", 0), ("trsv_cublas", 1)):
                tb = 1e30
                for rep in range(4):
                    x = r.clone()
                    ms = C.c_double()
                    assert lib.rcc_dense_trsv(0, C.c_void_p(buf.data_ptr()), n, ld, C.c_void_p(x.data_ptr()), cub, C.byref(ms)) == 0
                    tb = min(tb, ms.value)
                out[tname + "_ms"] = tb
            del buf
    print(json.dumps(out), flush=True)
    del S
    if n >= 20000:
        torch.cuda.empty_cache()
