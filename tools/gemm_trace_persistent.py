#!/usr/bin/env python
"""Per-tile timeline of the persistent tcgen05 GEMM (diagnostic; -DEIMS_GEMM_TRACE build).
   python tools/gemm_trace_persistent.py M N K [a_mn b_mn accumulate]   (on a B200)"""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "computational-chemistry-ai_b200", "build.py"))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)


def main():
    M, N, K = (int(x) for x in sys.argv[1:4])
    a_mn, b_mn, acc = (int(x) for x in (sys.argv[4:7] + ["0", "1", "0"][len(sys.argv[4:7]):]))
    path = os.environ.get("EIMS_TRACE_LIB") or (b.build_variant("trace", ["-DEIMS_GEMM_TRACE"]) if "--no-build" not in sys.argv else b.OUT.replace(".so", "_trace.so"))
    lib = C.CDLL(path)
    dev = "cuda"
    A = torch.randn((K, M) if a_mn else (M, K), device=dev)
    B = torch.randn((K, N) if b_mn else (N, K), device=dev)
    Cm = torch.zeros(M, N, device=dev)
    vp = C.c_void_p
    lib.eims_gemm.argtypes = [C.c_int32, vp, C.c_int32, C.c_int32, vp, C.c_int32, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32,
                              C.c_int32, vp, vp, vp, vp, C.c_int32, C.c_int32, vp]
    st = vp(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        rc = lib.eims_gemm(0, vp(A.data_ptr()), A.shape[1], a_mn, vp(B.data_ptr()), B.shape[1], b_mn, vp(Cm.data_ptr()), N, M, N, K,
                           None, None, None, None, 0, acc, st)
        assert rc == 0, rc
        torch.cuda.synchronize()
    n = 8 * 128
    buf = (C.c_ulonglong * n)()
    assert lib.eims_debug_trace_read(buf, n) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(8, 128)
    names = ["mma: tile reached", "mma: accumulator free", "mma: first stage full", "mma: last stage full", "epi: waiting", "epi: accumulator full",
             "epi: tile stored"]
    for c in (0, 1):
        base = t[c, 0]
        print(f"--- CTA {c} (cycles since the grid-dependency wait)")
        for tile in range(8):
            row = t[c, 16 + tile * 8: 16 + tile * 8 + 7]
            if not row.any():
                break
            print(f" tile {tile}: " + "  ".join(f"{nm} {v - base}" for nm, v in zip(names, row) if v))


if __name__ == "__main__":
    main()
