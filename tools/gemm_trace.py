#!/usr/bin/env python
"""Pipeline timeline of the tcgen05 GEMM (diagnostic): builds the -DEIMS_GEMM_TRACE variant of the
library, runs one GEMM of the given shape through the C ABI and prints the SM-clock stamps of
the first CTAs.   python tools/gemm_trace.py M N K [a_mn b_mn accumulate]   (on a B200)"""
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
    path = b.build_variant("trace", ["-DEIMS_GEMM_TRACE"]) if "--no-build" not in sys.argv else b.OUT.replace(".so", "_trace.so")
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
    names = {0: "start", 1: "prologue done", 2: "pdl wait done", 3: "accum ready", 4: "tile staged", 5: "stats done", 6: "stores done", 7: "end"}
    for c in (0, 1, 5):
        base = t[c, 0]
        print(f"--- CTA {c} (cycles since kernel start of this CTA)")
        ev = [(t[c, k] - base, names[k]) for k in names if t[c, k]]
        for i in range(12):
            for off, nm in ((0, "g%d empty acquired kb%d"), (1, "g%d stored kb%d"), (2, "MMA%s full kb%d")):
                v = t[c, 8 + i * 4 + off]
                if v:
                    ev.append((v - base, nm % (("" if off == 2 else i % 2), i)))
        for v, nm in sorted(ev):
            print(f"{v:8d}  {nm}")


if __name__ == "__main__":
    main()
