#!/usr/bin/env python
"""Back-to-back timing of the tcgen05 GEMM through the C ABI (diagnostic).
   python tools/gemm_time.py M N K [a_mn b_mn accumulate]   (on a B200)"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "computational-chemistry-ai_b200"))
from eims_b200 import _lib  # noqa: E402
from eims_b200._lib import check, ptr  # noqa: E402


def main():
    M, N, K = (int(x) for x in sys.argv[1:4])
    a_mn, b_mn, acc = (int(x) for x in (sys.argv[4:7] + ["0", "1", "0"][len(sys.argv[4:7]):]))
    lib = _lib.load()
    dev = "cuda"
    nbuf = 8  # rotate over buffers so that A/C are not trivially L2-hot beyond what a step sees
    As = [torch.randn((K, M) if a_mn else (M, K), device=dev) for _ in range(nbuf)]
    B = torch.randn((K, N) if b_mn else (N, K), device=dev)
    Cs = [torch.zeros(M, N, device=dev) for _ in range(nbuf)]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run(i):
        A, Cm = As[i % nbuf], Cs[i % nbuf]
        check(lib.eims_gemm(0, ptr(A), A.shape[1], a_mn, ptr(B), B.shape[1], b_mn, ptr(Cm), N, M, N, K, None, None, None, None, 0, acc, st))

    for i in range(20):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record()
    for i in range(n):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} acc={acc}: {us:.2f} us/launch, {2.0 * M * N * K / us * 1e-6:.1f} TFLOP/s (1x flops)")


if __name__ == "__main__":
    main()
