#!/usr/bin/env python
"""Where a GCN-layer GEMM launch spends its time beyond the tile (diagnostic, on a B200): static vs device-side M
(dead CTAs of the capacity grid), short K (fill + epilogue only), with ReLU / row scale."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "computational-chemistry-ai_b200"))
from eims_b200 import _lib  # noqa: E402
from eims_b200._lib import check, ptr  # noqa: E402


def timeit(lib, M, N, K, a_mn, b_mn, acc, live=None, n=200):
    dev = "cuda"
    nbuf = 8
    As = [torch.randn((K, M) if a_mn else (M, K), device=dev) for _ in range(nbuf)]
    B = torch.randn((K, N) if b_mn else (N, K), device=dev)
    Cs = [torch.zeros(M, N, device=dev) for _ in range(nbuf)]
    md = torch.tensor([live], dtype=torch.int32, device=dev) if live is not None else None
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run(i):
        A, Cm = As[i % nbuf], Cs[i % nbuf]
        check(lib.eims_gemm(0, ptr(A), A.shape[1], a_mn, ptr(B), B.shape[1], b_mn, ptr(Cm), N, M, N, K,
                            ptr(md) if md is not None else None, None, None, None, 0, acc, st))
    for i in range(20):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    m = live if live is not None else M
    print(f"M={M} live={live} N={N} K={K} a_mn={a_mn} b_mn={b_mn} acc={acc}: {us:.2f} us/launch, {2.0 * m * N * K / us * 1e-6:.1f} TFLOP/s")


def main():
    lib = _lib.load()
    timeit(lib, 16900, 256, 256, 0, 1, 0)
    timeit(lib, 32768, 256, 256, 0, 1, 0, live=16900)
    timeit(lib, 16900, 256, 32, 0, 1, 0)
    timeit(lib, 16900, 256, 64, 0, 1, 0)
    timeit(lib, 16900, 256, 128, 0, 1, 0)
    timeit(lib, 16900, 256, 512, 0, 1, 0)
    timeit(lib, 18944, 256, 256, 0, 1, 0)   # 148 tiles
    timeit(lib, 256, 256, 16900, 1, 1, 1)   # wgrad split-K
    timeit(lib, 16900, 256, 256, 0, 0, 0)   # dgrad (B = W stored [N,K])


if __name__ == "__main__":
    main()
