import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/computational-chemistry-ai_b200')
import ctypes as C, numpy as np, torch
from eims_b200 import _lib
from eims_b200._lib import check, ptr
DEV = "cuda:0"
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
shapes = [(24, 2048, 2048), (24, 1024, 2048), (24, 1000, 1024), (1024, 2048, 24), (2048, 2048, 24), (1000, 1024, 24),
          (24, 1024, 1000), (24, 2048, 1024), (24, 2048, 2048), (1560, 1024, 1024), (1024, 1024, 1560), (512, 2048, 2048)]
for (M, N, K) in shapes:
    for a_mn, b_mn in [(0, 0), (0, 1), (1, 0), (1, 1)]:
        for acc in (0, 1, 2):
            g = torch.Generator(device="cpu").manual_seed(M + N + K)
            A = torch.randn(M, K, generator=g, dtype=torch.float64)
            B = torch.randn(K, N, generator=g, dtype=torch.float64)
            Ad = (A.t().contiguous() if a_mn else A.contiguous()).float().to(DEV)
            Bd = (B.contiguous() if b_mn else B.t().contiguous()).float().to(DEV)
            out = torch.zeros(M, N, device=DEV) if acc != 0 else torch.full((M, N), 3.0, device=DEV)
            check(_lib.load().eims_gemm(0, ptr(Ad), Ad.stride(0), a_mn, ptr(Bd), Bd.stride(0), b_mn, ptr(out), N, M, N, K, None, None,
                                        None, None, 0, acc, st()))
            torch.cuda.synchronize()
            ref = (Ad.double().t() if a_mn else Ad.double()) @ (Bd.double() if b_mn else Bd.double().t())
            err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
            flag = "   <<<<<< BAD" if err > 2e-5 else ""
            if flag or (a_mn, b_mn, acc) == (0, 0, 0):
                print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} acc={acc}: err {err:.2e}{flag}")
