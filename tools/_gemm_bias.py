import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/computational-chemistry-ai_b200')
import ctypes as C, numpy as np, torch
from eims_b200 import _lib
from eims_b200._lib import check, ptr
DEV = "cuda:0"
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
M, N = 128, 256
for K in (32, 256, 1024, 2048):
    for mode in ("pos", "neg", "mixed"):
        g = torch.Generator(device="cpu").manual_seed(K)
        A = torch.randn(M, K, generator=g, dtype=torch.float64)
        B = torch.randn(K, N, generator=g, dtype=torch.float64)
        if mode != "mixed":
            A, B = A.abs(), B.abs()
        if mode == "neg":
            A = -A
        Ad = A.float().to(DEV).contiguous(); Bd = B.t().contiguous().float().to(DEV)
        ref = Ad.double() @ Bd.double().t()
        for backend in (0, 1):
            out = torch.zeros(M, N, device=DEV)
            check(_lib.load().eims_gemm(backend, ptr(Ad), K, 0, ptr(Bd), K, 0, ptr(out), N, M, N, K, None, None, None, None, 0, 0, st()))
            torch.cuda.synchronize()
            e = (out.double() - ref)
            scale = ref.abs().mean().item()
            print(f"K={K:5d} {mode:5s} backend={'tc' if backend==0 else 'simt'}: mean signed err/|ref| {e.mean().item()/scale:+.3e}  rms {e.pow(2).mean().sqrt().item()/scale:.3e}  max {e.abs().max().item()/ref.abs().max().item():.3e}")
