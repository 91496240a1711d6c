"""The planes GEMM (csrc/gemm_tma.cu: tf32 hi / lo operand planes, cp.async.bulk.tensor, CTA pairs, persistent) through
the C ABI, against fp64 matmuls: the three GraphConv products (GCN:316,359 forward; its data and weight gradients)
alone and as the grouped data + weight gradient launch, at ragged live sizes with canaries behind the live range."""
import ctypes as C

import numpy as np
import pytest
import torch

from eims_b200 import _lib
from eims_b200._lib import GemmProblem, check, ptr

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-5   # 3xTF32 with one accumulator per tile: ~2e-6 measured at K = 256 .. 17 k


def problem(A, a_mn, B, b_mn, Cm, M, N, K, m_dev=None, k_dev=None, rs=None, bias=None, relu=0, acc=0):
    return GemmProblem(A.data_ptr(), A.shape[1], a_mn, B.data_ptr(), B.shape[1], b_mn, Cm.data_ptr(), Cm.shape[1], M, N, K,
                       m_dev.data_ptr() if m_dev is not None else None, k_dev.data_ptr() if k_dev is not None else None,
                       rs.data_ptr() if rs is not None else None, bias.data_ptr() if bias is not None else None, relu, acc)


def run(p0, p1=None):
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    b1 = C.byref(p1) if p1 is not None else None
    need = lib.eims_gemm_planes_scratch_bytes(C.byref(p0), b1)
    scratch = torch.empty(need, dtype=torch.uint8, device=DEV)
    rc = lib.eims_gemm_planes(C.byref(p0), b1, ptr(scratch), scratch.numel(), st)
    torch.cuda.synchronize()
    return rc


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("atoms,cap,H", [(300, 512, 256), (16900, 32768, 256), (1000, 1024, 512), (129, 4096, 256), (4096, 4096, 256),
                                         (3000, 3072, 1024)])   # hidden 1024: K = 1024, main + correction accumulators
def test_graphconv_products_vs_fp64(atoms, cap, H):
    g = torch.Generator(device=DEV).manual_seed(atoms + H)
    X = torch.randn(cap, H, device=DEV, generator=g)
    Q = torch.randn(cap, H, device=DEV, generator=g) * 0.01
    X[atoms:] = 0   # what the producers guarantee: zeros behind the live rows (up to the next multiple of 32)
    Q[atoms:] = 0
    W = torch.randn(H, H, device=DEV, generator=g) / H ** 0.5
    rs = torch.rand(cap, device=DEV, generator=g) + 0.5
    bias = torch.randn(H, device=DEV, generator=g)
    nd = torch.tensor([atoms], dtype=torch.int32, device=DEV)
    canary = 777.0
    # forward: relu((a W) * c + b), live rows from the device
    Z = torch.full((cap, H), canary, device=DEV)
    check(run(problem(X, 0, W, 1, Z, cap, H, H, m_dev=nd, rs=rs, bias=bias, relu=1)))
    ref = torch.relu((X[:atoms].double() @ W.double()) * rs[:atoms, None].double() + bias.double())
    assert rel(Z[:atoms], ref) < TOL
    assert bool((Z[atoms:] == canary).all())
    # data gradient q W^T and weight gradient a^T q, separately and as one grouped launch
    refd = Q[:atoms].double() @ W.double().T
    refw = X[:atoms].double().T @ Q[:atoms].double()
    for grouped in (False, True):
        DA = torch.full((cap, H), canary, device=DEV)
        DW = torch.full((H, H), 0.25, device=DEV)   # the weight gradient is ADDED into C
        pd = problem(Q, 0, W, 0, DA, cap, H, H, m_dev=nd)
        pw = problem(X, 1, Q, 1, DW, H, H, cap, k_dev=nd, acc=1)
        if grouped:
            check(run(pd, pw))
        else:
            check(run(pd))
            check(run(pw))
        assert rel(DA[:atoms], refd) < TOL, grouped
        assert bool((DA[atoms:] == canary).all())
        assert rel(DW - 0.25, refw) < TOL, grouped


def test_planes_match_in_kernel_split():
    """Same split, same products: the planes kernel and the in-kernel-split kernel (eims_gemm) agree to accumulation
    order (one accumulator per tile against main + correction accumulators)."""
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device=DEV).manual_seed(3)
    M, H = 2000, 256
    A = torch.randn(M, H, device=DEV, generator=g)
    W = torch.randn(H, H, device=DEV, generator=g)
    C1 = torch.zeros(M, H, device=DEV)
    C2 = torch.zeros(M, H, device=DEV)
    check(run(problem(A, 0, W, 1, C1, M, H, H)))
    check(lib.eims_gemm(0, ptr(A), H, 0, ptr(W), H, 1, ptr(C2), H, M, H, H, None, None, None, None, 0, 0, st))
    torch.cuda.synchronize()
    ref = A.double() @ W.double()
    assert rel(C1, ref) < TOL and rel(C2, ref) < TOL
    assert float((C1 - C2).abs().max()) < 1e-4 * float(ref.abs().max())


def test_unsupported_shapes_are_refused():
    A = torch.zeros(256, 256, device=DEV)
    W = torch.zeros(256, 128, device=DEV)
    Cm = torch.zeros(256, 128, device=DEV)
    assert run(problem(A, 0, W, 1, Cm, 256, 128, 256)) == _lib.ERR_ARG      # N must be a multiple of 256
    A = torch.zeros(256, 4096, device=DEV)
    W = torch.zeros(4096, 256, device=DEV)
    Cm = torch.zeros(256, 256, device=DEV)
    assert run(problem(A, 0, W, 1, Cm, 256, 256, 4096)) == _lib.ERR_ARG     # K of a store problem <= 2048
    p0 = problem(A, 0, W, 1, Cm, 256, 256, 256)
    assert run(p0, p0) != 0                                                  # two store problems in one launch
