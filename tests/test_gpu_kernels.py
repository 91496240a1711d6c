"""Per-kernel parity: each CUDA kernel, called through the C ABI, against the CPU oracle
(or torch fp64 for the GEMMs) on the same seeded inputs.  Needs a B200 (`-m gpu`)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from eims_b200 import _lib
from eims_b200._lib import check, ptr
from eims_b200.engine import DeviceDataset, DevicePeaks, FlatParams, ModelDims, Plan, make_step, topk_peaks
from eims_b200.synth import MolTable, dense_spectra, peaks_as_lists, synth_molecules, synth_peaks
from oracle import gcn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
_BK = {"tcgen05": _lib.GEMM_TCGEN05, "simt": _lib.GEMM_FP32_SIMT}
BACKENDS = [_BK[b] for b in os.environ.get("EIMS_TEST_BACKENDS", "tcgen05,simt").split(",")]


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dev(a, dt=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dt is not None:
        t = t.to(dt)
    return t.to(DEV)


def make_dims(B, N, E, zero=0):
    return dev(np.array([B, N, E, zero, 0, 0, 0, 0], np.int32))


def graph_arrays(table, ids):
    mols = [table.mol(int(i)) for i in ids]
    b = O.batch_graphs(mols)
    rowptr, col, deg = O.csr_by_dst(b["src"], b["dst"], b["num_nodes"])
    return b, rowptr, col, deg, O.degree_norm(deg)


# ------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("n_mols,max_atoms,shuffle", [(1, 9, False), (64, 64, True), (700, 40, True), (33, 128, True), (1500, 24, True)])
def test_csr_build_bit_exact(n_mols, max_atoms, shuffle):
    table = synth_molecules(max(n_mols, 8) * 2, max_atoms=max_atoms, seed=5)
    rng = np.random.default_rng(0)
    ids = rng.permutation(table.num_mols)[:n_mols] if shuffle else np.arange(n_mols)
    b, rowptr, col, deg, norm = graph_arrays(table, ids)
    N, E = b["num_nodes"], len(b["src"])
    d = ModelDims(hidden_dim=64, max_mz=100)
    plan = Plan(d, n_mols, N + 7, E + 5, DEV)
    ds = DeviceDataset(table, None, DEV)
    plan.batch_build(ds, dev(ids, torch.int32))
    assert plan.check() == (N, E)
    i32 = torch.int32
    assert np.array_equal(plan.buffer("src", i32)[:E].cpu().numpy(), b["src"])
    assert np.array_equal(plan.buffer("dst", i32)[:E].cpu().numpy(), b["dst"])
    assert np.array_equal(plan.buffer("rowptr", i32)[:N + 1].cpu().numpy(), rowptr)
    assert np.array_equal(plan.buffer("col", i32)[:E].cpu().numpy(), col)
    assert np.array_equal(plan.buffer("gptr", i32)[:n_mols + 1].cpu().numpy(), O.graph_ptr(b["batch_num_nodes"]))
    assert np.array_equal(np.diff(plan.buffer("eptr", i32)[:n_mols + 1].cpu().numpy()), b["batch_num_edges"])
    assert np.array_equal(plan.buffer("gid", i32)[:N].cpu().numpy(), np.repeat(np.arange(n_mols), b["batch_num_nodes"]))
    got_norm = plan.buffer("norm")[:N].cpu().numpy()
    assert np.array_equal(got_norm.view(np.int32), norm.view(np.int32))  # raw bits
    assert np.array_equal(plan.buffer("x")[:N * 6].cpu().numpy().view(np.int32), b["feat"].reshape(-1).view(np.int32))


def test_csr_degree_norm_all_degrees():
    """star graphs with centre degree 1..64: fl(1/fl(sqrt(d))) must equal torch.pow(d,-0.5) bitwise."""
    feats, bb, be, nptr, bptr = [], [], [], [0], [0]
    for dgr in range(1, 65):
        feats.append(np.ones((dgr + 1, 6), np.float32))
        bb += [0] * dgr
        be += list(range(1, dgr + 1))
        nptr.append(nptr[-1] + dgr + 1)
        bptr.append(bptr[-1] + dgr)
    table = MolTable(np.array(nptr, np.int64), np.array(bptr, np.int64), np.concatenate(feats), np.array(bb, np.int32), np.array(be, np.int32))
    b, rowptr, col, deg, norm = graph_arrays(table, np.arange(64))
    plan = Plan(ModelDims(hidden_dim=64, max_mz=100), 64, nptr[-1], 2 * bptr[-1], DEV)
    plan.batch_build(DeviceDataset(table, None, DEV), None, 64)
    plan.check()
    got = plan.buffer("norm")[:nptr[-1]].cpu().numpy()
    assert np.array_equal(got.view(np.int32), norm.view(np.int32))
    assert np.array_equal(plan.buffer("col", torch.int32)[:2 * bptr[-1]].cpu().numpy(), col)


def test_csr_errors():
    d = ModelDims(hidden_dim=64, max_mz=100)
    # an atom without bonds: DGL GraphConv raises -> ZeroInDegreeError
    table = MolTable(np.array([0, 1, 3], np.int64), np.array([0, 0, 1], np.int64), np.ones((3, 6), np.float32),
                     np.array([0], np.int32), np.array([1], np.int32))
    plan = Plan(d, 4, 16, 16, DEV)
    plan.batch_build(DeviceDataset(table, None, DEV), None, 2)
    with pytest.raises(_lib.ZeroInDegreeError):
        plan.check()
    # capacity overflow is reported, nothing is written out of bounds
    big = synth_molecules(8, max_atoms=64, seed=1)
    plan.batch_build(DeviceDataset(big, None, DEV), None, 4)
    with pytest.raises(_lib.EimsError) as e:
        plan.check()
    assert e.value.code == _lib.ERR_CAPACITY
    with pytest.raises(_lib.EimsError):
        plan.batch_build(DeviceDataset(big, None, DEV), None, 5)  # > max_graphs
    # empty batch
    plan.batch_build(DeviceDataset(big, None, DEV), None, 0)
    assert plan.check() == (0, 0)


# ------------------------------------------------------------------------------- GEMM
def gemm(backend, A, a_mn, B, b_mn, M, N, K, row_scale=None, bias=None, relu=0, C0=None, m_dev=None, k_dev=None, Mcap=None, Kcap=None):
    out = torch.zeros(Mcap or M, N, device=DEV) if C0 is None else C0
    check(_lib.load().eims_gemm(backend, ptr(A), A.stride(0), a_mn, ptr(B), B.stride(0), b_mn, ptr(out), N,
                                Mcap or M, N, Kcap or K, ptr(m_dev), ptr(k_dev), ptr(row_scale), ptr(bias), relu,
                                int(C0 is not None), stream()))
    return out


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (300, 256, 256), (77, 1000, 516), (512, 64, 1000)])
def test_gemm_layouts(backend, a_mn, b_mn, M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g, dtype=torch.float64)
    B = torch.randn(K, N, generator=g, dtype=torch.float64)
    ref = A @ B
    Ad = (A.t().contiguous() if a_mn else A.contiguous()).float().to(DEV)
    Bd = (B.contiguous() if b_mn else B.t().contiguous()).float().to(DEV)
    out = gemm(backend, Ad, a_mn, Bd, b_mn, M, N, K)
    torch.cuda.synchronize()
    ref32 = (Ad.double().t() if a_mn else Ad.double()) @ (Bd.double() if b_mn else Bd.double().t())
    err = (out.double() - ref32).abs().max().item() / ref32.abs().max().item()
    # the tensor core truncates when it aligns partial sums, so 3xTF32 lands at ~5e-6 for K ~ 1000
    # (single-pass TF32 would be ~2e-4); the CUDA-core path is plain fp32
    assert err < (2e-5 if backend == _lib.GEMM_TCGEN05 else 2e-6), err


@pytest.mark.parametrize("backend", BACKENDS)
def test_gemm_epilogue_dynamic_and_splitk(backend):
    g = torch.Generator(device="cpu").manual_seed(1)
    Mcap, M, N, K = 640, 517, 256, 256
    A = torch.randn(Mcap, K, generator=g).to(DEV)
    W = torch.randn(K, N, generator=g).to(DEV)
    rs = torch.rand(Mcap, generator=g).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    m_dev = dev(np.array([M], np.int32))
    out = torch.full((Mcap, N), 7.0, device=DEV)
    check(_lib.load().eims_gemm(backend, ptr(A), K, 0, ptr(W), N, 1, ptr(out), N, Mcap, N, K, ptr(m_dev), None,
                                ptr(rs), ptr(bias), 1, 0, stream()))
    ref = torch.relu((A[:M].double() @ W.double()) * rs[:M, None].double() + bias.double())
    assert (out[:M].double() - ref).abs().max().item() / ref.abs().max().item() < 1e-5
    assert bool((out[M:] == 7.0).all())  # rows past the live size are untouched
    # split-K accumulate with K on the device: dW[F,H] += A^T Q over K = M rows
    Q = torch.randn(Mcap, N, generator=g).to(DEV)
    dW = torch.ones(K, N, device=DEV)
    check(_lib.load().eims_gemm(backend, ptr(A), K, 1, ptr(Q), N, 1, ptr(dW), N, K, N, Mcap, None, ptr(m_dev),
                                None, None, 0, 1, stream()))
    ref = 1.0 + A[:M].double().t() @ Q[:M].double()
    assert (dW.double() - ref).abs().max().item() / ref.abs().max().item() < 1e-5


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_persistent_kernel(a_mn, b_mn):
    """Launches with >= 4 tiles of 128x256 per SM take the persistent kernel (one CTA per SM walking its
    tiles, double-buffered TMEM accumulator): all four storage orders, a live size on the device that leaves
    a ragged last tile and hundreds of dead tiles, row scale + bias + ReLU, a K that is not a multiple of 32,
    and the split-K accumulate form."""
    g = torch.Generator(device="cpu").manual_seed(5 + 2 * a_mn + b_mn)
    Mcap, M, N, K = 80000, 40777, 512, 200           # 625 x 2 tiles of capacity, 319 x 2 live
    A = torch.randn(Mcap, K, generator=g)
    B = torch.randn(K, N, generator=g)
    Ad = (A.t().contiguous() if a_mn else A.contiguous()).to(DEV)
    Bd = (B.contiguous() if b_mn else B.t().contiguous()).to(DEV)
    rs = torch.rand(Mcap, generator=g).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    m_dev = dev(np.array([M], np.int32))
    out = torch.full((Mcap, N), 7.0, device=DEV)
    check(_lib.load().eims_gemm(_lib.GEMM_TCGEN05, ptr(Ad), Ad.stride(0), a_mn, ptr(Bd), Bd.stride(0), b_mn, ptr(out), N,
                                Mcap, N, K, ptr(m_dev), None, ptr(rs), ptr(bias), 1, 0, stream()))
    ref = torch.relu((A[:M].double() @ B.double()) * rs[:M, None].cpu().double() + bias.cpu().double())
    assert (out[:M].cpu().double() - ref).abs().max().item() / ref.abs().max().item() < 1e-5
    assert bool((out[M:] == 7.0).all())               # dead tiles and the rows past M of the ragged tile are untouched
    # accumulate into C (every CTA adds its tiles with red.global.add.v4.f32)
    out2 = torch.ones(Mcap, N, device=DEV)
    check(_lib.load().eims_gemm(_lib.GEMM_TCGEN05, ptr(Ad), Ad.stride(0), a_mn, ptr(Bd), Bd.stride(0), b_mn, ptr(out2), N,
                                Mcap, N, K, ptr(m_dev), None, None, None, 0, 1, stream()))
    ref2 = 1.0 + A[:M].double() @ B.double()
    assert (out2[:M].cpu().double() - ref2).abs().max().item() / ref2.abs().max().item() < 1e-5
    assert bool((out2[M:] == 1.0).all())


# ------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("H", [64, 256, 1024])
def test_spmm_forward_backward(H):
    table = synth_molecules(40, max_atoms=64, seed=3)
    b, rowptr, col, deg, norm = graph_arrays(table, np.arange(40))
    N, E = b["num_nodes"], len(b["src"])
    rng = np.random.default_rng(1)
    h = rng.standard_normal((N, H)).astype(np.float32)
    scale = rng.standard_normal(H).astype(np.float32)
    shift = rng.standard_normal(H).astype(np.float32)
    dims = make_dims(40, N, E)
    args = (dims, dev(rowptr, torch.int32), dev(col, torch.int32), dev(norm))
    lib = _lib.load()
    out = torch.empty(N, H, device=DEV)
    hd = dev(h)
    # forward without BN: bit-exact with index_add_ in edge order
    check(lib.eims_spmm_norm(*map(ptr, args), ptr(hd), H, None, None, 0.0, 0, 0, 0, 0, ptr(out), N, stream()))
    g = O.Graph(b["src"], b["dst"], b["batch_num_nodes"])
    s = torch.from_numpy(h) * torch.from_numpy(norm)[:, None]
    ref = torch.zeros(N, H).index_add_(0, g.dst, s[g.src])
    assert np.array_equal(out.cpu().numpy().view(np.int32), ref.numpy().view(np.int32))
    # forward with BN apply
    scale_d, shift_d = dev(scale), dev(shift)
    check(lib.eims_spmm_norm(*map(ptr, args), ptr(hd), H, ptr(scale_d), ptr(shift_d), 0.0, 0, 0, 0, 0, ptr(out), N, stream()))
    hb = torch.from_numpy(h).double() * torch.from_numpy(scale).double() + torch.from_numpy(shift).double()
    ref = torch.zeros(N, H, dtype=torch.float64).index_add_(0, g.dst, (hb * torch.from_numpy(norm).double()[:, None])[g.src])
    assert (out.cpu().double() - ref).abs().max().item() < 2e-5 * ref.abs().max().item()
    # backward form: (A da) * c
    check(lib.eims_spmm_norm(*map(ptr, args), ptr(hd), H, None, None, 0.0, 0, 0, 0, 1, ptr(out), N, stream()))
    ref = torch.zeros(N, H).index_add_(0, g.dst, torch.from_numpy(h)[g.src]) * torch.from_numpy(norm)[:, None]
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-6, atol=1e-6)
    # dropout: forward mask on the gathered rows == materialised mask; backward uses the same mask
    p, seed, step, site = 0.2, 1234567, 3, 1
    mask = torch.empty(N, H, device=DEV)
    check(lib.eims_dropout_mask(p, seed, step, site, N, H, ptr(mask), stream()))
    check(lib.eims_spmm_norm(*map(ptr, args), ptr(hd), H, None, None, p, seed, step, site, 0, ptr(out), N, stream()))
    hm = torch.from_numpy(h) * mask.cpu() * np.float32(1.0 / (1.0 - p))
    ref = torch.zeros(N, H).index_add_(0, g.dst, (hm * torch.from_numpy(norm)[:, None])[g.src])
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=2e-6, atol=1e-6)
    keep = mask.mean().item()
    assert abs(keep - 0.8) < 4 * (0.16 / (N * H)) ** 0.5 + 1e-3
    check(lib.eims_spmm_norm(*map(ptr, args), ptr(hd), H, None, None, p, seed, step, site, 1, ptr(out), N, stream()))
    ref = torch.zeros(N, H).index_add_(0, g.dst, torch.from_numpy(h)[g.src]) * torch.from_numpy(norm)[:, None] * mask.cpu() / (1.0 - p)
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=2e-6, atol=1e-6)


# ------------------------------------------------------------------------------- BN stats
@pytest.mark.parametrize("H,tile_rows,max_atoms", [(256, 64, 64), (256, 16, 64), (128, 64, 40), (1024, 128, 128), (1024, 24, 128), (384, 64, 30)])
def test_spmm_molecule_tiles_bit_identical_to_gather(H, tile_rows, max_atoms):
    """The molecule-tile aggregation (neighbour rows staged in shared memory by cp.async.bulk) must reproduce the
    gather kernel - itself pinned to torch's index_add_ order above - bit for bit: plain, with BatchNorm apply + dropout
    (forward form) and in the backward form with the output mask; a small tile sends the bigger molecules down the
    in-kernel gather path, so both paths and their mix are covered; empty trailing capacity stays untouched."""
    n = 90
    table = synth_molecules(n, max_atoms=max_atoms, seed=13)
    b, rowptr, col, deg, norm = graph_arrays(table, np.arange(n))
    N, E = b["num_nodes"], len(b["src"])
    gptr = np.concatenate([[0], np.cumsum(b["batch_num_nodes"])]).astype(np.int32)
    rng = np.random.default_rng(2)
    hd = dev(rng.standard_normal((N, H)).astype(np.float32))
    scale_d, shift_d = dev(rng.standard_normal(H).astype(np.float32)), dev(rng.standard_normal(H).astype(np.float32))
    dims = make_dims(n, N, E)
    args = (dims, dev(rowptr, torch.int32), dev(col, torch.int32), dev(norm))
    gp = dev(gptr, torch.int32)
    lib = _lib.load()
    cases = [(None, None, 0.0, 0), (scale_d, shift_d, 0.0, 0), (scale_d, shift_d, 0.2, 0), (None, None, 0.0, 1), (None, None, 0.35, 1)]
    for sc, sh, p, mode in cases:
        ref = torch.full((N + 3, H), 7.0, device=DEV)
        got = torch.full((N + 3, H), 7.0, device=DEV)
        check(lib.eims_spmm_norm(*map(ptr, args), ptr(hd), H, ptr(sc), ptr(sh), p, 99, 5, 1, mode, ptr(ref), N, stream()))
        check(lib.eims_spmm_norm_mol(ptr(dims), ptr(gp), *map(ptr, args[1:]), ptr(hd), H, ptr(sc), ptr(sh), p, 99, 5, 1, mode, ptr(got), N,
                                     n + 5, tile_rows, stream()))
        torch.cuda.synchronize()
        assert torch.equal(got.view(torch.int32), ref.view(torch.int32)), (H, tile_rows, p, mode)


@pytest.mark.parametrize("N,H", [(1, 64), (2, 64), (1000, 256), (4097, 1024)])
def test_bn_stats(N, H):
    rng = np.random.default_rng(2)
    z = np.maximum(rng.standard_normal((N, H)) * 3 + 1.0, 0).astype(np.float32)
    z[:, 0] = 5.0 + 1e-3 * rng.standard_normal(N)  # nearly constant column (cancellation check)
    z[:, 1] = 0.0                                    # dead column
    gamma = rng.standard_normal(H).astype(np.float32)
    beta = rng.standard_normal(H).astype(np.float32)
    lib = _lib.load()
    cap = N + 100
    part = torch.zeros(lib.eims_bn_scratch_floats(H, cap), device=DEV)
    rm, rv = torch.zeros(H, device=DEV), torch.ones(H, device=DEV)
    mean, invstd, scale, shift = (torch.empty(H, device=DEV) for _ in range(4))
    zd, dims_d, gamma_d, beta_d = dev(z), make_dims(1, N, 0), dev(gamma), dev(beta)
    for _ in range(2):  # twice: the ticket counter must reset itself
        check(lib.eims_bn_stats(ptr(dims_d), ptr(zd), H, ptr(gamma_d), ptr(beta_d), ptr(rm), ptr(rv),
                                ptr(mean), ptr(invstd), ptr(scale), ptr(shift), ptr(part), cap, stream()))
    zt = torch.from_numpy(z).double()
    m = zt.mean(0)
    var = zt.var(0, unbiased=False)
    np.testing.assert_allclose(mean.cpu().numpy(), m.numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(invstd.cpu().numpy(), (1 / torch.sqrt(var + 1e-5)).numpy(), rtol=2e-6)
    if N > 1:
        # two momentum-0.1 updates from (0, 1): exact values in fp64, and torch's own fp32 BatchNorm
        unb = zt.var(0, unbiased=True)
        rm64 = 0.9 * (0.1 * m) + 0.1 * m
        rv64 = 0.9 * (0.9 + 0.1 * unb) + 0.1 * unb
        np.testing.assert_allclose(rm.cpu().numpy(), rm64.numpy(), rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(rv.cpu().numpy(), rv64.numpy(), rtol=2e-6, atol=1e-7)
        rm_ref, rv_ref = torch.zeros(H), torch.ones(H)
        for _ in range(2):
            torch.nn.functional.batch_norm(torch.from_numpy(z), rm_ref, rv_ref, None, None, True, 0.1, 1e-5)
        np.testing.assert_allclose(rm.cpu().numpy(), rm_ref.numpy(), rtol=5e-5, atol=1e-6)
        np.testing.assert_allclose(rv.cpu().numpy(), rv_ref.numpy(), rtol=5e-5, atol=1e-6)
    ref_scale = torch.from_numpy(gamma).double() / torch.sqrt(var + 1e-5)
    np.testing.assert_allclose(scale.cpu().numpy(), ref_scale.numpy(), rtol=3e-6, atol=1e-7)
    np.testing.assert_allclose(shift.cpu().numpy(), (torch.from_numpy(beta).double() - m * ref_scale).numpy(), rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------- K5
@pytest.mark.parametrize("pooling", ["sum", "mean", "max", "combined"])
def test_readout_first_argmax(pooling):
    H = 256
    table = synth_molecules(30, max_atoms=64, seed=8)
    b = O.batch_graphs([table.mol(i) for i in range(30)])
    N = b["num_nodes"]
    rng = np.random.default_rng(3)
    z = np.maximum(rng.standard_normal((N, H)), 0).astype(np.float32)  # ReLU zeros => many exact ties
    gptr = O.graph_ptr(b["batch_num_nodes"])
    d = O.Dims(hidden_dim=H, pooling=pooling)
    pd = d.pool_dim
    out = torch.empty(30, pd, device=DEV)
    arg = torch.full((30, H), -1, dtype=torch.int32, device=DEV)
    keep = (make_dims(30, N, 0), dev(gptr, torch.int32), dev(z))  # keep device inputs alive across the launch
    check(_lib.load().eims_readout(ptr(keep[0]), ptr(keep[1]), ptr(keep[2]), H, None, None,
                                   _lib.POOLING[pooling], ptr(out), ptr(arg), 30, stream()))
    g = O.Graph(b["src"], b["dst"], b["batch_num_nodes"])
    h = torch.from_numpy(z)
    if pooling in ("max", "combined"):
        mx, a = O.segment_max_first(h, g)
        assert np.array_equal(arg.cpu().numpy(), a.numpy())
        got_mx = out[:, -H:].cpu()
        assert torch.equal(got_mx, mx)
        ties = (h == mx[g.gid]).sum().item() - 30 * H
        assert ties > 0  # the case really has ties
    if pooling in ("sum", "mean", "combined"):
        S = torch.zeros(30, H, dtype=torch.float64).index_add_(0, g.gid, h.double())
        if pooling == "mean":
            S = S / g.batch_num_nodes[:, None]
        np.testing.assert_allclose(out[:, :H].cpu().numpy(), S.numpy(), rtol=2e-6, atol=1e-6)


# ------------------------------------------------------------------------------- K7
@pytest.mark.parametrize("loss_kind", ["mse", "cosine"])
@pytest.mark.parametrize("B,M", [(1, 100), (37, 1000), (16, 500)])
def test_loss_kernel(loss_kind, B, M):
    rng = np.random.default_rng(4)
    logits = (rng.standard_normal((B, M)) * 2).astype(np.float32)
    pk = synth_peaks(B + 5, M, seed=6)
    targets = dense_spectra(*pk, M)
    rows = rng.permutation(B + 5)[:B].astype(np.int32)
    prob, dl = torch.empty(B, M, device=DEV), torch.empty(B, M, device=DEV)
    rl, rc = torch.empty(B, device=DEV), torch.empty(B, device=DEV)
    keep = (make_dims(B, 0, 0), dev(logits), dev(targets), dev(rows))  # keep device inputs alive across the launch
    check(_lib.load().eims_loss_mse_cos(ptr(keep[0]), ptr(keep[1]), ptr(keep[2]), ptr(keep[3]), M,
                                        _lib.LOSS[loss_kind], ptr(prob), ptr(dl), ptr(rl), ptr(rc), B, stream()))
    u = torch.from_numpy(logits).double().requires_grad_(True)
    t = torch.from_numpy(targets[rows]).double()
    p = torch.sigmoid(u)
    loss = O.mse_loss(p, t) if loss_kind == "mse" else O.cosine_loss(p, t)
    loss.backward()
    np.testing.assert_allclose(prob.cpu().numpy(), p.detach().numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(rl.sum().item() / (B * M), O.mse_loss(p, t).item(), rtol=1e-5)
    cos_ref = O.cosine_similarity_batch(p.detach().float(), t.float(), "cupy")
    np.testing.assert_allclose(rc.cpu().numpy(), cos_ref, rtol=1e-5)
    gref = u.grad.numpy()
    assert np.abs(dl.cpu().numpy() - gref).max() <= 1e-5 * np.abs(gref).max()


# ------------------------------------------------------------------------------- K8
def test_adamw_matches_torch():
    n = 10007
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    total = 12
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=1e-4)
    sch = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-3, total_steps=total)
    from eims_b200.engine import onecycle_schedule
    table = onecycle_schedule(total)
    p, m, v = p0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for k in range(total):
        grad = torch.randn(n, generator=g) * (0.1 + k)
        ref.grad = grad.clone()
        opt.step()
        sch.step()
        gd = (grad * 4).to(DEV)  # grad_scale 0.25 undoes the x4 (data-parallel mean)
        st = make_step(lr=table[k][0], beta1=table[k][1], grad_scale=0.25, step=k + 1)
        check(_lib.load().eims_adamw_flat(ptr(p), ptr(gd), ptr(m), ptr(v), n, C.byref(st), stream()))
        assert bool((gd == 0).all())  # zero_grad fused
    err = (p.cpu() - ref.detach()).abs().max().item()
    assert err < 2e-6, err


# ------------------------------------------------------------------------------- peak binning (GCN:166-205)
def _golden(name):
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)))


def test_peaks_to_spectrum_golden():
    """The device kernel reproduces the reference script's own outputs bit for bit: its NumPy
    branch (float64 rounding) and its CuPy branch run with cp = numpy (float32 rounding)."""
    g = _golden("binning_f32.npz")
    for name in ("a", "b"):
        ptr_, mz, inten, M = g[f"{name}_ptr"], g[f"{name}_mz"], g[f"{name}_inten"], int(g[f"{name}_max_mz"])
        out64 = DevicePeaks(ptr_, mz, inten.astype(np.float32), DEV).to_spectrum(M).cpu().numpy()
        assert np.array_equal(out64, g[f"{name}_spec_f64"])
        out32 = DevicePeaks(ptr_, mz.astype(np.float32), inten.astype(np.float32), DEV).to_spectrum(M).cpu().numpy()
        assert np.array_equal(out32, g[f"{name}_spec_f32"])
    b = _golden("binning.npz")
    flat, lens = b["peaks_flat"].reshape(-1, 2), b["peaks_len"]
    peaks, o = [], 0
    for n in lens:
        peaks.append([tuple(r) for r in flat[o:o + n]])
        o += n
    assert np.array_equal(DevicePeaks.from_lists(peaks, DEV).to_spectrum(100).cpu().numpy(), b["spec"])
    pk = synth_peaks(8, 100, seed=5)
    assert np.array_equal(DevicePeaks(*pk, DEV).to_spectrum(100).cpu().numpy(), b["spec2"])


@pytest.mark.parametrize("n,M,f64", [(1, 4, True), (300, 1000, True), (300, 1000, False), (5000, 500, True), (64, 4096, False)])
def test_peaks_to_spectrum_oracle(n, M, f64):
    """Random peak lists (ties, duplicates, out-of-range and negative m/z, non-positive and NaN
    intensities, empty spectra) and a row gather against the oracle; bit-exact."""
    rng = np.random.default_rng(n + M)
    k = rng.integers(0, 200, size=n)
    k[rng.random(n) < 0.05] = 0
    ptr_ = np.zeros(n + 1, np.int64)
    np.cumsum(k, out=ptr_[1:])
    tot = int(ptr_[-1])
    mz = rng.uniform(-2.0, M + 2.0, size=tot)
    half = rng.random(tot) < 0.3
    mz[half] = np.floor(mz[half]) + 0.5
    inten = rng.uniform(-10.0, 999.0, size=tot).astype(np.float32)
    inten[rng.random(tot) < 0.02] = 0.0
    mz = mz.astype(np.float64 if f64 else np.float32)
    rows = rng.permutation(n)[: max(1, n // 2)].astype(np.int32)
    dp = DevicePeaks(ptr_, mz, inten, DEV)
    got = dp.to_spectrum(M, rows=dev(rows, torch.int32)).cpu().numpy()
    lists = peaks_as_lists(ptr_, mz, inten)
    ref_fn = O.peaks_to_spectrum_batch if f64 else O.peaks_to_spectrum_batch_f32
    ref = ref_fn([lists[r] for r in rows], M).astype(np.float32)
    assert np.array_equal(got, ref)
    # NaN intensities never replace a bin (Python's max keeps its first argument)
    if tot:
        inten2 = inten.copy()
        inten2[:: 7] = np.nan
        got2 = DevicePeaks(ptr_, mz, inten2, DEV).to_spectrum(M).cpu().numpy()
        keep = ~np.isnan(inten2)
        ptr2 = np.zeros(n + 1, np.int64)
        np.cumsum([int(keep[ptr_[i]:ptr_[i + 1]].sum()) for i in range(n)], out=ptr2[1:])
        ref2 = ref_fn(peaks_as_lists(ptr2, mz[keep], inten2[keep]), M).astype(np.float32)
        assert np.array_equal(got2, ref2)


def test_loss_with_peak_targets_matches_dense():
    """The loss kernel binning its targets from peak lists == the same kernel reading the dense rows."""
    B, M = 37, 1000
    d = ModelDims(hidden_dim=64, max_mz=M)
    table = synth_molecules(B, max_atoms=12, seed=3)
    pk = synth_peaks(B, M, seed=4)
    dense = dense_spectra(*pk, M)
    plan = Plan(d, B, int(table.node_ptr[-1]), int(2 * table.bond_ptr[-1]), DEV)
    fp = FlatParams(d, DEV)
    fp.load_state_dict(O.init_params(O.Dims(6, 64, 3, M, "combined", 0.0), 0))
    ids = dev(np.random.default_rng(0).permutation(B), torch.int32)
    res = []
    for ds in (DeviceDataset(table, dense, DEV), DeviceDataset(table, None, DEV, peaks=DevicePeaks(*pk, DEV))):
        plan.batch_build(ds, ids, B)
        plan.forward(fp, True, make_step())
        plan.loss(plan._targets(ds), ids, "mse", True)
        res.append((plan.buffer("dlogits", torch.float32, (B, M)).clone(), plan.buffer("row_loss", torch.float32, (B,)).clone(),
                    plan.buffer("row_cos", torch.float32, (B,)).clone()))
        plan.set_peak_targets(None)
    for a, b in zip(*res):
        assert torch.equal(a, b)


# ------------------------------------------------------------------------------- top-k peaks (GCN:610-613)
@pytest.mark.parametrize("n,M,k", [(1, 4, 4), (700, 1000, 5), (64, 500, 1), (33, 4096, 32), (5, 100, 100)])
def test_topk_peaks(n, M, k):
    rng = np.random.default_rng(n * M + k)
    x = rng.random((n, M)).astype(np.float32)
    x[:, ::7] = np.round(x[:, ::7], 1)           # plenty of exact ties
    if n > 2:
        x[1] = 0.25                                # a constant row: ties everywhere
    bins, vals = topk_peaks(dev(x), k)
    bins, vals = bins.cpu().numpy(), vals.cpu().numpy()
    for r in range(n):
        ref = np.argsort(x[r], kind="stable")[-k:][::-1]          # GCN:613 with a stable sort
        assert np.array_equal(bins[r], ref), r
        assert np.array_equal(vals[r], x[r][ref])
        # np.argmax (GCN:612) is the first maximum; the top-1 bin holds the same value
        assert x[r][bins[r, 0]] == x[r][np.argmax(x[r])]
