"""Known-answer tests that freeze the DGL behaviours this repo relies on but cannot run (SURVEY Appendix B;
`dgl` is not installable here): hand-computed small graphs, checked against BOTH restatements - the oracle
(oracle/gcn_oracle.py) and the shim the golden fixtures were recorded with (oracle/dgl_shim.py).  CPU only."""
import math

import numpy as np
import pytest
import torch

from oracle import dgl_shim
from oracle import gcn_oracle as O


def bonds_to_edges(bonds):
    """GCN:139-143: bond k -> edges 2k = begin->end, 2k+1 = end->begin."""
    src, dst = [], []
    for b, e in bonds:
        src += [b, e]
        dst += [e, b]
    return np.array(src, np.int64), np.array(dst, np.int64)


CASES = {
    # name: (n atoms, bonds, expected normalised adjacency D^-1/2 A D^-1/2 written out by hand)
    "ethane": (2, [(0, 1)], [[0, 1], [1, 0]]),
    "propane": (3, [(0, 1), (1, 2)], [[0, 1 / math.sqrt(2), 0], [1 / math.sqrt(2), 0, 1 / math.sqrt(2)], [0, 1 / math.sqrt(2), 0]]),
    "cyclopropane": (3, [(0, 1), (1, 2), (2, 0)], [[0, .5, .5], [.5, 0, .5], [.5, .5, 0]]),
    "isobutane": (4, [(0, 1), (0, 2), (0, 3)], [[0, 1 / math.sqrt(3)] + [1 / math.sqrt(3)] * 2, [1 / math.sqrt(3), 0, 0, 0],
                                                  [1 / math.sqrt(3), 0, 0, 0], [1 / math.sqrt(3), 0, 0, 0]]),
}


@pytest.mark.parametrize("name", list(CASES))
def test_graphconv_is_normalised_adjacency_without_self_loops(name):
    """GraphConv(norm='both') = ((A (h*c)) W) * c + b with c = deg^-1/2 and NO self-loops: on a hand-written
    graph the aggregate equals A_hat @ X with the closed-form A_hat (zero diagonal)."""
    n, bonds, a_hat = CASES[name]
    a_hat = np.array(a_hat, np.float64)
    assert np.all(np.diag(a_hat) == 0)
    src, dst = bonds_to_edges(bonds)
    g = O.Graph(src, dst, [n])
    x = torch.arange(1, n * 3 + 1, dtype=torch.float64).reshape(n, 3)
    W = torch.tensor([[1., 2.], [0., -1.], [.5, .25]], dtype=torch.float64)
    b = torch.tensor([.1, -.2], dtype=torch.float64)
    r, a = O.graph_conv(g, x, W, b)
    # a = A (x * c_src): the source-side half of the normalisation; r carries both
    # (the degree normalisation itself is computed in float32, as torch.pow on the float32 degrees does)
    np.testing.assert_allclose(r.numpy(), a_hat @ x.numpy() @ W.numpy() + b.numpy(), rtol=1e-6)
    # the shim the golden fixtures were recorded with agrees
    conv = dgl_shim.GraphConv(3, 2)
    with torch.no_grad():
        conv.weight.copy_(W.float())
        conv.bias.copy_(b.float())
    out = conv(dgl_shim.graph((torch.from_numpy(src), torch.from_numpy(dst)), num_nodes=n), x.float())
    np.testing.assert_allclose(out.detach().numpy(), r.numpy(), rtol=1e-5)


def test_edge_order_and_batch_offsets():
    """dgl.batch: node ids of graph g are offset by the atoms before it, edges stay in graph order, then bond order,
    begin->end before end->begin (GCN:139-143, 295)."""
    m0 = (np.zeros((2, 6), np.float32), np.array([0], np.int32), np.array([1], np.int32))
    m1 = (np.zeros((3, 6), np.float32), np.array([2, 0], np.int32), np.array([1, 1], np.int32))
    b = O.batch_graphs([m0, m1])
    assert b["src"].tolist() == [0, 1, 4, 3, 2, 3]
    assert b["dst"].tolist() == [1, 0, 3, 4, 3, 2]
    assert b["batch_num_nodes"].tolist() == [2, 3] and b["num_nodes"] == 5
    rowptr, col, deg = O.csr_by_dst(b["src"], b["dst"], 5)
    assert rowptr.tolist() == [0, 1, 2, 3, 5, 6] and deg.tolist() == [1, 1, 1, 2, 1]
    assert col.tolist() == [1, 0, 3, 4, 2, 3]  # sources of each destination row, ascending edge id


def test_degree_norm_bits_and_clamp():
    """torch.pow(deg.clamp(min=1), -0.5) on the CPU = fl(1/fl(sqrt(d))): differs from the correctly rounded
    1/sqrt(d) at d = 6 and 7 (and at ten larger degrees up to 64), which is why K1 uses __fdiv_rn(1, __fsqrt_rn(d))
    and not rsqrtf; degree 0 is clamped to 1 (norm 1)."""
    d = np.arange(0, 65)
    got = O.degree_norm(d)
    ref = torch.pow(torch.from_numpy(d).float().clamp(min=1), -0.5).numpy()
    assert np.array_equal(got.view(np.int32), ref.view(np.int32))
    two_step = (np.float32(1) / np.sqrt(np.maximum(d, 1).astype(np.float32))).astype(np.float32)
    assert np.array_equal(got.view(np.int32), two_step.view(np.int32))
    exact = (1.0 / np.sqrt(np.maximum(d, 1).astype(np.float64))).astype(np.float32)
    diff = [int(k) for k in d[got != exact]]
    assert [k for k in diff if k <= 8] == [6, 7]  # the degrees a molecule can have; 12 of the 64 values differ in all
    assert got[0] == 1.0


def test_isolated_atom_raises_like_graphconv():
    g = O.Graph(np.array([0, 1]), np.array([1, 0]), [3])  # atom 2 has no bond
    with pytest.raises(O.ZeroInDegreeError):
        O.graph_conv(g, torch.ones(3, 2), torch.ones(2, 2), torch.zeros(2))


def test_max_pool_gradient_goes_to_the_first_arg_max():
    """DGL's SegmentCmp keeps the FIRST maximum of a segment; torch's scatter_reduce('amax') would split the
    gradient between ties."""
    g = O.Graph(np.array([0, 1, 1, 2, 3, 4]), np.array([1, 0, 2, 1, 4, 3]), [3, 2])
    h = torch.tensor([[1., 5.], [3., 5.], [3., 2.], [7., 0.], [7., 0.]], requires_grad=True)
    mx, arg = O.segment_max_first(h, g)
    assert mx.tolist() == [[3., 5.], [7., 0.]]
    assert arg.tolist() == [[1, 0], [3, 3]]
    mx.sum().backward()
    assert h.grad.tolist() == [[0., 1.], [1., 0.], [0., 0.], [1., 1.], [0., 0.]]


def test_batchnorm_running_statistics_after_one_step():
    """nn.BatchNorm1d over ALL nodes of the batch: momentum 0.1, running_var from the UNBIASED batch variance,
    normalisation with the biased one, eps 1e-5."""
    d = O.Dims(6, 64, 1, 100, "combined", 0.0)
    sd = O.init_params(d, 0)
    mols = [(np.random.default_rng(1).random((3, 6)).astype(np.float32), np.array([0, 1], np.int32), np.array([1, 2], np.int32)),
            (np.random.default_rng(2).random((2, 6)).astype(np.float32), np.array([0], np.int32), np.array([1], np.int32))]
    g, feat = O.Graph.from_mols(mols)
    _, aux = O.forward(sd, g, feat, d, True, keep=True)
    z = aux["z0"].double()
    n = z.shape[0]
    mean, var_b = z.mean(0), z.var(0, unbiased=False)
    np.testing.assert_allclose(sd["batch_norms.0.running_mean"].numpy(), 0.1 * mean.numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(sd["batch_norms.0.running_var"].numpy(), 0.9 + 0.1 * var_b.numpy() * n / (n - 1), rtol=1e-5)
    np.testing.assert_allclose(aux["bn0"].detach().double().numpy(), ((z - mean) / torch.sqrt(var_b + 1e-5)).numpy(), rtol=1e-4, atol=1e-5)
    assert int(sd["batch_norms.0.num_batches_tracked"]) == 1
