"""Stand-ins for `dgl` and `rdkit` so that the UNMODIFIED reference script
(/root/reference/templates/ms-pred-gcn-eims-cupy.py) can be imported and run offline.
TEST INFRASTRUCTURE - only tests/golden/make_golden.py and tests use it.

`dgl` is not installable here and the reference does not pin a version, so this module
restates the small DGL surface the script touches from DGL's published source
(python/dgl/{batch.py, nn/pytorch/conv/graphconv.py, nn/pytorch/glob.py, ops/segment.py,
backend/pytorch/sparse.py}).  These semantics are the *unpinned* part of the oracle; all
torch-side behaviour (BatchNorm1d, LayerNorm, Linear, MSELoss, AdamW, OneCycleLR, the
script's own control flow) is the reference's real code when run through this shim.

The rdkit stand-in is duck-typed: `FakeMol` answers exactly the calls `mol_to_dgl_graph` /
`get_atom_features` make (GCN:113-153).
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch
import torch.nn as nn


# ----------------------------------------------------------------------------- dgl
class DGLError(Exception):
    pass


class DGLGraph:
    def __init__(self, src, dst, num_nodes, batch_num_nodes=None, batch_num_edges=None):
        self._src = torch.as_tensor(src, dtype=torch.int64).reshape(-1)
        self._dst = torch.as_tensor(dst, dtype=torch.int64).reshape(-1)
        self._n = int(num_nodes)
        self.ndata = {}
        self._bnn = torch.tensor([self._n]) if batch_num_nodes is None else batch_num_nodes
        self._bne = torch.tensor([len(self._src)]) if batch_num_edges is None else batch_num_edges

    def num_nodes(self):
        return self._n

    number_of_nodes = num_nodes

    def num_edges(self):
        return len(self._src)

    def edges(self):
        return self._src, self._dst

    @property
    def batch_size(self):
        return len(self._bnn)

    def batch_num_nodes(self):
        return self._bnn

    def batch_num_edges(self):
        return self._bne

    def in_degrees(self):
        return torch.bincount(self._dst, minlength=self._n)

    def out_degrees(self):
        return torch.bincount(self._src, minlength=self._n)

    def to(self, device):
        g = DGLGraph(self._src.to(device), self._dst.to(device), self._n, self._bnn, self._bne)
        g.ndata = {k: v.to(device) for k, v in self.ndata.items()}
        return g

    def local_scope(self):
        import contextlib
        return contextlib.nullcontext()


def graph(data, num_nodes=None):
    src, dst = data
    return DGLGraph(list(src), list(dst), num_nodes)


def batch(graphs):
    """dgl.batch: cumulative node offsets, edges concatenated in list order."""
    off, srcs, dsts, nn_, ne_ = 0, [], [], [], []
    for g in graphs:
        srcs.append(g._src + off)
        dsts.append(g._dst + off)
        nn_.append(g._n)
        ne_.append(len(g._src))
        off += g._n
    out = DGLGraph(torch.cat(srcs), torch.cat(dsts), off, torch.tensor(nn_), torch.tensor(ne_))
    keys = graphs[0].ndata.keys()
    out.ndata = {k: torch.cat([g.ndata[k] for g in graphs], dim=0) for k in keys}
    return out


class GraphConv(nn.Module):
    """dgl.nn.GraphConv defaults: norm='both', weight, bias, no activation,
    allow_zero_in_degree=False; weight is (in_feats, out_feats), xavier-uniform; bias 0."""

    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True, activation=None,
                 allow_zero_in_degree=False):
        super().__init__()
        self._in_feats, self._out_feats, self._norm = in_feats, out_feats, norm
        self._allow_zero_in_degree = allow_zero_in_degree
        self.weight = nn.Parameter(torch.Tensor(in_feats, out_feats))
        self.bias = nn.Parameter(torch.Tensor(out_feats))
        nn.init.xavier_uniform_(self.weight)
        nn.init.zeros_(self.bias)

    def forward(self, g, feat):
        if not self._allow_zero_in_degree and bool((g.in_degrees() == 0).any()):
            raise DGLError("There are 0-in-degree nodes in the graph, output for those nodes will be invalid.")
        src, dst = g.edges()
        degs = g.out_degrees().to(feat).clamp(min=1)
        norm = torch.pow(degs, -0.5)
        feat_src = feat * norm.reshape((-1,) + (1,) * (feat.dim() - 1))
        if self._in_feats > self._out_feats:
            feat_src = torch.matmul(feat_src, self.weight)
            rst = torch.zeros(g.num_nodes(), feat_src.shape[1], dtype=feat.dtype, device=feat.device).index_add_(0, dst, feat_src[src])
        else:
            agg = torch.zeros(g.num_nodes(), feat_src.shape[1], dtype=feat.dtype, device=feat.device).index_add_(0, dst, feat_src[src])
            rst = torch.matmul(agg, self.weight)
        degs = g.in_degrees().to(feat).clamp(min=1)
        norm = torch.pow(degs, -0.5)
        rst = rst * norm.reshape((-1,) + (1,) * (feat.dim() - 1))
        return rst + self.bias


def _gid(g):
    return torch.repeat_interleave(torch.arange(g.batch_size), g.batch_num_nodes())


class SumPooling(nn.Module):
    def forward(self, g, feat):
        return torch.zeros(g.batch_size, feat.shape[1], dtype=feat.dtype).index_add_(0, _gid(g), feat)


class AvgPooling(nn.Module):
    def forward(self, g, feat):
        s = torch.zeros(g.batch_size, feat.shape[1], dtype=feat.dtype).index_add_(0, _gid(g), feat)
        return s / g.batch_num_nodes().to(feat.dtype)[:, None]


class MaxPooling(nn.Module):
    """segment_reduce('max'); backward scatters to the saved arg-max, which DGL's CPU
    kernel (strict '<' compare while scanning the segment) makes the first maximum."""

    def forward(self, g, feat):
        N, H = feat.shape
        gid = _gid(g)[:, None].expand(N, H)
        B = g.batch_size
        mx = torch.full((B, H), -float("inf"), dtype=feat.dtype).scatter_reduce(0, gid, feat.detach(), "amax")
        node = torch.arange(N)[:, None].expand(N, H)
        cand = torch.where(feat.detach() == mx[_gid(g)], node, torch.full_like(node, N))
        arg = torch.full((B, H), N, dtype=torch.int64).scatter_reduce(0, gid, cand, "amin")
        return feat.gather(0, arg)


# --------------------------------------------------------------------------- rdkit
class FakeAtom:
    def __init__(self, row):
        self._r = row

    def GetAtomicNum(self):
        return int(self._r[0])

    def GetDegree(self):
        return int(self._r[1])

    def GetFormalCharge(self):
        return int(self._r[2])

    def GetHybridization(self):
        return int(self._r[3])  # the script calls int() on the enum (GCN:119)

    def GetIsAromatic(self):
        return bool(self._r[4])

    def GetTotalNumHs(self):
        return int(self._r[5])


class FakeBond:
    def __init__(self, b, e):
        self._b, self._e = int(b), int(e)

    def GetBeginAtomIdx(self):
        return self._b

    def GetEndAtomIdx(self):
        return self._e


class FakeMol:
    def __init__(self, feat, begin, end):
        self._feat, self._b, self._e = np.asarray(feat), np.asarray(begin), np.asarray(end)

    def GetAtoms(self):
        return [FakeAtom(r) for r in self._feat]

    def GetBonds(self):
        return [FakeBond(b, e) for b, e in zip(self._b, self._e)]

    def GetNumAtoms(self):
        return len(self._feat)


def install():
    """Register the stand-ins as `dgl`, `dgl.nn`, `rdkit`, `rdkit.Chem`, ... in sys.modules."""
    dgl = types.ModuleType("dgl")
    dgl.graph, dgl.batch, dgl.DGLGraph, dgl.DGLError = graph, batch, DGLGraph, DGLError
    dglnn = types.ModuleType("dgl.nn")
    dglnn.GraphConv, dglnn.SumPooling, dglnn.AvgPooling, dglnn.MaxPooling = GraphConv, SumPooling, AvgPooling, MaxPooling
    dgl.nn = dglnn
    rdkit = types.ModuleType("rdkit")
    chem = types.ModuleType("rdkit.Chem")
    chem.MolFromMolFile = lambda *a, **k: None
    chem.MolFromSmiles = lambda *a, **k: None
    allchem = types.ModuleType("rdkit.Chem.AllChem")
    desc = types.ModuleType("rdkit.Chem.Descriptors")
    chem.AllChem, chem.Descriptors = allchem, desc
    rdkit.Chem = chem
    sys.modules.update({"dgl": dgl, "dgl.nn": dglnn, "rdkit": rdkit, "rdkit.Chem": chem,
                        "rdkit.Chem.AllChem": allchem, "rdkit.Chem.Descriptors": desc})
    return dgl


def load_reference(path="/root/reference/templates/ms-pred-gcn-eims-cupy.py"):
    """Import the reference script as a module (its `__main__` guard keeps it passive)."""
    import importlib.util
    install()
    spec = importlib.util.spec_from_file_location("ref_gcn_eims", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
