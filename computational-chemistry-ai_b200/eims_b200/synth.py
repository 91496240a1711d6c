"""Synthetic EI-MS molecules (no RDKit needed).

The reference script gets its molecules from RDKit (`mol_to_dgl_graph`,
/root/reference/templates/ms-pred-gcn-eims-cupy.py:124-153) and its spectra from MSP
peak lists (`load_peaks`, :260-278).  Neither data set nor RDKit is available offline,
so the benchmarks and parity tests use the generator below: connected
molecular graphs with 1 <= degree <= 4, six raw atom descriptors in the order of
`get_atom_features` (:113-122), and peak lists in the `(mz, intensity)` form that
`peaks_to_spectrum_batch` (:166-205) consumes.

Everything is vectorised over molecules so that 1 M molecules take seconds.
Only NumPy is used; the result is a `MolTable` of flat arrays (the packed layout the
device-resident dataset uses).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_Z_CHOICES = np.array([6, 7, 8, 9, 16, 17], dtype=np.float32)
_Z_PROBS = np.array([0.72, 0.08, 0.12, 0.03, 0.02, 0.03])


@dataclass
class MolTable:
    """Packed molecules: molecule g owns atoms [node_ptr[g], node_ptr[g+1]) and
    bonds [bond_ptr[g], bond_ptr[g+1]); bond ends are molecule-local atom indices."""

    node_ptr: np.ndarray  # int64 [G+1]
    bond_ptr: np.ndarray  # int64 [G+1]
    feat: np.ndarray  # float32 [sum n, 6]
    bond_begin: np.ndarray  # int32 [sum bonds]
    bond_end: np.ndarray  # int32 [sum bonds]

    @property
    def num_mols(self) -> int:
        return len(self.node_ptr) - 1

    def num_atoms(self, g: int) -> int:
        return int(self.node_ptr[g + 1] - self.node_ptr[g])

    def mol(self, g: int):
        """(feat[n,6], begin[b], end[b]) of molecule g."""
        a0, a1 = int(self.node_ptr[g]), int(self.node_ptr[g + 1])
        b0, b1 = int(self.bond_ptr[g]), int(self.bond_ptr[g + 1])
        return self.feat[a0:a1], self.bond_begin[b0:b1], self.bond_end[b0:b1]

    def select(self, ids) -> "MolTable":
        ids = np.asarray(ids, dtype=np.int64)
        n = (self.node_ptr[ids + 1] - self.node_ptr[ids]).astype(np.int64)
        b = (self.bond_ptr[ids + 1] - self.bond_ptr[ids]).astype(np.int64)
        node_ptr = np.zeros(len(ids) + 1, np.int64)
        bond_ptr = np.zeros(len(ids) + 1, np.int64)
        np.cumsum(n, out=node_ptr[1:])
        np.cumsum(b, out=bond_ptr[1:])
        a_idx = _ranges(self.node_ptr[ids], n)
        b_idx = _ranges(self.bond_ptr[ids], b)
        return MolTable(node_ptr, bond_ptr, self.feat[a_idx], self.bond_begin[b_idx], self.bond_end[b_idx])


def _ranges(starts: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """concat(arange(s, s+l) for s, l in zip(starts, lens)) without a Python loop."""
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, np.int64)
    ends = np.cumsum(lens)
    out = np.ones(total, np.int64)
    out[0] = starts[0]
    nz = lens > 0
    first = (ends - lens)[nz]
    s = starts[nz]
    prev_last = np.concatenate([[0], (s + lens[nz] - 1)[:-1]])
    out[first] = s - prev_last
    out[0] = s[0]
    return np.cumsum(out)


_CHUNK = 32768


def _synth_chunk(G: int, A: int, min_atoms: int, rng):
    """One chunk of molecules; returns (n, parent[G,A], ring[G,3,2], deg[G,A])."""
    n = rng.integers(min_atoms, A + 1, size=G)
    deg = np.zeros((G, A), np.int32)
    parent = np.zeros((G, A), np.int32)
    adj = np.zeros((G, A, A), bool)
    rows = np.arange(G)
    window = 6
    for v in range(1, A):
        live = n > v
        lo = max(0, v - window)
        cand = deg[:, lo:v] < 4
        # atom v-1 has degree <= 1 here (or is atom 0), so `cand` is never all-False
        score = rng.random((G, v - lo))
        score[~cand] = -1.0
        pick = (lo + np.argmax(score, axis=1)).astype(np.int32)
        parent[:, v] = pick
        r = rows[live]
        p = pick[live]
        deg[r, p] += 1
        deg[r, v] += 1
        adj[r, v, p] = True
        adj[r, p, v] = True
    ring = np.full((G, 3, 2), -1, np.int32)
    n_ring_target = rng.integers(0, 4, size=G)
    for k in range(3):
        a = (rng.random(G) * n).astype(np.int32)
        b = (rng.random(G) * n).astype(np.int32)
        ok = (k < n_ring_target) & (a != b) & (deg[rows, a] < 4) & (deg[rows, b] < 4) & ~adj[rows, a, b]
        r = rows[ok]
        ring[r, k, 0] = a[ok]
        ring[r, k, 1] = b[ok]
        adj[r, a[ok], b[ok]] = True
        adj[r, b[ok], a[ok]] = True
        deg[r, a[ok]] += 1
        deg[r, b[ok]] += 1
    return n, parent, ring, deg


def synth_molecules(num_mols: int, max_atoms: int = 64, seed: int = 1234, min_atoms: int = 2) -> MolTable:
    """Random connected molecules: a spanning tree (atom v bonds to a uniformly chosen
    earlier atom of degree < 4 among the previous 6) plus up to 3 ring-closure bonds
    between non-bonded atoms of degree < 4.  Bond order: tree bonds (begin = parent,
    end = v) in atom order, then ring closures.  Generated in fixed chunks of 32768
    molecules so memory stays bounded at 1 M molecules."""
    rng = np.random.Generator(np.random.PCG64(seed))
    A = max_atoms
    parts = [_synth_chunk(min(_CHUNK, num_mols - i), A, min_atoms, rng) for i in range(0, num_mols, _CHUNK)]
    n = np.concatenate([p[0] for p in parts])
    parent = np.concatenate([p[1] for p in parts])
    ring = np.concatenate([p[2] for p in parts])
    deg = np.concatenate([p[3] for p in parts])
    G = num_mols
    rows = np.arange(G)
    tree_cnt = n - 1
    ring_ok = ring[:, :, 0] >= 0
    bonds_per = tree_cnt + ring_ok.sum(1)
    node_ptr = np.zeros(G + 1, np.int64)
    bond_ptr = np.zeros(G + 1, np.int64)
    np.cumsum(n, out=node_ptr[1:])
    np.cumsum(bonds_per, out=bond_ptr[1:])
    tot_b = int(bond_ptr[-1])
    bb = np.empty(tot_b, np.int32)
    be = np.empty(tot_b, np.int32)
    mol_of_tree = np.repeat(rows, tree_cnt)
    v_of_tree = _ranges(np.ones(G, np.int64), tree_cnt.astype(np.int64))
    pos_tree = bond_ptr[mol_of_tree] + (v_of_tree - 1)
    bb[pos_tree] = parent[mol_of_tree, v_of_tree]
    be[pos_tree] = v_of_tree
    slot = np.cumsum(ring_ok, axis=1) - 1
    gi, ki = np.nonzero(ring_ok)
    pos_ring = bond_ptr[gi] + tree_cnt[gi] + slot[gi, ki]
    bb[pos_ring] = ring[gi, ki, 0]
    be[pos_ring] = ring[gi, ki, 1]
    # features, in the order of get_atom_features (reference :113-122)
    tot_n = int(node_ptr[-1])
    mol_of_atom = np.repeat(rows, n)
    idx_in_mol = (np.arange(tot_n) - node_ptr[mol_of_atom]).astype(np.int64)
    feat = np.zeros((tot_n, 6), np.float32)
    feat[:, 0] = rng.choice(_Z_CHOICES, size=tot_n, p=_Z_PROBS)
    d = deg[mol_of_atom, idx_in_mol]
    feat[:, 1] = d
    feat[:, 3] = rng.integers(2, 5, size=tot_n)
    feat[:, 4] = rng.integers(0, 2, size=tot_n)
    feat[:, 5] = np.clip(4 - d, 0, 3)
    return MolTable(node_ptr, bond_ptr, feat, bb, be)


def synth_peaks(num_mols: int, max_mz: int = 1000, seed: int = 4321):
    """Per molecule K ~ U[10,150] peaks at integer m/z ~ U[0,max_mz) with intensity
    U(0,999].  Returned packed: (peak_ptr int64[G+1], mz float32[P], inten float32[P])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    k = rng.integers(10, 151, size=num_mols)
    ptr = np.zeros(num_mols + 1, np.int64)
    np.cumsum(k, out=ptr[1:])
    tot = int(ptr[-1])
    mz = rng.integers(0, max_mz, size=tot).astype(np.float32)
    inten = (999.0 * (1.0 - rng.random(tot))).astype(np.float32)
    return ptr, mz, inten


def peaks_as_lists(ptr, mz, inten):
    """The reference's `peaks_list` form: list of list of (mz, intensity) tuples."""
    return [list(zip(mz[ptr[g]:ptr[g + 1]].tolist(), inten[ptr[g]:ptr[g + 1]].tolist()))
            for g in range(len(ptr) - 1)]


def dense_spectra(ptr, mz, inten, max_mz: int) -> np.ndarray:
    """Vectorised equivalent of `peaks_to_spectrum_batch` (reference :193-205): round
    half-to-even, drop bins outside [0,max_mz), max-merge duplicates, divide each row
    by its max (1.0 when the row is all zero).  Bit-identical to the reference's NumPy
    branch (checked in tests/test_oracle_golden.py)."""
    G = len(ptr) - 1
    out = np.zeros((G, max_mz), np.float32)
    row = np.repeat(np.arange(G), np.diff(ptr))
    b = np.round(mz).astype(np.int64)
    ok = (b >= 0) & (b < max_mz)
    np.maximum.at(out, (row[ok], b[ok]), inten[ok])
    mx = out.max(axis=1, keepdims=True)
    mx = np.where(mx > 0, mx, np.float32(1.0)).astype(np.float32)
    return out / mx
