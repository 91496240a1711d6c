"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
script (/root/reference/templates/ms-pred-gcn-eims-cupy.py) on seeded synthetic inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

`dgl` and `rdkit` are not installable offline, so `oracle/dgl_shim.py` stands in for them
(see its docstring for what that leaves unpinned).  Everything recorded here is produced
by the reference's own classes and functions: `mol_to_dgl_graph`, `collate_fn`,
`GCNSpectrum.forward`, `nn.MSELoss`, `CuPySpectrumProcessor`, `train_model`.
Weights come from `oracle.gcn_oracle.init_params` (NumPy RNG, portable) and are loaded
into the reference model with `load_state_dict`, so tests can rebuild them from a seed.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "computational-chemistry-ai_b200"))

from eims_b200.synth import dense_spectra, peaks_as_lists, synth_molecules, synth_peaks  # noqa: E402
from oracle import dgl_shim  # noqa: E402
from oracle.gcn_oracle import Dims, init_params, onecycle_table  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)  # fixed summation order inside torch's CPU kernels


def ref_config(ref, d: Dims, **kw):
    c = ref.Config()
    c.hidden_dim, c.num_gcn_layers, c.max_mz, c.pooling, c.dropout = d.hidden_dim, d.num_gcn_layers, d.max_mz, d.pooling, d.dropout
    c.use_cupy = False
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def ref_model(ref, d: Dims, seed):
    m = ref.GCNSpectrum(d.node_feat_dim, ref_config(ref, d))
    m.load_state_dict(init_params(d, seed))
    return m


def ref_graphs(ref, table):
    return [ref.mol_to_dgl_graph(dgl_shim.FakeMol(*table.mol(g))) for g in range(table.num_mols)]


def fwd_bwd_case(ref, d: Dims, n_mols, max_atoms, seed, store_full):
    table = synth_molecules(n_mols, max_atoms=max_atoms, seed=seed)
    pk = synth_peaks(n_mols, d.max_mz, seed=seed + 1)
    target = torch.from_numpy(dense_spectra(*pk, d.max_mz))
    graphs = ref_graphs(ref, table)
    bg, tgt = ref.collate_fn(list(zip(graphs, target)))
    model = ref_model(ref, d, seed=0)
    model.train()
    pred = model(bg, bg.ndata["feat"])
    loss = torch.nn.MSELoss()(pred, tgt)
    loss.backward()
    out = dict(
        n_mols=n_mols, max_atoms=max_atoms, seed=seed,
        src=bg.edges()[0].numpy(), dst=bg.edges()[1].numpy(),
        batch_num_nodes=bg.batch_num_nodes().numpy(), batch_num_edges=bg.batch_num_edges().numpy(),
        feat=bg.ndata["feat"].numpy(), target=tgt.numpy(),
        pred_train=pred.detach().numpy(), loss=np.float64(loss.item()),
        cos_torch=ref.CuPySpectrumProcessor(d.max_mz, False).cosine_similarity_batch(pred.detach(), tgt).numpy(),
    )
    sd_after = model.state_dict()
    for l in range(d.num_gcn_layers):
        out[f"rm{l}"] = sd_after[f"batch_norms.{l}.running_mean"].numpy().copy()
        out[f"rv{l}"] = sd_after[f"batch_norms.{l}.running_var"].numpy().copy()
    for n, p in model.named_parameters():
        g = p.grad.detach().numpy()
        out[f"gnorm:{n}"] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
        if store_full:
            out[f"grad:{n}"] = g
        else:
            out[f"gslice:{n}"] = g.reshape(-1)[:: max(1, g.size // 256)][:256].copy()
    model.eval()
    with torch.no_grad():
        out["pred_eval"] = model(bg, bg.ndata["feat"]).numpy()
    return out


def pooling_case(ref, pooling):
    d = Dims(hidden_dim=64, max_mz=100, pooling=pooling)
    table = synth_molecules(5, max_atoms=10, seed=77)
    graphs = ref_graphs(ref, table)
    bg = sys.modules["dgl"].batch(graphs)
    model = ref_model(ref, d, seed=3)
    model.train()
    p_train = model(bg, bg.ndata["feat"]).detach().numpy()
    model.eval()
    with torch.no_grad():
        p_eval = model(bg, bg.ndata["feat"]).numpy()
    return p_train, p_eval


def train_case(ref):
    d = Dims(hidden_dim=64, max_mz=100, dropout=0.0)
    n_train, n_val, bs, epochs = 12, 4, 4, 2
    table = synth_molecules(n_train + n_val, max_atoms=12, seed=2024)
    pk = synth_peaks(n_train + n_val, d.max_mz, seed=2025)
    spectra = ref.CuPySpectrumProcessor(d.max_mz, False).peaks_to_spectrum_batch(peaks_as_lists(*pk))
    graphs = ref_graphs(ref, table)
    items = [(g, torch.FloatTensor(s)) for g, s in zip(graphs, spectra)]
    mk = lambda lo, hi: torch.utils.data.DataLoader(items[lo:hi], batch_size=bs, shuffle=False, collate_fn=ref.collate_fn, num_workers=0)
    cfg = ref_config(ref, d, batch_size=bs, num_epochs=epochs, use_mixed_precision=False)
    model = ref_model(ref, d, seed=1)
    model, hist = ref.train_model(model, mk(0, n_train), mk(n_train, n_train + n_val), cfg)
    out = dict(n_train=n_train, n_val=n_val, batch_size=bs, epochs=epochs, spectra=spectra)
    for k, v in hist.items():
        out[f"hist:{k}"] = np.asarray(v, np.float64)
    for n, t in model.state_dict().items():
        out[f"sd:{n}"] = t.numpy()
    return out


def main():
    ref = dgl_shim.load_reference()
    small = Dims(hidden_dim=64, max_mz=100)
    np.savez_compressed(os.path.join(HERE, "fwd_bwd_small.npz"), **fwd_bwd_case(ref, small, 6, 12, 11, True))
    full = Dims(hidden_dim=256, max_mz=1000)
    np.savez_compressed(os.path.join(HERE, "fwd_bwd_full.npz"), **fwd_bwd_case(ref, full, 16, 64, 12, False))
    pools = {}
    for p in ("sum", "mean", "max", "combined"):
        pools[f"{p}:train"], pools[f"{p}:eval"] = pooling_case(ref, p)
    np.savez_compressed(os.path.join(HERE, "pooling.npz"), **pools)
    np.savez_compressed(os.path.join(HERE, "train_small.npz"), **train_case(ref))
    # spectrum binning: the reference's NumPy branch on peak lists with duplicates,
    # out-of-range and half-integer m/z (round-half-even)
    peaks = [[(10.5, 5.0), (11.5, 7.0), (10.4, 9.0), (-0.6, 3.0), (99.5, 4.0), (99.4, 2.0), (250.0, 1.0)],
             [], [(3.0, 0.0)], [(0.49, 1.0), (0.5, 2.0), (1.5, 4.0), (2.5, 8.0)]]
    spec = ref.CuPySpectrumProcessor(100, False).peaks_to_spectrum_batch(peaks)
    pk = synth_peaks(8, 100, seed=5)
    spec2 = ref.CuPySpectrumProcessor(100, False).peaks_to_spectrum_batch(peaks_as_lists(*pk))
    np.savez_compressed(os.path.join(HERE, "binning.npz"), spec=spec, spec2=spec2,
                        peaks_flat=np.asarray([p for pl in peaks for p in pl], np.float64),
                        peaks_len=np.asarray([len(pl) for pl in peaks]))
    np.savez_compressed(os.path.join(HERE, "onecycle20.npz"), table=np.asarray(onecycle_table(20), np.float64))
    # feature extraction through the duck-typed Mol
    t = synth_molecules(3, max_atoms=8, seed=9)
    feats = [np.stack([ref.get_atom_features(a) for a in dgl_shim.FakeMol(*t.mol(g)).GetAtoms()]) for g in range(3)]
    np.savez_compressed(os.path.join(HERE, "features.npz"), **{f"f{g}": f for g, f in enumerate(feats)})
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
