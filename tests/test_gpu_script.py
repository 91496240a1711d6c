"""The drop-in script surface on a B200: written the way the reference's own tests would be
(build graphs, collate, call the model, train_model, predict), checked against the golden
fixtures recorded from the reference script and against the oracle."""
import os

import numpy as np
import pytest
import torch

from eims_b200 import _lib
from eims_b200 import script as S
from eims_b200.synth import dense_spectra, synth_molecules, synth_peaks
from oracle import dgl_shim
from oracle import gcn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def cfg(**kw):
    c = S.Config()
    c.hidden_dim, c.max_mz, c.dropout = 64, 100, 0.0
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def test_forward_eval_train_and_autograd(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "fwd_bwd_small.npz")))
    table = synth_molecules(6, max_atoms=12, seed=11)
    graphs = [S.mol_to_dgl_graph(dgl_shim.FakeMol(*table.mol(i))) for i in range(6)]
    targets = torch.from_numpy(g["target"])
    model = S.GCNSpectrum(6, cfg()).to(DEV)
    model.load_state_dict(O.init_params(O.Dims(hidden_dim=64, max_mz=100), 0))
    bg, tg = S.collate_fn(list(zip(graphs, targets)))
    bg = bg.to(DEV)
    model.train()
    pred = model(bg, bg.ndata["feat"])                       # GCN:426
    loss = torch.nn.MSELoss()(pred, tg.to(DEV))              # GCN:427
    loss.backward()                                          # GCN:428 (through eims_backward)
    assert rel_err(pred.detach().cpu().numpy(), g["pred_train"]) < 1e-4
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * float(g["loss"])
    grads = model._views(model.flat.grad)
    for n, gr in grads.items():
        assert rel_err(gr.cpu().numpy(), g[f"grad:{n}"]) < 1e-4, n
    sd = model.state_dict()
    for l in range(3):
        assert rel_err(sd[f"batch_norms.{l}.running_mean"].cpu().numpy(), g[f"rm{l}"]) < 1e-4
        assert int(sd[f"batch_norms.{l}.num_batches_tracked"]) == 1
    model.eval()
    with torch.no_grad():
        ev = model(bg, bg.ndata["feat"])
    assert rel_err(ev.cpu().numpy(), g["pred_eval"]) < 1e-4
    cos = S.CuPySpectrumProcessor(100, True).cosine_similarity_batch(pred.detach(), tg.to(DEV))
    np.testing.assert_allclose(cos.cpu().numpy(), g["cos_torch"], rtol=1e-4)
    # predict path: one molecule per call == batched
    single = np.stack([S.predict_graphs(model, [gr])[0] for gr in graphs])
    assert rel_err(single, ev.cpu().numpy()) < 1e-5


def test_isolated_atom_raises_like_dgl():
    model = S.GCNSpectrum(6, cfg()).to(DEV).eval()
    lone = S.mol_to_dgl_graph(dgl_shim.FakeMol(np.ones((1, 6), np.float32), [], []))
    ok = S.MolGraph(np.ones((2, 6), np.float32), [0], [1])
    with pytest.raises(_lib.ZeroInDegreeError):
        with torch.no_grad():
            model(S.batch([ok, lone]).to(DEV), None)


def test_train_model_matches_reference_history(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "train_small.npz")))
    n_train, n_val, bs, epochs = (int(g[k]) for k in ("n_train", "n_val", "batch_size", "epochs"))
    table = synth_molecules(n_train + n_val, max_atoms=12, seed=2024)
    spectra = dense_spectra(*synth_peaks(n_train + n_val, 100, seed=2025), 100)
    graphs = [S.mol_to_dgl_graph(dgl_shim.FakeMol(*table.mol(i))) for i in range(n_train + n_val)]
    items = [(gr, torch.from_numpy(s)) for gr, s in zip(graphs, spectra)]
    mk = lambda lo, hi: torch.utils.data.DataLoader(items[lo:hi], batch_size=bs, shuffle=False, collate_fn=S.collate_fn, num_workers=0)
    c = cfg(batch_size=bs, num_epochs=epochs)
    model = S.GCNSpectrum(6, c).to(DEV)
    model.load_state_dict(O.init_params(O.Dims(hidden_dim=64, max_mz=100), 1))
    model, hist = S.train_model(model, mk(0, n_train), mk(n_train, n_train + n_val), c, verbose=False)
    for k, v in hist.items():
        np.testing.assert_allclose(v, g[f"hist:{k}"], rtol=2e-4)
    sd = model.state_dict()
    assert list(sd) == [k[3:] for k in g if k.startswith("sd:")]
    for n in ("spectrum_predictor.8.weight", "spectrum_predictor.4.weight", "gcn_layers.2.weight", "batch_norms.2.running_var"):
        assert rel_err(sd[n].cpu().numpy(), g[f"sd:{n}"]) < 2e-3, n
    assert int(sd["batch_norms.0.num_batches_tracked"]) == int(g["sd:batch_norms.0.num_batches_tracked"])


def test_processor_device_branch_matches_reference_cupy_branch(golden_dir):
    """CuPySpectrumProcessor(use_cupy=True) runs the device kernel and equals the reference's CuPy
    branch (recorded with cp = numpy); use_cupy=False equals its NumPy branch."""
    from eims_b200.synth import peaks_as_lists
    g = dict(np.load(os.path.join(golden_dir, "binning_f32.npz")))
    for name in ("a", "b"):
        peaks = peaks_as_lists(g[f"{name}_ptr"], g[f"{name}_mz"], g[f"{name}_inten"])
        M = int(g[f"{name}_max_mz"])
        assert np.array_equal(S.CuPySpectrumProcessor(M, True).peaks_to_spectrum_batch(peaks), g[f"{name}_spec_f32"])
        assert np.array_equal(S.CuPySpectrumProcessor(M, False).peaks_to_spectrum_batch(peaks).astype(np.float32), g[f"{name}_spec_f64"])


def test_predict_spectrum_batch(monkeypatch):
    """Batched SMILES prediction == the reference-style one-call-per-molecule loop; invalid SMILES
    give None as in GCN:497-502; the device top-k equals the printed report of GCN:610-613."""
    import sys
    dgl_shim.install()
    table = synth_molecules(9, max_atoms=14, seed=21)
    mols = {f"M{g}": dgl_shim.FakeMol(*table.mol(g)) for g in range(9)}
    monkeypatch.setattr(sys.modules["rdkit.Chem"], "MolFromSmiles", lambda smi, *a, **k: mols.get(smi))
    config = cfg()
    model = S.GCNSpectrum(6, config).to(DEV)
    model.load_state_dict(O.init_params(O.Dims(6, 64, 3, 100, "combined", 0.0), 3))
    smiles = ["M0", "bad", "M3", "M8", "M1", "also bad", "M2"]
    spectra, bins, vals = S.predict_spectrum_batch(model, smiles, config, batch_size=3, top_k=5)
    for smi, sp, bi, va in zip(smiles, spectra, bins, vals):
        one = S.predict_spectrum(model, smi, config)
        if smi not in mols:
            assert sp is None and one is None and bi is None
            continue
        assert rel_err(sp, one) < 1e-5
        assert np.array_equal(bi, np.argsort(sp, kind="stable")[-5:][::-1])
        assert np.array_equal(va, sp[bi])
    assert S.predict_spectrum_batch(model, ["bad"], config) == [None]
