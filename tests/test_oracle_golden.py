"""The oracle (oracle/gcn_oracle.py) against the fixtures recorded from the reference
script itself (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from eims_b200.synth import dense_spectra, peaks_as_lists, synth_molecules, synth_peaks
from oracle import gcn_oracle as O

torch.set_num_threads(1)


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def rebuild_inputs(g, d):
    table = synth_molecules(int(g["n_mols"]), max_atoms=int(g["max_atoms"]), seed=int(g["seed"]))
    mols = [table.mol(i) for i in range(table.num_mols)]
    pk = synth_peaks(table.num_mols, d.max_mz, seed=int(g["seed"]) + 1)
    return mols, dense_spectra(*pk, d.max_mz)


@pytest.mark.parametrize("name,d", [("fwd_bwd_small.npz", O.Dims(hidden_dim=64, max_mz=100)),
                                    ("fwd_bwd_full.npz", O.Dims(hidden_dim=256, max_mz=1000))])
def test_batching_forward_backward(golden_dir, name, d):
    g = load(golden_dir, name)
    mols, target = rebuild_inputs(g, d)
    b = O.batch_graphs(mols)
    # integer work: bit-exact against the reference's mol_to_dgl_graph + collate_fn
    for k in ("src", "dst", "batch_num_nodes", "batch_num_edges"):
        assert np.array_equal(b[k], g[k]), k
    assert np.array_equal(b["feat"].view(np.int32), g["feat"].view(np.int32))
    assert np.array_equal(target.view(np.int32), g["target"].view(np.int32))
    graph, feat = O.Graph.from_mols(mols)
    sd = O.init_params(d, 0)
    pred, loss, grads, _ = O.loss_and_grads(sd, graph, feat, torch.from_numpy(target), d, update_running=True)
    np.testing.assert_allclose(pred.numpy(), g["pred_train"], rtol=2e-6, atol=1e-7)
    assert abs(float(loss) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    for l in range(d.num_gcn_layers):
        np.testing.assert_allclose(sd[f"batch_norms.{l}.running_mean"].numpy(), g[f"rm{l}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(sd[f"batch_norms.{l}.running_var"].numpy(), g[f"rv{l}"], rtol=1e-6, atol=1e-7)
    for n, gr in grads.items():
        gn = float(np.sqrt((gr.numpy().astype(np.float64) ** 2).sum()))
        assert abs(gn - float(g[f"gnorm:{n}"])) <= 1e-5 * max(float(g[f"gnorm:{n}"]), 1e-12), n
        if f"grad:{n}" in g:
            ref = g[f"grad:{n}"]
            assert np.abs(gr.numpy() - ref).max() <= 1e-5 * np.abs(ref).max() + 1e-9, n
        else:
            ref = g[f"gslice:{n}"]
            got = gr.numpy().reshape(-1)[:: max(1, gr.numel() // 256)][:256]
            assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max() + 1e-9, n
    pe, _ = O.forward(sd, graph, feat, d, False)
    np.testing.assert_allclose(pe.detach().numpy(), g["pred_eval"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(O.cosine_similarity_batch(pred, torch.from_numpy(target), "torch"), g["cos_torch"], rtol=1e-6)


def test_pooling_modes(golden_dir):
    g = load(golden_dir, "pooling.npz")
    table = synth_molecules(5, max_atoms=10, seed=77)
    graph, feat = O.Graph.from_mols([table.mol(i) for i in range(5)])
    for p in ("sum", "mean", "max", "combined"):
        d = O.Dims(hidden_dim=64, max_mz=100, pooling=p)
        sd = O.init_params(d, 3)
        tr, _ = O.forward(sd, graph, feat, d, True, update_running=True)
        np.testing.assert_allclose(tr.detach().numpy(), g[f"{p}:train"], rtol=2e-6, atol=1e-7)
        ev, _ = O.forward(sd, graph, feat, d, False)
        np.testing.assert_allclose(ev.detach().numpy(), g[f"{p}:eval"], rtol=2e-6, atol=1e-7)


def test_training_loop(golden_dir):
    """oracle Trainer == reference train_model (AdamW + OneCycleLR + BN buffers + history)."""
    g = load(golden_dir, "train_small.npz")
    d = O.Dims(hidden_dim=64, max_mz=100, dropout=0.0)
    n_train, n_val, bs, epochs = (int(g[k]) for k in ("n_train", "n_val", "batch_size", "epochs"))
    table = synth_molecules(n_train + n_val, max_atoms=12, seed=2024)
    pk = synth_peaks(n_train + n_val, d.max_mz, seed=2025)
    spectra = O.peaks_to_spectrum_batch(peaks_as_lists(*pk), d.max_mz)
    assert np.array_equal(spectra.view(np.int32), g["spectra"].astype(np.float32).view(np.int32))
    assert np.array_equal(dense_spectra(*pk, d.max_mz).view(np.int32), spectra.astype(np.float32).view(np.int32))
    spectra = torch.from_numpy(spectra.astype(np.float32))
    steps_per_epoch = (n_train + bs - 1) // bs
    tr = O.Trainer(O.init_params(d, 1), d, total_steps=epochs * steps_per_epoch)
    hist = {k: [] for k in ("train_loss", "val_loss", "train_cosine", "val_cosine")}
    for _ in range(epochs):
        tl = tc = 0.0
        for s in range(0, n_train, bs):
            ids = list(range(s, min(s + bs, n_train)))
            graph, feat = O.Graph.from_mols([table.mol(i) for i in ids])
            pred, loss = tr.step(graph, feat, spectra[ids])
            tl += loss
            tc += float(O.cosine_similarity_batch(pred, spectra[ids], "torch").mean())
        vl = vc = 0.0
        nvb = 0
        for s in range(n_train, n_train + n_val, bs):
            ids = list(range(s, min(s + bs, n_train + n_val)))
            graph, feat = O.Graph.from_mols([table.mol(i) for i in ids])
            pred = tr.predict(graph, feat)
            vl += float(O.mse_loss(pred, spectra[ids]))
            vc += float(O.cosine_similarity_batch(pred, spectra[ids], "torch").mean())
            nvb += 1
        hist["train_loss"].append(tl / steps_per_epoch)
        hist["train_cosine"].append(tc / steps_per_epoch)
        hist["val_loss"].append(vl / nvb)
        hist["val_cosine"].append(vc / nvb)
    for k, v in hist.items():
        np.testing.assert_allclose(v, g[f"hist:{k}"], rtol=2e-5)
    for n, t in tr.state_dict().items():
        ref = g[f"sd:{n}"]
        if n.endswith("num_batches_tracked"):
            assert int(t) == int(ref)
        else:
            assert np.abs(t.numpy() - ref).max() <= 2e-5 * np.abs(ref).max() + 1e-8, n


def test_binning_and_schedule(golden_dir):
    g = load(golden_dir, "binning.npz")
    flat, lens = g["peaks_flat"].reshape(-1, 2), g["peaks_len"]
    peaks, o = [], 0
    for n in lens:
        peaks.append([tuple(r) for r in flat[o:o + n]])
        o += n
    assert np.array_equal(O.peaks_to_spectrum_batch(peaks, 100), g["spec"])
    pk = synth_peaks(8, 100, seed=5)
    assert np.array_equal(O.peaks_to_spectrum_batch(peaks_as_lists(*pk), 100), g["spec2"])
    assert np.array_equal(dense_spectra(*pk, 100), g["spec2"].astype(np.float32))
    t = load(golden_dir, "onecycle20.npz")["table"]
    np.testing.assert_allclose(np.asarray(O.onecycle_table(20)), t, rtol=1e-12)
    # SURVEY Appendix A.5 spot values
    np.testing.assert_allclose(t[0], [4e-5, 0.95], rtol=1e-6)
    np.testing.assert_allclose(t[5], [1e-3, 0.85], rtol=1e-6)
    np.testing.assert_allclose(t[19], [4e-9, 0.95], rtol=1e-5)


def test_features(golden_dir):
    g = load(golden_dir, "features.npz")
    t = synth_molecules(3, max_atoms=8, seed=9)
    for i in range(3):
        assert np.array_equal(t.mol(i)[0], g[f"f{i}"])


def test_binning_float32_branch(golden_dir):
    """oracle restatement of the CuPy branch (GCN:170-191) == the reference run with cp = numpy."""
    g = load(golden_dir, "binning_f32.npz")
    for name in ("a", "b"):
        peaks = peaks_as_lists(g[f"{name}_ptr"], g[f"{name}_mz"], g[f"{name}_inten"])
        M = int(g[f"{name}_max_mz"])
        assert np.array_equal(O.peaks_to_spectrum_batch_f32(peaks, M), g[f"{name}_spec_f32"])
        assert np.array_equal(O.peaks_to_spectrum_batch(peaks, M).astype(np.float32), g[f"{name}_spec_f64"])
    pk = synth_peaks(8, 100, seed=5)
    assert np.array_equal(O.peaks_to_spectrum_batch_f32(peaks_as_lists(*pk), 100), g["c_spec_f32"])
