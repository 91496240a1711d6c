"""CPU-only tests of the host layer that mirrors the reference script's interface."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from eims_b200 import script as S
from eims_b200.engine import ModelDims, onecycle_schedule, state_dict_order
from eims_b200.hostpath import PackedHostBatch
from eims_b200.synth import peaks_as_lists, synth_molecules, synth_peaks
from oracle import dgl_shim
from oracle import gcn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_featurise_and_collate_match_reference_semantics(golden_dir):
    t = synth_molecules(6, max_atoms=12, seed=11)
    graphs = [S.mol_to_dgl_graph(dgl_shim.FakeMol(*t.mol(i))) for i in range(6)]
    assert S.mol_to_dgl_graph(None) is None
    g = dict(np.load(os.path.join(golden_dir, "fwd_bwd_small.npz")))
    spectra = [torch.zeros(4)] * 6
    bg, sp = S.collate_fn(list(zip(graphs, spectra)))
    src, dst = bg.edges()
    assert np.array_equal(src.numpy(), g["src"]) and np.array_equal(dst.numpy(), g["dst"])  # the reference's own dgl.batch output
    assert np.array_equal(bg.batch_num_nodes().numpy(), g["batch_num_nodes"])
    assert np.array_equal(bg.batch_num_edges().numpy(), g["batch_num_edges"])
    assert np.array_equal(bg.ndata["feat"].numpy(), g["feat"])
    assert sp.shape == (6, 4) and bg.batch_size == 6
    f = dict(np.load(os.path.join(golden_dir, "features.npz")))
    t3 = synth_molecules(3, max_atoms=8, seed=9)
    for i in range(3):
        got = np.stack([S.get_atom_features(a) for a in dgl_shim.FakeMol(*t3.mol(i)).GetAtoms()])
        assert got.dtype == np.float32 and np.array_equal(got, f[f"f{i}"])
    # molecule without bonds: an edgeless graph is built (GCN:146-147); the model raises later
    lone = S.mol_to_dgl_graph(dgl_shim.FakeMol(np.ones((1, 6), np.float32), [], []))
    assert lone.num_nodes() == 1 and lone.num_edges() == 0


def test_packed_host_batch_layout():
    t = synth_molecules(5, max_atoms=9, seed=3)
    tg = np.random.default_rng(0).random((5, 100)).astype(np.float32)
    hb = PackedHostBatch(t, tg, pin=False)
    raw = hb.buf.numpy()
    get = lambda name, dt, n: raw[hb.offsets[name]:hb.offsets[name] + n * np.dtype(dt).itemsize].view(dt)
    assert np.array_equal(get("node_ptr", np.int64, 6), t.node_ptr)
    assert np.array_equal(get("bond_begin", np.int32, len(t.bond_begin)), t.bond_begin)
    assert np.array_equal(get("feat", np.float32, t.feat.size), t.feat.reshape(-1))
    assert np.array_equal(get("targets", np.float32, tg.size), tg.reshape(-1))
    assert all(o % 256 == 0 for o in hb.offsets.values()) and hb.num_edges == 2 * len(t.bond_begin)


def test_packed_host_batch_with_peak_lists():
    """Targets as peak lists: the batch carries peak_ptr / mz / intensity instead of dense rows (3x fewer bytes
    at 1000 bins), m/z in the precision it was given in (float32 = the reference's CuPy branch, else float64)."""
    from eims_b200.synth import synth_peaks
    t = synth_molecules(5, max_atoms=9, seed=3)
    ptr_, mz, inten = synth_peaks(5, 1000, seed=4)
    dense = PackedHostBatch(t, np.zeros((5, 1000), np.float32), pin=False)
    for mz_in, is64 in ((mz, 0), (mz.astype(np.float64), 1)):
        hb = PackedHostBatch(t, None, pin=False, peaks=(ptr_, mz_in, inten))
        raw = hb.buf.numpy()
        get = lambda name, dt, n: raw[hb.offsets[name]:hb.offsets[name] + n * np.dtype(dt).itemsize].view(dt)
        assert hb.has_peaks and not hb.has_targets and hb.mz_is_f64 == is64
        assert np.array_equal(get("peak_ptr", np.int64, 6), ptr_)
        assert np.array_equal(get("peak_mz", np.float64 if is64 else np.float32, len(mz)), mz_in)
        assert np.array_equal(get("peak_inten", np.float32, len(inten)), inten)
        assert "targets" not in hb.offsets and all(o % 256 == 0 for o in hb.offsets.values())
        assert hb.nbytes < dense.nbytes / 2


def test_spectrum_processor_matches_reference(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, "binning.npz")))
    flat, lens = g["peaks_flat"].reshape(-1, 2), g["peaks_len"]
    peaks, o = [], 0
    for n in lens:
        peaks.append([tuple(r) for r in flat[o:o + n]])
        o += n
    proc = S.CuPySpectrumProcessor(100, False)  # the NumPy branch (GCN:193-205); use_cupy=True is the device kernel
    assert np.array_equal(proc.peaks_to_spectrum_batch(peaks), g["spec"])
    pk = synth_peaks(8, 100, seed=5)
    assert np.array_equal(proc.peaks_to_spectrum_batch(peaks_as_lists(*pk)), g["spec2"])
    a, b = torch.rand(4, 100), torch.rand(4, 100)
    np.testing.assert_allclose(proc.cosine_similarity_batch(a, b).numpy(), O.cosine_similarity_batch(a, b, "torch"), rtol=1e-6)


def test_model_surface_and_checkpoint_format(tmp_path):
    cfg = S.Config()
    assert (cfg.max_mz, cfg.hidden_dim, cfg.num_gcn_layers, cfg.dropout, cfg.pooling, cfg.batch_size, cfg.num_epochs,
            cfg.learning_rate, cfg.weight_decay, cfg.model_save_path) == (500, 256, 3, 0.2, "combined", 64, 100, 1e-3, 1e-4, "gcn_eims_model.pth")
    model = S.GCNSpectrum(6, cfg)
    assert sum(p.numel() for p in model.parameters()) == 658_932  # SURVEY A.6, reference defaults
    sd = model.state_dict()
    spec = {n: (s, dt) for n, s, dt in O.state_dict_spec(O.Dims(max_mz=500))}
    assert list(sd) == list(spec)
    for n, t in sd.items():
        assert tuple(t.shape) == tuple(spec[n][0]) and t.dtype == spec[n][1], n
    # initial distributions: GraphConv xavier bound, zero biases, unit norms
    w0 = sd["gcn_layers.0.weight"]
    assert w0.abs().max() <= np.sqrt(6.0 / (6 + 256)) + 1e-6 and float(sd["gcn_layers.1.bias"].abs().max()) == 0.0
    assert float(sd["batch_norms.0.running_var"].min()) == 1.0 and float(sd["spectrum_predictor.1.weight"].min()) == 1.0
    # checkpoint dict exactly as GCN:589-593; round trip through torch.save / load_state_dict
    path = tmp_path / "ck.pth"
    torch.save({"model_state_dict": sd, "config": cfg.__dict__, "history": {k: [] for k in ("train_loss", "val_loss", "train_cosine", "val_cosine")}}, path)
    ck = torch.load(path)
    m2 = S.GCNSpectrum(6, S.Config(**ck["config"]))
    m2.load_state_dict(ck["model_state_dict"])
    assert all(torch.equal(a, b) for a, b in zip(m2.state_dict().values(), sd.values()))
    # an oracle / reference state dict loads too
    m2.load_state_dict(O.init_params(O.Dims(max_mz=500), 3))
    with pytest.raises(RuntimeError):
        m2.load_state_dict({"nope": torch.zeros(1)})
    # no CPU path: forward on a CPU model fails loudly
    t = synth_molecules(2, max_atoms=5, seed=1)
    g = S.batch([S.MolGraph(*t.mol(i)) for i in range(2)])
    with pytest.raises(RuntimeError, match="no CPU path"):
        model.eval()
        model(g, g.ndata["feat"])


def test_cli_surface():
    p = S.build_parser()
    a = p.parse_args([])
    assert (a.mode, a.data_dir, a.batch_size, a.num_epochs, a.use_cupy, a.smiles, a.msp_file, a.mol_dir) == \
        ("train", "processed_data", 64, 100, True, None, None, None)
    assert p.parse_args(["--mode", "predict", "--smiles", "CCO"]).smiles == "CCO"
    with pytest.raises(SystemExit):
        p.parse_args(["--mode", "bogus"])
    script = os.path.join(ROOT, "computational-chemistry-ai_b200", "ms_pred_gcn_eims_b200.py")
    out = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "Demo mode - showing example usage" in out.stdout and "--mode train --data_dir processed_data" in out.stdout
    out = subprocess.run([sys.executable, script, "--mode", "preprocess"], capture_output=True, text=True, timeout=120)
    assert "Mode not implemented" in out.stdout


def test_onecycle_schedule_equals_torch():
    for total in (4, 20, 37, 300):
        np.testing.assert_allclose(np.array(onecycle_schedule(total)), np.array(O.onecycle_table(total)), rtol=1e-12)
    assert len(state_dict_order(ModelDims())) == 31


def test_bench_knows_the_work_of_every_stage():
    """bench.py reports achieved GB/s / TFLOP/s per stage: every stage name the library can time must have an
    algorithmic-work entry (a missing one would silently drop that stage from the roofline table)."""
    import ctypes as C
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from eims_b200 import _lib
    lib = _lib.load()
    lib.eims_plan_stage_name.restype = C.c_char_p
    names = [lib.eims_plan_stage_name(k).decode() for k in range(lib.eims_plan_num_stages())]
    work = bench.stage_work(16896.0, 34806.0, 512, 787432)
    assert set(names) <= set(work), sorted(set(names) - set(work))
    for name, (bound, w) in work.items():
        assert bound in ("hbm", "tensor") and w >= 0, name
    # the grouped launches count both gradients
    assert work["gemm_gcn_bwd"][1] == work["gemm_gcn_wgrad"][1] + work["gemm_gcn_dgrad"][1]


def test_host_packer_equals_python_collate():
    """`eims_host_pack_batch` (C, one call) produces byte for byte the packed batch the Python collate
    (MolTable.select + PackedHostBatch) builds: dense targets and both peak-list precisions; capacity errors."""
    import numpy as np
    from eims_b200 import _lib
    from eims_b200.hostpath import HostDataset, HostPacker, PackedHostBatch
    from eims_b200.synth import _ranges, dense_spectra, synth_molecules, synth_peaks
    n, M = 300, 200
    table = synth_molecules(n, max_atoms=24, seed=5)
    pk = synth_peaks(n, M, seed=6)
    dense = dense_spectra(*pk, M)
    ids = np.random.default_rng(1).permutation(n)[:77].astype(np.int32)
    ref = PackedHostBatch(table.select(ids), dense[ids], pin=False)
    got = HostPacker(HostDataset(table, dense), M, 1 << 20, n_buffers=2, pin=False).pack(ids)
    assert got.nbytes == ref.nbytes and {k: v for k, v in got.offsets.items()} == ref.offsets
    assert (got.num_graphs, got.num_nodes, got.num_edges, got.feat_dim) == (ref.num_graphs, ref.num_nodes, ref.num_edges, ref.feat_dim)
    for name, o in ref.offsets.items():   # compare the live bytes of every section (padding is unspecified)
        nxt = min([v for v in ref.offsets.values() if v > o] + [ref.nbytes])
        live = {"node_ptr": 8 * 78, "bond_ptr": 8 * 78, "bond_begin": 2 * ref.num_edges, "bond_end": 2 * ref.num_edges,
                "feat": 4 * ref.num_nodes * 6, "targets": 4 * 77 * M}[name]
        assert live <= nxt - o
        assert np.array_equal(got.buf.numpy()[o:o + live], ref.buf.numpy()[o:o + live]), name
    for mz_dt in (np.float32, np.float64):
        kk = np.diff(pk[0])[ids]
        pp = np.zeros(len(ids) + 1, np.int64)
        np.cumsum(kk, out=pp[1:])
        sel = _ranges(pk[0][ids], kk)
        refp = PackedHostBatch(table.select(ids), None, pin=False, peaks=(pp, pk[1][sel].astype(mz_dt), pk[2][sel]))
        gotp = HostPacker(HostDataset(table, None, peaks=(pk[0], pk[1].astype(mz_dt), pk[2])), M, 1 << 20, 2, pin=False).pack(ids)
        assert gotp.has_peaks and not gotp.has_targets and gotp.mz_is_f64 == int(mz_dt == np.float64)
        assert gotp.nbytes == refp.nbytes and gotp.offsets == refp.offsets
        nb = int(pp[-1])
        for name, live in (("peak_ptr", 8 * 78), ("peak_mz", np.dtype(mz_dt).itemsize * nb), ("peak_inten", 4 * nb)):
            o = refp.offsets[name]
            assert np.array_equal(gotp.buf.numpy()[o:o + live], refp.buf.numpy()[o:o + live]), name
    with pytest.raises(_lib.EimsError):
        HostPacker(HostDataset(table, dense), M, 1024, 1, pin=False).pack(ids)
    with pytest.raises(_lib.EimsError):
        HostPacker(HostDataset(table, dense), M, 1 << 20, 1, pin=False).pack(np.array([n], np.int32))
