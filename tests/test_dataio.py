"""MSP reader, packed dataset cache and `--mode preprocess`.  CPU only."""
import os

import numpy as np

from eims_b200 import dataio
from eims_b200 import script as S
from eims_b200.synth import synth_molecules, synth_peaks

NIST = """Name: Ethanol
Formula: C2H6O
MW: 46
CAS#: 64-17-5;  NIST#: 1234;  ID: 7
Comment: test "with; semicolons"
Num Peaks: 7
26 98; 27 224; 29 298; 30 60; 31 999;
45 514 "M-H"; 46 217

Name: No peaks here
ID: 8
Num Peaks: 0

NAME: Acetone
ID: 9
Num Peaks: 3
43\t999
58,271
15 1.5e2
"""


def test_parse_msp_multi_pair_lines(tmp_path):
    p = tmp_path / "lib.msp"
    p.write_text(NIST)
    recs = dataio.parse_msp(str(p))
    assert [dataio.record_field(r, "name") for r in recs] == ["Ethanol", "No peaks here", "Acetone"]
    assert recs[0]["peaks"] == [(26.0, 98.0), (27.0, 224.0), (29.0, 298.0), (30.0, 60.0), (31.0, 999.0), (45.0, 514.0), (46.0, 217.0)]
    assert recs[1]["peaks"] == []
    assert recs[2]["peaks"] == [(43.0, 999.0), (58.0, 271.0), (15.0, 150.0)]
    assert dataio.record_field(recs[0], "CAS#").startswith("64-17-5")
    # the reference's reader drops the multi-pair record (float("98;") raises inside its bare except) ...
    assert dataio.load_peaks_reference(str(p)) is None
    assert S.OptimizedEIMSDataset.load_peaks(str(p)) is None
    # ... and reads what write_reference_msp writes
    q = tmp_path / "one.msp"
    dataio.write_reference_msp(str(q), recs[0])
    assert dataio.load_peaks_reference(str(q)) == recs[0]["peaks"]
    assert dataio.load_peaks_reference(str(tmp_path / "missing.msp")) is None


def test_preprocess_layout(tmp_path, capsys):
    (tmp_path / "mols").mkdir()
    (tmp_path / "mols" / "ID7.MOL").write_text("ethanol molblock\n")
    (tmp_path / "mols" / "Acetone.mol").write_text("acetone molblock\n")
    (tmp_path / "lib.msp").write_text(NIST)
    out = tmp_path / "processed"
    S.main(["--mode", "preprocess", "--msp_file", str(tmp_path / "lib.msp"), "--mol_dir", str(tmp_path / "mols"), "--data_dir", str(out)])
    assert "Preprocessed 2 of 3 records" in capsys.readouterr().out
    mols = sorted(os.listdir(out / "mol_files"))
    assert mols == ["Acetone.mol", "ID7.mol"]
    # the train mode's pairing rule (GCN:546-548) finds the spectra
    for m in mols:
        msp = str(out / "mol_files" / m).replace("mol_files", "msp_files").replace(".mol", ".msp")
        assert dataio.load_peaks_reference(msp)
    assert (out / "mol_files" / "ID7.mol").read_text() == "ethanol molblock\n"
    # without the two inputs the reference's message stays
    S.main(["--mode", "preprocess"])
    assert "Mode not implemented" in capsys.readouterr().out


def test_packed_cache_round_trip(tmp_path):
    table = synth_molecules(20, max_atoms=12, seed=1)
    ptr, mz, inten = synth_peaks(20, 100, seed=2)
    path = str(tmp_path / "cache.npz")
    dataio.save_packed(path, table, ptr, mz, inten, [f"m{i}.mol" for i in range(20)])
    t2, (p2, m2, i2), names = dataio.load_packed(path)
    for a in ("node_ptr", "bond_ptr", "feat", "bond_begin", "bond_end"):
        assert np.array_equal(getattr(table, a), getattr(t2, a))
    assert np.array_equal(p2, ptr) and np.array_equal(m2, mz.astype(np.float64)) and np.array_equal(i2, inten)
    assert names[3] == "m3.mol"
    # pack_graphs(list of MolGraph) reproduces the table
    graphs = [S.MolGraph(*table.mol(g)) for g in range(20)]
    t3 = dataio.pack_graphs(graphs)
    for a in ("node_ptr", "bond_ptr", "feat", "bond_begin", "bond_end"):
        assert np.array_equal(getattr(table, a), getattr(t3, a))
    # a dataset built from the cache needs neither RDKit nor the text files
    cfg = S.Config()
    cfg.max_mz, cfg.use_cupy = 100, False
    ds = S.OptimizedEIMSDataset([], [], cfg, cache_path=path)
    assert len(ds) == 20
    g, spec = ds[5]
    assert np.array_equal(g.ndata["feat"].numpy(), table.mol(5)[0])
    from eims_b200.synth import dense_spectra
    assert np.array_equal(spec.numpy(), dense_spectra(ptr, mz, inten, 100)[5])


def test_reference_msp_writer_round_trips_floats_exactly(tmp_path):
    """`{:g}` rounded 101.49996 to '101.5' (bin 102 instead of 101) and truncated large intensities."""
    from eims_b200.dataio import load_peaks_reference, write_reference_msp
    peaks = [(101.49996, 1234567.25), (57.0, 999.0), (0.1 + 0.2, 1e-7), (300.5000001, 3.0)]
    path = str(tmp_path / "x.msp")
    write_reference_msp(path, {"fields": {"Name": "x"}, "peaks": peaks})
    back = load_peaks_reference(path)
    assert [(float(a), float(b)) for a, b in back] == [(float(a), float(b)) for a, b in peaks]
