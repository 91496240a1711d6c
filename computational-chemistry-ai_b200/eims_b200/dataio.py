"""On-disk formats either side of the hot path: the NIST MSP reader, the
packed binary dataset cache, and `--mode preprocess`.

* `load_peaks_reference` is the reference's reader verbatim in behaviour (GCN:260-278): one
  `mz intensity` pair per line after a `Num Peaks:` line, any parse error drops the molecule.
* `parse_msp` reads real NIST MSP libraries: many records per file, `Key: value` header
  lines, peak lines holding one OR several `mz intensity` pairs separated by `;`, `,`, tabs
  or blanks, optional quoted annotations.  The reference's loader raises on such a line
  (`float("41;")`) and its bare `except` then silently drops the molecule.
* `save_packed` / `load_packed`: one `.npz` with the flat arrays the device dataset is made of
  (features, bonds, offsets, peak lists), so a later run skips RDKit and the MSP text entirely.
* `preprocess`: splits one MSP library + a directory of MOL files into the
  `data_dir/mol_files/*.mol` / `data_dir/msp_files/*.msp` layout `--mode train` expects
  (GCN:546-548), i.e. the mode the reference advertises (GCN:626-627) but does not implement
  (GCN:619-620).

MOL parsing itself stays with RDKit (featurisation is host-side Python by design).
"""
from __future__ import annotations

import os
import re
import shutil

import numpy as np

from .synth import MolTable

_PAIR = re.compile(r"([-+]?\d*\.?\d+(?:[eE][-+]?\d+)?)[\s,:]+([-+]?\d*\.?\d+(?:[eE][-+]?\d+)?)")
_QUOTED = re.compile(r'"[^"]*"')


def load_peaks_reference(msp_file):
    """GCN:260-278, unchanged in behaviour (including what it rejects)."""
    peaks = []
    try:
        with open(msp_file, "r") as f:
            reading = False
            for line in f.readlines():
                if reading:
                    parts = line.strip().split()
                    if len(parts) >= 2:
                        peaks.append((float(parts[0]), float(parts[1])))
                elif "Num Peaks:" in line or "NUM PEAKS:" in line:
                    reading = True
        return peaks if peaks else None
    except Exception:
        return None


def parse_peak_line(line: str):
    """All `mz intensity` pairs of one MSP peak line (annotations in quotes are ignored)."""
    line = _QUOTED.sub(" ", line)
    out = []
    for chunk in line.split(";"):
        m = _PAIR.search(chunk)
        if m:
            out.append((float(m.group(1)), float(m.group(2))))
    return out


def parse_msp(path_or_lines):
    """Records of a NIST MSP library: list of dicts {"fields": {key: value}, "peaks": [(mz, i), ...]}.
    A record starts at a `Name:` line (or after a blank line) and ends at the next blank line /
    `Name:`; keys are kept as written, lookups below are case-insensitive."""
    if isinstance(path_or_lines, (str, os.PathLike)):
        with open(path_or_lines, "r", errors="replace") as f:
            lines = f.read().splitlines()
    else:
        lines = list(path_or_lines)
    records, cur, in_peaks = [], None, False

    def close():
        nonlocal cur, in_peaks
        if cur is not None and (cur["fields"] or cur["peaks"]):
            records.append(cur)
        cur, in_peaks = None, False

    for raw in lines:
        line = raw.strip()
        if not line:
            close()
            continue
        is_name = line.lower().startswith("name:")
        if is_name and cur is not None and (in_peaks or cur["fields"]):
            close()
        if cur is None:
            cur = {"fields": {}, "peaks": []}
        if not in_peaks and ":" in line and not line[0].isdigit():
            key, _, val = line.partition(":")
            key, val = key.strip(), val.strip()
            # NIST writes several fields on one line: "CAS#: 64-17-5;  NIST#: 1234;  ID: 7"
            if ";" in val and not key.lower().startswith("comment") and '"' not in val:
                first, *rest = val.split(";")
                subs = [r.partition(":") for r in rest]
                if all(sep and re.fullmatch(r"[A-Za-z][A-Za-z0-9#_ ]*", k.strip()) for k, sep, _ in subs):
                    val = first.strip()
                    for k, _, v in subs:
                        cur["fields"][k.strip()] = v.strip()
            cur["fields"][key] = val
            if key.lower() == "num peaks":
                in_peaks = True
                # some writers put the first pairs on the same line
                cur["peaks"].extend(parse_peak_line(val) if ";" in val else [])
            continue
        cur["peaks"].extend(parse_peak_line(line))
    close()
    return records


def record_field(rec, *names):
    low = {k.lower(): v for k, v in rec["fields"].items()}
    for n in names:
        if n.lower() in low and low[n.lower()] != "":
            return low[n.lower()]
    return None


def record_key(rec, index):
    """File stem used to pair a record with its MOL file: ID / NIST# / DB# / CAS# / sanitised Name."""
    cands = []
    for names, fmt in ((("ID",), "ID{}"), (("ID",), "{}"), (("NIST#", "NISTNO"), "{}"), (("DB#",), "{}"), (("CAS#", "CASNO"), "{}")):
        v = record_field(rec, *names)
        if v:
            cands.append(fmt.format(v.split(";")[0].strip()))
    name = record_field(rec, "Name")
    if name:
        cands.append(re.sub(r"[^A-Za-z0-9._-]+", "_", name).strip("_"))
    cands.append(f"record{index}")
    return cands


def write_reference_msp(path, rec):
    """One record in the only form the reference's reader accepts: one pair per line (GCN:266-274)."""
    with open(path, "w") as f:
        for k, v in rec["fields"].items():
            if k.lower() != "num peaks":
                f.write(f"{k}: {v}\n")
        f.write(f"Num Peaks: {len(rec['peaks'])}\n")
        for mz, inten in rec["peaks"]:
            # repr() round-trips a float exactly: "{:g}" keeps 6 significant digits, which moved 101.49996 to
            # "101.5" (bin 102 instead of 101) and truncated intensities above 1e6
            f.write(f"{float(mz)!r} {float(inten)!r}\n")


def preprocess(msp_file, mol_dir, data_dir, verbose=True):
    """`--mode preprocess`: pair every MSP record with `<mol_dir>/<key>.mol` (case-insensitive
    stem and extension) and write `data_dir/mol_files/<key>.mol` + `data_dir/msp_files/<key>.msp`.
    Records without peaks or without a MOL file are skipped and counted.  Returns the stats."""
    recs = parse_msp(msp_file)
    stems = {}
    for fn in os.listdir(mol_dir):
        stem, ext = os.path.splitext(fn)
        if ext.lower() in (".mol", ".sdf"):
            stems.setdefault(stem.lower(), fn)
    mol_out, msp_out = os.path.join(data_dir, "mol_files"), os.path.join(data_dir, "msp_files")
    os.makedirs(mol_out, exist_ok=True)
    os.makedirs(msp_out, exist_ok=True)
    stats = {"records": len(recs), "written": 0, "no_peaks": 0, "no_mol": 0}
    for i, rec in enumerate(recs):
        if not rec["peaks"]:
            stats["no_peaks"] += 1
            continue
        hit = next((c for c in record_key(rec, i) if c.lower() in stems), None)
        if hit is None:
            stats["no_mol"] += 1
            continue
        stem = os.path.splitext(stems[hit.lower()])[0]
        shutil.copyfile(os.path.join(mol_dir, stems[hit.lower()]), os.path.join(mol_out, stem + ".mol"))
        write_reference_msp(os.path.join(msp_out, stem + ".msp"), rec)
        stats["written"] += 1
    if verbose:
        print(f"Preprocessed {stats['written']} of {stats['records']} records into {data_dir} "
              f"({stats['no_mol']} without a MOL file, {stats['no_peaks']} without peaks)")
    return stats


# ------------------------------------------------------------------------------ packed dataset cache
_CACHE_VERSION = 1


def save_packed(path, table: MolTable, peak_ptr, peak_mz, peak_inten, names=None):
    """The dataset as the flat arrays the device holds: node_ptr / bond_ptr (int64), feat (f32),
    bond_begin / bond_end (int32), peak_ptr (int64), peak_mz (f64), peak_inten (f32)."""
    np.savez(path, version=np.int64(_CACHE_VERSION), node_ptr=table.node_ptr.astype(np.int64),
             bond_ptr=table.bond_ptr.astype(np.int64), feat=np.ascontiguousarray(table.feat, np.float32),
             bond_begin=table.bond_begin.astype(np.int32), bond_end=table.bond_end.astype(np.int32),
             peak_ptr=np.asarray(peak_ptr, np.int64), peak_mz=np.asarray(peak_mz, np.float64),
             peak_inten=np.asarray(peak_inten, np.float32),
             names=np.asarray(names if names is not None else [], dtype=str))


def load_packed(path):
    """-> (MolTable, (peak_ptr, peak_mz, peak_inten), names)"""
    with np.load(path, allow_pickle=False) as z:
        if int(z["version"]) != _CACHE_VERSION:
            raise ValueError(f"{path}: unsupported dataset cache version {int(z['version'])}")
        table = MolTable(z["node_ptr"], z["bond_ptr"], z["feat"], z["bond_begin"], z["bond_end"])
        return table, (z["peak_ptr"], z["peak_mz"], z["peak_inten"]), [str(n) for n in z["names"]]


def pack_graphs(graphs):
    """MolTable of a list of MolGraph objects (script.mol_to_dgl_graph results)."""
    n = np.array([g.num_nodes() for g in graphs], np.int64)
    b = np.array([len(g._bb) for g in graphs], np.int64)
    node_ptr, bond_ptr = np.zeros(len(graphs) + 1, np.int64), np.zeros(len(graphs) + 1, np.int64)
    np.cumsum(n, out=node_ptr[1:])
    np.cumsum(b, out=bond_ptr[1:])
    feat = np.concatenate([g._feat for g in graphs]) if graphs else np.zeros((0, 6), np.float32)
    bb = np.concatenate([g._bb for g in graphs]) if graphs else np.zeros(0, np.int32)
    be = np.concatenate([g._be for g in graphs]) if graphs else np.zeros(0, np.int32)
    return MolTable(node_ptr, bond_ptr, feat.astype(np.float32), bb.astype(np.int32), be.astype(np.int32))
