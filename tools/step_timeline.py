#!/usr/bin/env python
"""In-situ duration of every launch of a training step (diagnostic, on a B200).

Builds the -DEIMS_TIMELINE variant of the library (block 0 of every kernel stamps %globaltimer right
after its grid-dependency wait, i.e. when the previous kernel of the chain has completed), runs
BASELINE configs[1] steps on one stream with programmatic dependent launch left ON and prints the
median gap between consecutive stamps, labelled with the launch sequence of plan.cu.
    python tools/step_timeline.py [--steps 20] [--hidden 256 --layers 3 --batch 512 --max-atoms 64]
"""
import argparse
import ctypes as C
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "computational-chemistry-ai_b200"))
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "computational-chemistry-ai_b200", "build.py"))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)


def sequence(L, H=256):
    s = ["k1_build", "layer0_fwd"]
    for l in range(1, L):
        s += [f"spmm_fwd{l}", f"gemm_gcn_fwd{l}"]
    s += ["readout", "gemm_head1", "ln1", "gemm_head2", "ln2", "gemm_head3", "loss",
          "colsum", "gemm_bwd_h3", "ln_bwd2", "gemm_bwd_h2", "ln_bwd1", "gemm_bwd_h1"]
    for l in range(L - 1, -1, -1):
        if l == L - 1 or H > 256:  # below the top layer the statistics ride on the SpMM (H <= 256 only)
            s += [f"bn_bwd_stats{l}"]
        s += [f"bn_bwd_apply{l}"]
        if l > 0:
            s += [f"gemm_bwd_gcn{l}", f"spmm_bwd{l}"]
    s += ["adamw"]
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--max-atoms", type=int, default=64)
    ap.add_argument("--no-build", action="store_true")
    a = ap.parse_args()
    path = b.OUT.replace(".so", "_timeline.so") if a.no_build else b.build_variant("timeline", ["-DEIMS_TIMELINE"])
    from eims_b200 import _lib
    _lib.LIB_PATH = path
    import torch
    from bench import init_weights
    from eims_b200.engine import DeviceDataset, FlatParams, ModelDims, Plan, make_step, onecycle_schedule
    from eims_b200.synth import dense_spectra, synth_molecules, synth_peaks

    lib = _lib.load()
    readers = [getattr(lib, f"eims_debug_timeline_read_{tu}") for tu in ("graph", "dense", "gemm_tc", "dp_fused", "plan")]
    for r in readers:
        r.restype, r.argtypes = C.c_int, [C.c_void_p, C.c_int32]

    def read_all():
        out = []
        buf = (C.c_ulonglong * 8192)()
        for r in readers:
            n = r(buf, 8192)
            assert n >= 0
            out.extend(buf[:n])
        return np.sort(np.array(out, dtype=np.int64))
    dev = torch.device("cuda", 0)
    M, B = 1000, a.batch
    n_mols = 20000
    table = synth_molecules(n_mols, max_atoms=a.max_atoms, seed=1234)
    targets = dense_spectra(*synth_peaks(n_mols, M, seed=4321), M)
    ds = DeviceDataset(table, targets, dev)
    d = ModelDims(6, a.hidden, a.layers, M, "combined", 0.2)
    plan = Plan(d, B, B * a.max_atoms, 2 * (B * a.max_atoms + 3 * B), dev)
    fp = FlatParams(d, dev)
    init_weights(fp, d)
    perm = torch.from_numpy(np.random.default_rng(0).permutation(n_mols).astype(np.int32)).to(dev)
    sched = onecycle_schedule(1000)
    metrics = torch.zeros(8, device=dev)
    nb = n_mols // B

    def step(k):
        i = k % nb
        plan.train_step(ds, perm[i * B:(i + 1) * B], fp, make_step(lr=sched[k][0], beta1=sched[k][1], step=k + 1, seed=1), metrics)

    for k in range(20):
        step(k)
    torch.cuda.synchronize()
    read_all()  # reset
    for k in range(20, 20 + a.steps):
        step(k)
    t = read_all()
    n = len(t)
    seq = sequence(a.layers, a.hidden)
    per = len(seq)
    assert n == per * a.steps, (n, per, a.steps)
    t = t.reshape(a.steps, per)
    # gap k = stamp of launch k+1 minus stamp of launch k = time launch k held the chain
    flat = t.reshape(-1)
    gaps = np.diff(flat)
    gaps = np.concatenate([gaps, [gaps[-per]]]).reshape(a.steps, per)[1:-1]
    med = np.median(gaps, axis=0) / 1e3
    print(f"step {np.median(np.diff(t[:, 0])) / 1e3:.1f} us (median), {per} launches")
    for nm, g in zip(seq, med):
        print(f"  {nm:16s} {g:7.2f} us")
    groups = {}
    for nm, g in zip(seq, med):
        key = nm.rstrip("0123456789")
        groups[key] = groups.get(key, 0.0) + g
    print("by kernel class:")
    for key, g in sorted(groups.items(), key=lambda kv: -kv[1]):
        print(f"  {key:16s} {g:7.2f} us  {100 * g / med.sum():5.1f} %")


if __name__ == "__main__":
    main()
