"""Golden vectors for the float32 ("CuPy") branch of CuPySpectrumProcessor.peaks_to_spectrum_batch
(GCN:170-191), produced by the UNMODIFIED reference script with its array module `cp` bound to
NumPy - the alias the script itself sets up when CuPy is missing (GCN:59) - plus the one
function NumPy lacks, `asnumpy` (identity).  Build container only:

    python tests/golden/make_golden_binning_f32.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "computational-chemistry-ai_b200"))

from eims_b200.synth import peaks_as_lists, synth_peaks  # noqa: E402
from oracle import dgl_shim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def fractional_peaks(n, max_mz, seed):
    """Peak lists with fractional, half-integer, negative and out-of-range m/z and duplicates."""
    rng = np.random.Generator(np.random.PCG64(seed))
    k = rng.integers(0, 60, size=n)
    ptr = np.zeros(n + 1, np.int64)
    np.cumsum(k, out=ptr[1:])
    tot = int(ptr[-1])
    mz = rng.uniform(-3.0, max_mz + 3.0, size=tot)
    half = rng.random(tot) < 0.3
    mz[half] = np.floor(mz[half]) + 0.5           # exact ties -> round-half-even matters
    near = rng.random(tot) < 0.1
    mz[near] = np.floor(mz[near]) + 0.49999999    # differs between float32 and float64 rounding
    inten = rng.uniform(-5.0, 999.0, size=tot)
    return ptr, mz, inten


def main():
    ref = dgl_shim.load_reference()
    npx = types.ModuleType("numpy_as_cupy")
    npx.__dict__.update(np.__dict__)
    npx.asnumpy = lambda a: a
    ref.cp = npx
    ref.CUPY_AVAILABLE = True
    out = {}
    for name, (n, M, seed) in {"a": (16, 100, 3), "b": (40, 1000, 4)}.items():
        ptr, mz, inten = fractional_peaks(n, M, seed)
        spec = ref.CuPySpectrumProcessor(M, True).peaks_to_spectrum_batch(peaks_as_lists(ptr, mz, inten))
        spec64 = ref.CuPySpectrumProcessor(M, False).peaks_to_spectrum_batch(peaks_as_lists(ptr, mz, inten))
        out.update({f"{name}_ptr": ptr, f"{name}_mz": mz, f"{name}_inten": inten, f"{name}_spec_f32": np.asarray(spec, np.float32),
                    f"{name}_spec_f64": np.asarray(spec64, np.float32), f"{name}_max_mz": np.int64(M)})
        print(name, "rows where the two branches differ:", int((np.asarray(spec) != np.asarray(spec64)).any(axis=1).sum()))
    pk = synth_peaks(8, 100, seed=5)
    out["c_spec_f32"] = np.asarray(ref.CuPySpectrumProcessor(100, True).peaks_to_spectrum_batch(peaks_as_lists(*pk)), np.float32)
    np.savez_compressed(os.path.join(HERE, "binning_f32.npz"), **out)
    print("binning_f32.npz", os.path.getsize(os.path.join(HERE, "binning_f32.npz")))


if __name__ == "__main__":
    main()
