"""CPU-only: the C-ABI library loads and exports every symbol include/eims_b200.h declares;
pure-host entry points (parameter layout, argument validation) behave."""
import ctypes as C
import os
import re

import pytest

from eims_b200 import _lib
from eims_b200.engine import ModelDims, param_offsets, param_spec, state_dict_order

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "eims_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eims_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/eims_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)
    assert lib.eims_version() == 100


def test_param_layout_matches_reference_state_dict():
    # parameter counts from SURVEY Appendix A.6 (torch-verified for the reference module)
    assert param_offsets(ModelDims(max_mz=500))[-1] == 658_932
    assert param_offsets(ModelDims(max_mz=1000))[-1] == 787_432
    assert param_offsets(ModelDims(hidden_dim=1024, num_gcn_layers=6, max_mz=1000))[-1] == 12_593_128
    d = ModelDims(max_mz=1000)
    off, spec = param_offsets(d), param_spec(d)
    assert len(spec) == 22 and len(off) == 23
    for (name, shape), a, b in zip(spec, off[:-1], off[1:]):
        n = 1
        for s in shape:
            n *= s
        assert b - a == n, name
    order = state_dict_order(d)
    assert order[0] == "gcn_layers.0.weight" and order[-1] == "spectrum_predictor.8.bias" and len(order) == 22 + 9
    assert order.index("batch_norms.0.num_batches_tracked") == 10
    # single pooling => first Linear is (2H, H)  (GCN:325-342)
    assert dict(param_spec(ModelDims(pooling="max")))["spectrum_predictor.0.weight"] == (512, 256)


@pytest.mark.parametrize("bad", [dict(hidden_dim=100), dict(max_mz=1001), dict(num_gcn_layers=0), dict(node_feat_dim=9),
                                 dict(dropout=1.0)])
def test_bad_dims_are_rejected_with_a_message(bad):
    lib = _lib.load()
    cd = ModelDims(**bad).c()
    assert lib.eims_param_count(C.byref(cd)) == -1
    assert len(lib.eims_last_error()) > 0
    h = C.c_void_p()
    assert lib.eims_plan_create(C.byref(cd), 4, 64, 64, C.byref(h)) == _lib.ERR_ARG


def test_plan_sizes_without_a_device():
    lib = _lib.load()
    cd = ModelDims(max_mz=1000).c()
    h = C.c_void_p()
    assert lib.eims_plan_create(C.byref(cd), 512, 512 * 40, 512 * 90, C.byref(h)) == 0
    nbytes = lib.eims_plan_workspace_bytes(h)
    assert 100e6 < nbytes < 400e6  # ~8 activations of N*H*4 bytes + head buffers
    assert lib.eims_plan_buffer(h, b"rowptr", None, None) == _lib.ERR_STATE  # not bound yet
    assert lib.eims_plan_destroy(h) == 0
