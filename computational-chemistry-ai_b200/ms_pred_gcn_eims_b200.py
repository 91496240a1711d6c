#!/usr/bin/env python
"""Drop-in for turnDeep/Computational-Chemistry-AI `templates/ms-pred-gcn-eims-cupy.py`:
same entry points, CLI flags, checkpoint format and no-argument demo banner (GCN:517-633),
with the GPU work done by hand-written sm_100a kernels behind include/eims_b200.h.

    python ms_pred_gcn_eims_b200.py --mode train --data_dir processed_data
    python ms_pred_gcn_eims_b200.py --mode predict --smiles 'CC(C)CC1=CC=C(C=C1)C(C)C'
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from eims_b200.script import (Config, CuPySpectrumProcessor, GCNSpectrum, OptimizedEIMSDataset,  # noqa: E402,F401
                              collate_fn, demo_banner, get_atom_features, main, mol_to_dgl_graph,
                              predict_spectrum, train_model)

if __name__ == "__main__":
    if len(sys.argv) == 1:
        demo_banner()
    else:
        main()
