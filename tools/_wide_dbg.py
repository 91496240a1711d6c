import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/computational-chemistry-ai_b200'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from test_gpu_e2e import *
d = ModelDims(hidden_dim=1024, num_gcn_layers=6, max_mz=1000, dropout=0.0)
for backend in ("tcgen05", "simt"):
    table, targets, plan, ds, fp, sd = setup(d, 24, 128, 77, backend, wseed=5)
    prob, loss, cos, grads = gpu_fwd_bwd(plan, ds, fp, None, make_step())
    if backend == "tcgen05":
        graph, feat = O.Graph.from_mols([table.mol(i) for i in range(24)])
        tt = torch.from_numpy(targets)
        p64, l64, g64, _ = O.loss_and_grads(sd, graph, feat, tt, odims(d), dtype=torch.float64)
        p32, l32, g32, _ = O.loss_and_grads(sd, graph, feat, tt, odims(d))
    print(backend, "prob", rel_err(prob, p64.numpy()), "fp32 oracle", rel_err(p32.numpy(), p64.numpy()))
    for n in grads:
        print(f"  {n:32s} gpu {rel_err(grads[n], g64[n].numpy()):.2e}   fp32-oracle {rel_err(g32[n].numpy(), g64[n].numpy()):.2e}")
