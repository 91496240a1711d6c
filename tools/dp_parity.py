#!/usr/bin/env python
"""Two-or-more-rank parity of the fused data-parallel step (csrc/dp_fused.cu: all-reduce + AdamW + parameter
broadcast in one kernel over NVLink peer memory) against the shard-sequential ORACLE:

    torchrun --nproc-per-node 2 tools/dp_parity.py [--branch multicast|peer] [--graph] [--out file.json]

Every rank runs `steps` optimiser steps (dropout 0) on its shard through the product path; rank 0 replays the same
steps with `oracle.Trainer.step(world_shards=[shard of rank 0, shard of rank 1, ...])` - per-shard BatchNorm
statistics, gradients averaged, torch AdamW + OneCycleLR - and compares the parameters on the elements whose
gradient is above fp32 noise (where BatchNorm makes the loss invariant to a parameter the true gradient is 0 and
Adam turns rounding noise into lr-sized steps on both sides; same rule as tests/test_gpu_e2e.py) at 1e-4.
Parameters must be bit-identical across ranks.  Test infrastructure: imports oracle/."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "computational-chemistry-ai_b200"))
from eims_b200.dist import FusedP2PAdamW, broadcast_params, shard_epoch, train_step_fused  # noqa: E402
from eims_b200.engine import DeviceDataset, FlatParams, GraphedTrainStep, ModelDims, Plan, make_step, onecycle_schedule  # noqa: E402
from eims_b200.synth import dense_spectra, synth_molecules, synth_peaks  # noqa: E402
from oracle import gcn_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--branch", default="multicast", choices=["multicast", "peer"])
    ap.add_argument("--graph", action="store_true", help="replay the steps from captured CUDA graphs")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    d = ModelDims(6, 128, 3, 200, "combined", 0.0)
    od = O.Dims(6, 128, 3, 200, "combined", 0.0)
    n_mols, batch, steps = 2048, 48, a.steps
    table = synth_molecules(n_mols, max_atoms=32, seed=3)
    spectra = dense_spectra(*synth_peaks(n_mols, d.max_mz, seed=4), d.max_mz)
    ds = DeviceDataset(table, spectra, dev)
    ids_h = shard_epoch(n_mols, world, rank, batch, epoch=0, seed=1)
    ids = torch.from_numpy(ids_h).to(dev)
    sched = onecycle_schedule(steps)
    plan = Plan(d, batch, batch * 32, 2 * (batch * 32 + 3 * batch), dev)
    fp = FlatParams(d, dev)
    sd0 = O.init_params(od, 0)
    if rank == 0:
        fp.load_state_dict(sd0)
    broadcast_params(fp)
    fused = FusedP2PAdamW(fp, d.num_gcn_layers, overlap=False, use_multicast=(a.branch == "multicast"))
    if a.branch == "multicast" and not fused.multicast:
        print("note: no multicast address on this box; the peer-load branch ran instead", flush=True)
    metrics = torch.zeros(8, device=dev)
    mk = lambda k: make_step(lr=sched[k][0], beta1=sched[k][1], grad_scale=1.0 / world, step=k + 1, seed=5)
    if a.graph:
        gs = GraphedTrainStep(plan, ds, fp, batch, metrics, fused=fused)
        gs.capture(ids[0], mk(0))
        for k in range(steps):
            gs.step(mk(k), ids[min(k + 1, steps - 1)])
    else:
        for k in range(steps):
            train_step_fused(plan, ds, ids[k], fp, mk(k), fused, metrics, next_ids=ids[k + 1] if k + 1 < steps else None)
    torch.cuda.synchronize()
    plan.check()
    lost = fused.lost_peer()
    p = fp.params.clone()
    gathered = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(gathered, p)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    ok, report = True, None
    if rank == 0:
        tr = O.Trainer(sd0, od, total_steps=steps)
        reliable = None
        for k in range(steps):
            shards = []
            for r in range(world):
                idl = shard_epoch(n_mols, world, r, batch, epoch=0, seed=1)[k]
                graph, feat = O.Graph.from_mols([table.mol(int(i)) for i in idl])
                shards.append((graph, feat, torch.from_numpy(spectra[idl])))
            tr.step(None, None, None, world_shards=shards)
            og = {n: tr.sd[n].grad.abs() for n in tr.names}
            okm = {n: (g > 1e-3 * g.max()) for n, g in og.items()}
            reliable = okm if reliable is None else {n: reliable[n] & okm[n] for n in okm}
        got = {n: t.cpu() for n, t in fp.named_params().items()}
        worst, frac = {}, {}
        for n in tr.names:
            ref = tr.sd[n].detach()
            m = reliable[n]
            frac[n] = float(m.float().mean())
            if m.any():
                worst[n] = float((got[n] - ref).abs()[m].max() / ref.abs().max().clamp_min(1e-30))
        report = {"world": world, "branch": "multicast" if fused.multicast else "peer", "graph": bool(a.graph), "steps": steps,
                  "ranks_bit_identical": bool(same), "lost_peer": lost, "worst_rel_err": max(worst.values()),
                  "per_tensor": worst, "reliable_fraction_mean": float(np.mean(list(frac.values()))), "last_loss": float(metrics[4])}
        ok = same and lost == 0 and report["worst_rel_err"] < 1e-4 and report["reliable_fraction_mean"] > 0.3
        print(json.dumps(report), flush=True)
        if a.out:
            with open(a.out, "w") as fh:
                json.dump(report, fh, indent=1)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if not int(flag.item()):
        raise SystemExit("DP PARITY FAILED")
    if rank == 0:
        print("DP PARITY OK", flush=True)


if __name__ == "__main__":
    main()
