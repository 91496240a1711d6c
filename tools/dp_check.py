#!/usr/bin/env python
"""Data-parallel equivalence on real GPUs (run under torchrun, N >= 2): the fused NVLink kernel
(all-reduce + AdamW + broadcast, csrc/dp_fused.cu) against the NCCL all-reduce + AdamW path,
same data, same initial weights, 20 steps.  Every rank must end with the same parameters in
both paths (fp32 summation-order tolerance) and bit-identical parameters across ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "computational-chemistry-ai_b200"))
from bench import init_weights  # noqa: E402
from eims_b200.dist import FusedP2PAdamW, GradReducer, broadcast_params, shard_epoch, train_step_dp, train_step_fused  # noqa: E402
from eims_b200.engine import DeviceDataset, FlatParams, ModelDims, Plan, make_step, onecycle_schedule  # noqa: E402
from eims_b200.synth import dense_spectra, synth_molecules, synth_peaks  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    d = ModelDims(6, 128, 3, 200, "combined", 0.2)
    n_mols, batch, steps = 4096, 64, 20
    table = synth_molecules(n_mols, max_atoms=32, seed=3)
    targets = dense_spectra(*synth_peaks(n_mols, d.max_mz, seed=4), d.max_mz)
    ds = DeviceDataset(table, targets, dev)
    ids = torch.from_numpy(shard_epoch(n_mols, world, rank, batch, epoch=0, seed=1)).to(dev)
    sched = onecycle_schedule(steps)
    results = {}
    for path in ("nccl", "nccl2", "fused"):
        plan = Plan(d, batch, batch * 32, 2 * (batch * 32 + 3 * batch), dev)
        fp = FlatParams(d, dev)
        init_weights(fp, d)
        broadcast_params(fp)
        metrics = torch.zeros(8, device=dev)
        fused = FusedP2PAdamW(fp, d.num_gcn_layers) if path == "fused" else None
        reducer = GradReducer(fp.offsets, d.num_gcn_layers)
        for k in range(steps):
            st = make_step(lr=sched[k][0], beta1=sched[k][1], grad_scale=1.0 / world, step=k + 1, seed=5)
            if fused is not None:
                train_step_fused(plan, ds, ids[k], fp, st, fused, metrics, next_ids=ids[k + 1] if k + 1 < steps else None)
            else:
                train_step_dp(plan, ds, ids[k], fp, st, reducer, metrics)
        torch.cuda.synchronize()
        plan.check()
        p = fp.params.clone()
        gathered = [torch.empty_like(p) for _ in range(world)]
        dist.all_gather(gathered, p)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        results[path] = (p, float(metrics[4]), same, getattr(fused, "multicast", None))
        if rank == 0:
            print(f"{path}: last loss {float(metrics[4]):.6f}, ranks bit-identical: {same}, multicast: {getattr(fused, 'multicast', None)}", flush=True)
    # training is not bit-reproducible (split-K and BatchNorm sums use atomics) and AdamW turns
    # last-bit gradient differences into lr-sized parameter differences, so the yardstick for
    # "same result" is the distance between two runs of the SAME (NCCL) path
    a, a2, b = results["nccl"][0], results["nccl2"][0], results["fused"][0]
    noise = float((a - a2).abs().max() / a.abs().max())
    err = float((a - b).abs().max() / a.abs().max())
    if rank == 0:
        print(f"parameters after {steps} steps: fused vs nccl max rel err {err:.2e}; nccl vs nccl (run-to-run noise) {noise:.2e}", flush=True)
    ok = err < max(4 * noise, 1e-6) and results["fused"][2] and abs(results["nccl"][1] - results["fused"][1]) < 1e-4 * abs(results["nccl"][1])
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit("DP CHECK FAILED")
    if rank == 0:
        print("DP CHECK OK", flush=True)


if __name__ == "__main__":
    main()
