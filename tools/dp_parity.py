#!/usr/bin/env python
"""Two-or-more-rank parity of the fused data-parallel step (csrc/dp_fused.cu: all-reduce + AdamW + parameter
broadcast in one kernel over NVLink peer memory) against the shard-sequential ORACLE:

    torchrun --nproc-per-node 2 tools/dp_parity.py [--branch multicast|peer] [--graph] [--out file.json]

Part A - the exchange kernel itself, exact.  Every rank fills its gradient buffer with a seeded pseudo-random
vector (magnitudes over 8 decades) for 6 steps and runs the fused kernel; the expectation is NCCL all-reduce(sum) of
the same vectors followed by `eims_adamw_flat` (grad_scale 1/world) on the whole vector.  With two ranks a sum has
one order, so parameters and the Adam state of the own slice must match BIT FOR BIT (with more ranks the switch's
summation order is its own: 1e-6).  Covers the double-buffered gradients, their zeroing, both barriers, the broadcast.

Part B - the training trajectory against the ORACLE.  Every rank runs `steps` optimiser steps (dropout 0) on its
shard through the product path; rank 0 replays them with `oracle.Trainer.step(world_shards=[...])` - per-shard
BatchNorm statistics, gradients averaged, torch AdamW + OneCycleLR.  Per step the losses must agree to 1e-4.  After
the last step every parameter element whose gradient was above fp32 noise in every step (where the loss does not
depend on an element the true gradient is 0 and Adam turns rounding noise into lr-sized steps on BOTH sides; same
rule as tests/test_gpu_e2e.py) must be within 1e-4 of the oracle's relative to the tensor's scale, or within 5 % of
the distance any parameter can have travelled (the sum of the learning rates; biases start at zero and are all
"distance travelled").  The second clause is a sanity bound, not a parity claim: over several steps a ReLU
pre-activation within rounding distance of 0 falls on different sides in the two implementations,
which moves early-layer gradients by ~1e-3 of their maximum, and AdamW's normalised update turns that into a
percent-level change of the step of the weaker elements (measured at 2 ranks: 1.5 % of the distance on one element of
gcn_layers.2.weight, every head tensor within 1e-5).  Element-wise parity of the arithmetic is pinned elsewhere:
single-step gradients agree with the fp64 oracle to 2e-6 once the discrete decisions are shared
(tests/test_gpu_e2e.py::test_full_size_flip_aware_gradients), the AdamW kernel equals torch's (test_gpu_kernels.py),
and part A above pins the exchange itself bit for bit.  The loss, which sees every parameter, must track the oracle's
to 1e-4 at every step (measured: 1.4e-7).  Parameters must be bit-identical across ranks.  Test infrastructure: imports oracle/."""
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
    ap.add_argument("--overlap", action="store_true", help="two buckets: the head's exchange on a side stream under the GraphConv backward")
    ap.add_argument("--group", type=int, default=0, help="--graph: steps per graph (0 = single-step graphs, losses checked per step)")
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
    fused = FusedP2PAdamW(fp, d.num_gcn_layers, overlap=a.overlap, use_multicast=(a.branch == "multicast"))
    if a.branch == "multicast" and not fused.multicast:
        print("note: no multicast address on this box; the peer-load branch ran instead", flush=True)
    metrics = torch.zeros(8, device=dev)
    mk = lambda k: make_step(lr=sched[k][0], beta1=sched[k][1], grad_scale=1.0 / world, step=k + 1, seed=5)
    exact = kernel_exactness(fused, fp, rank, world, dev)
    # back to the common initial state for part B
    fp.load_state_dict(sd0)
    broadcast_params(fp)
    for t in fused.m + fused.v:
        t.zero_()
    for g in fused.sym_grads:
        g.zero_()
    torch.cuda.synchronize()
    dist.barrier()
    losses = []
    if a.graph:
        gs = GraphedTrainStep(plan, ds, fp, batch, metrics, fused=fused, group=a.group)
        gs.capture(ids[0], mk(0))
        for k in range(steps):
            gs.step(mk(k), ids[min(k + 1, steps - 1)])
            if not a.group:
                gs.flush()
                losses.append(float(metrics[4]))
        gs.flush()
        if a.group:
            losses = [None] * (steps - 1) + [float(metrics[4])]
    else:
        for k in range(steps):
            train_step_fused(plan, ds, ids[k], fp, mk(k), fused, metrics, next_ids=ids[k + 1] if k + 1 < steps else None)
            losses.append(float(metrics[4]))
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
        olosses, reliable = [], None
        for k in range(steps):
            shards = []
            for r in range(world):
                idl = shard_epoch(n_mols, world, r, batch, epoch=0, seed=1)[k]
                graph, feat = O.Graph.from_mols([table.mol(int(i)) for i in idl])
                shards.append((graph, feat, torch.from_numpy(spectra[idl])))
            _, ls = tr.step(None, None, None, world_shards=shards)
            olosses.append(ls[0])   # rank 0's shard
            og = {n: tr.sd[n].grad.abs() for n in tr.names}
            okm = {n: (g > 1e-3 * g.max()) for n, g in og.items()}
            reliable = okm if reliable is None else {n: reliable[n] & okm[n] for n in okm}
        got = {n: t.cpu() for n, t in fp.named_params().items()}
        travelled = float(sum(lr for lr, _ in sched))
        worst, bound, frac = {}, {}, {}
        for n in tr.names:
            ref = tr.sd[n].detach()
            m = reliable[n]
            frac[n] = float(m.float().mean())
            worst[n] = float((got[n] - ref).abs()[m].max()) if m.any() else 0.0
            bound[n] = max(1e-4 * float(ref.abs().max()), 5e-2 * travelled)
        loss_err = max(abs(x - y) / abs(y) for x, y in zip(losses, olosses) if x is not None)
        report = {"world": world, "branch": "multicast" if fused.multicast else "peer", "graph": bool(a.graph), "steps": steps,
                  "kernel_vs_nccl_plus_adamw": exact, "ranks_bit_identical": bool(same), "lost_peer": lost,
                  "loss_rel_err_max": loss_err, "losses": losses, "oracle_losses": olosses,
                  "param_abs_err_over_bound_max": max(worst[n] / bound[n] for n in worst),
                  "param_abs_err": worst, "param_bound": bound, "distance_travelled_sum_lr": travelled,
                  "reliable_fraction_mean": float(np.mean(list(frac.values())))}
        ok = (same and lost == 0 and exact["ok"] and loss_err < 1e-4 and report["param_abs_err_over_bound_max"] < 1.0
              and report["reliable_fraction_mean"] > 0.3)
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


def kernel_exactness(fused, fp, rank, world, dev):
    """Part A: the fused kernel on synthetic gradients against NCCL all-reduce + eims_adamw_flat."""
    import ctypes as C
    from eims_b200 import _lib
    from eims_b200._lib import check, ptr
    lib = _lib.load()
    n = fused.n_pad
    ref_p = fused.sym_params.clone()
    ref_m, ref_v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    sched = onecycle_schedule(6)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    per = n // world
    for k in range(6):
        g = torch.Generator(device="cpu").manual_seed(1000 * k + rank)
        grad = (torch.randn(n, generator=g) * torch.pow(10.0, torch.randint(-6, 3, (n,), generator=g).float())).to(dev)
        step = make_step(lr=sched[k][0], beta1=sched[k][1], grad_scale=1.0 / world, step=k + 1)
        fused.begin_step()
        fused.sym_grads[(fused.seq - 1) % 2].copy_(grad)
        fused.finish(step, st)
        total = grad.clone()
        dist.all_reduce(total)
        check(lib.eims_adamw_flat(ptr(ref_p), ptr(total), ptr(ref_m), ptr(ref_v), n, C.byref(step), st))
    torch.cuda.synchronize()
    sl = slice(rank * per, (rank + 1) * per)
    if world == 2:
        ok = bool(torch.equal(fused.sym_params, ref_p))
        if len(fused.buckets) == 1:   # (two buckets: every bucket keeps the Adam state of its own slice; the parameters say it all)
            ok = ok and bool(torch.equal(fused.m[0], ref_m[sl]) and torch.equal(fused.v[0], ref_v[sl]))
        err = float((fused.sym_params - ref_p).abs().max())
    else:
        err = float((fused.sym_params - ref_p).abs().max() / ref_p.abs().max())
        ok = err < 1e-6
    zeroed = bool(not fused.sym_grads[fused.seq % 2].any())   # the buffer of the NEXT step has been zeroed
    flag = torch.tensor([1.0 if (ok and zeroed) else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"ok": bool(flag.item() > 0), "bitwise": world == 2, "max_abs_err": err, "steps": 6}


if __name__ == "__main__":
    main()
