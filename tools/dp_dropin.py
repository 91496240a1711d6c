#!/usr/bin/env python
"""Data-parallel training THROUGH THE DROP-IN (`script.train_model(..., world_size=W)` under torchrun), 2+ GPUs:

    torchrun --nproc-per-node 2 tools/dp_dropin.py [--out file.json]

Every rank builds the same synthetic data set of MolGraphs (what `OptimizedEIMSDataset` holds, GCN:227-258), takes
its shard of every epoch through `script.ShardedLoader`, and calls the reference-shaped `train_model` (GCN:382-488),
which exchanges gradients with the fused NVLink kernel.  Checks: all ranks end with bit-identical parameters; the result equals - within the tolerances of tools/dp_parity.py - the shard-sequential ORACLE driven
with the same shards (dropout 0); rank 0's checkpoint round-trips through `torch.save` / `load_state_dict` in the
reference's format (GCN:589-605).  Test infrastructure: imports oracle/."""
import argparse
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "computational-chemistry-ai_b200"))
from eims_b200 import script as S  # noqa: E402
from eims_b200.dist import shard_epoch  # noqa: E402
from eims_b200.synth import dense_spectra, synth_molecules, synth_peaks  # noqa: E402
from oracle import gcn_oracle as O  # noqa: E402


class ListDataset(torch.utils.data.Dataset):
    def __init__(self, graphs, spectra):
        self.graphs, self.spectra = graphs, spectra

    def __len__(self):
        return len(self.graphs)

    def __getitem__(self, i):
        return self.graphs[i], torch.from_numpy(self.spectra[i])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    world, rank = S._init_data_parallel()
    if world < 2:
        raise SystemExit("run under torchrun with >= 2 ranks")
    dev = S.device
    cfg = S.Config()
    cfg.hidden_dim, cfg.max_mz, cfg.dropout, cfg.batch_size, cfg.num_epochs = 64, 100, 0.0, 16, 2
    n_train, n_val = 16 * world * 3, 32
    table = synth_molecules(n_train + n_val, max_atoms=20, seed=21)
    spectra = dense_spectra(*synth_peaks(n_train + n_val, cfg.max_mz, seed=22), cfg.max_mz)
    graphs = [S.MolGraph(*table.mol(i)) for i in range(n_train + n_val)]
    ds = ListDataset(graphs, spectra)
    train = torch.utils.data.Subset(ds, list(range(n_train)))
    val = torch.utils.data.Subset(ds, list(range(n_train, n_train + n_val)))
    train_loader = S.ShardedLoader(train, cfg.batch_size, world, rank, seed=3)
    val_loader = torch.utils.data.DataLoader(val, batch_size=cfg.batch_size, shuffle=False, collate_fn=S.collate_fn)
    od = O.Dims(6, 64, 3, 100, "combined", 0.0)
    sd0 = O.init_params(od, 0)
    model = S.GCNSpectrum(6, cfg).to(dev)
    if rank == 0:
        model.load_state_dict(sd0)     # the other ranks start from their own random init: train_model must broadcast rank 0's
    model, history = S.train_model(model, train_loader, val_loader, cfg, world_size=world, verbose=False)
    sd = model.state_dict()
    # trainable parameters must be bit-identical everywhere; BatchNorm running buffers are rank-local (DDP without
    # SyncBN: rank 0's are the ones the checkpoint keeps), so they - and the validation numbers - may differ
    flat = model.flat.data.detach().clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    ok, report = True, None
    if rank == 0:
        steps_per_epoch = len(train_loader)
        tr = O.Trainer(sd0, od, total_steps=cfg.num_epochs * steps_per_epoch, lr=cfg.learning_rate, weight_decay=cfg.weight_decay)
        olosses = []
        for ep in range(cfg.num_epochs):
            per_rank = [shard_epoch(n_train, world, r, cfg.batch_size, ep, seed=3) for r in range(world)]
            run = []
            for k in range(steps_per_epoch):
                shards = []
                for r in range(world):
                    idl = per_rank[r][k]
                    graph, feat = O.Graph.from_mols([table.mol(int(i)) for i in idl])
                    shards.append((graph, feat, torch.from_numpy(spectra[idl])))
                _, ls = tr.step(None, None, None, world_shards=shards)
                run.append(ls[0])
            olosses.append(float(np.mean(run)))
        loss_err = max(abs(x - y) / abs(y) for x, y in zip(history["train_loss"], olosses))
        # the oracle's validation loss in the reference's form: mean over val batches of the batch MSE
        vl = []
        for s0 in range(n_train, n_train + n_val, cfg.batch_size):
            g_, f_ = O.Graph.from_mols([table.mol(i) for i in range(s0, s0 + cfg.batch_size)])
            vl.append(float(O.mse_loss(tr.predict(g_, f_), torch.from_numpy(spectra[s0:s0 + cfg.batch_size]))))
        val_err = abs(history["val_loss"][-1] - float(np.mean(vl))) / float(np.mean(vl))
        with tempfile.TemporaryDirectory() as td:   # reference-format checkpoint (GCN:589-593) and back (GCN:599-605)
            path = os.path.join(td, "m.pth")
            torch.save({"model_state_dict": sd, "config": cfg.__dict__, "history": history}, path)
            ck = torch.load(path, weights_only=False)
            m2 = S.GCNSpectrum(6, S.Config(**ck["config"])).to(dev)
            m2.load_state_dict(ck["model_state_dict"])
            rt = all(torch.equal(m2.state_dict()[k].cpu(), sd[k].cpu()) for k in sd)
        report = {"world": world, "ranks_bit_identical": bool(same),
                  "train_loss": history["train_loss"], "oracle_train_loss": olosses, "train_loss_rel_err": loss_err,
                  "val_loss_rel_err_vs_oracle": val_err, "checkpoint_round_trip": bool(rt), "keys": len(sd)}
        ok = same and rt and loss_err < 1e-4 and val_err < 2e-3 and history["train_loss"][-1] < history["train_loss"][0]
        print(json.dumps(report), flush=True)
        if a.out:
            with open(a.out, "w") as fh:
                json.dump(report, fh, indent=1)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if not int(flag.item()):
        raise SystemExit("DP DROP-IN FAILED")
    if rank == 0:
        print("DP DROP-IN OK", flush=True)


if __name__ == "__main__":
    main()
