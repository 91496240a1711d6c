"""World-size-2 data-parallel logic on the CPU (gloo): sharding, flat-buffer bucketing,
gradient averaging and weight broadcast of eims_b200.dist, checked against the oracle running
the two shards sequentially (DDP semantics).  No CUDA kernel runs here; the
NCCL path with the kernels is exercised by bench.py under torchrun on the GPU box."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "computational-chemistry-ai_b200"))

from eims_b200.dist import GradReducer, broadcast_params, head_split, shard_epoch, stratified_epoch  # noqa: E402
from eims_b200.engine import FlatParams, ModelDims, param_offsets, param_spec  # noqa: E402
from eims_b200.synth import dense_spectra, synth_molecules, synth_peaks  # noqa: E402
from oracle import gcn_oracle as O  # noqa: E402

D = ModelDims(hidden_dim=64, max_mz=100, dropout=0.0)
OD = O.Dims(6, 64, 3, 100, "combined", 0.0)
N_MOLS, BATCH, WORLD = 64, 8, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard_inputs(table, spectra, ids):
    graph, feat = O.Graph.from_mols([table.mol(int(i)) for i in ids])
    return graph, feat, torch.from_numpy(spectra[np.asarray(ids)])


def _flat(grads, d):
    spec, off = param_spec(d), param_offsets(d)
    flat = torch.zeros(off[-1], dtype=torch.float32)
    for (name, shape), o in zip(spec, off):
        flat[o:o + int(np.prod(shape))] = grads[name].reshape(-1)
    return flat


def _worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        torch.set_num_threads(1)
        table = synth_molecules(N_MOLS, max_atoms=12, seed=5)
        spectra = dense_spectra(*synth_peaks(N_MOLS, 100, seed=6), 100)
        # identical weights everywhere: rank 1 starts from garbage and must receive rank 0's
        fp = FlatParams(D, "cpu")
        sd = O.init_params(OD, 0)
        if rank == 0:
            fp.load_state_dict(sd)
        else:
            fp.params.fill_(123.0)
            fp.bn_running.fill_(7.0)
        broadcast_params(fp)
        ref = FlatParams(D, "cpu")
        ref.load_state_dict(sd)
        assert torch.equal(fp.params, ref.params) and torch.equal(fp.bn_running, ref.bn_running)

        ids = shard_epoch(N_MOLS, WORLD, rank, BATCH, epoch=3, seed=11)
        graph, feat, tgt = _shard_inputs(table, spectra, ids[0])
        _, _, grads, _ = O.loss_and_grads(sd, graph, feat, tgt, OD)
        flat = _flat(grads, D)
        offsets = param_offsets(D)
        red = GradReducer(offsets, D.num_gcn_layers)
        assert red.world == WORLD and red.split == head_split(offsets, D.num_gcn_layers)
        red.head_ready(flat)       # no-op on CPU tensors: the one-shot path must still be complete
        red.finish(flat)
        mean = flat / WORLD        # eims_step.grad_scale = 1/world is applied by the AdamW kernel
        if rank == 0:
            np.save(out, mean.numpy())
    finally:
        dist.destroy_process_group()


def test_shard_epoch_partitions_the_permutation():
    shards = [shard_epoch(1000, 4, r, 32, epoch=2, seed=9) for r in range(4)]
    assert all(s.shape == (1000 // (4 * 32), 32) and s.dtype == np.int32 for s in shards)
    flat = np.concatenate([s.reshape(-1) for s in shards])
    assert len(np.unique(flat)) == len(flat)                      # disjoint
    other = shard_epoch(1000, 4, 0, 32, epoch=3, seed=9)
    assert not np.array_equal(other, shards[0])                   # reshuffled every epoch
    assert np.array_equal(shard_epoch(1000, 4, 0, 32, epoch=2, seed=9), shards[0])  # deterministic
    perm = np.random.Generator(np.random.PCG64([9, 2])).permutation(1000)
    assert np.array_equal(shards[1].reshape(-1), perm[1::4][: shards[1].size])
    with pytest.raises(ValueError):
        shard_epoch(100, 4, 0, 32, epoch=0)
    with pytest.raises(ValueError):
        shard_epoch(1000, 4, 4, 32, epoch=0)


def test_stratified_epoch_equalises_batch_work():
    table = synth_molecules(4096 + 37, max_atoms=64, seed=8)
    sizes = np.diff(table.node_ptr)
    ids = stratified_epoch(sizes, 128, epoch=1, seed=3)
    assert ids.shape == (4133 // 128, 128) and ids.dtype == np.int32
    flat = ids.reshape(-1)
    assert len(np.unique(flat)) == len(flat) and flat.min() >= 0 and flat.max() < len(sizes)   # each molecule at most once
    work = sizes[ids].sum(axis=1)
    rnd = sizes[np.random.default_rng(0).permutation(len(sizes))[: ids.size]].reshape(ids.shape).sum(axis=1)
    assert work.std() < 0.2 * rnd.std()          # far tighter than uniform shuffling
    assert abs(work.mean() - rnd.mean()) < 0.01 * rnd.mean()
    assert not np.array_equal(ids, stratified_epoch(sizes, 128, epoch=2, seed=3))
    assert np.array_equal(ids, stratified_epoch(sizes, 128, epoch=1, seed=3))


def test_head_bucket_is_the_tail_of_the_flat_buffer():
    off, spec = param_offsets(D), param_spec(D)
    split = head_split(off, D.num_gcn_layers)
    names_tail = [n for (n, _), o in zip(spec, off) if o >= split]
    assert names_tail and all(n.startswith("spectrum_predictor") for n in names_tail)
    assert all(not n.startswith("spectrum_predictor") for (n, _), o in zip(spec, off) if o < split)


@pytest.mark.timeout(300)
def test_two_rank_gradient_average_equals_sequential_oracle(tmp_path):
    out = str(tmp_path / "mean.npy")
    mp.spawn(_worker, args=(_free_port(), out), nprocs=WORLD, join=True)
    got = np.load(out)
    # the oracle: both shards on one process, gradients averaged (Trainer.step(world_shards=...) semantics)
    table = synth_molecules(N_MOLS, max_atoms=12, seed=5)
    spectra = dense_spectra(*synth_peaks(N_MOLS, 100, seed=6), 100)
    sd = O.init_params(OD, 0)
    acc = None
    for r in range(WORLD):
        ids = shard_epoch(N_MOLS, WORLD, r, BATCH, epoch=3, seed=11)
        graph, feat, tgt = _shard_inputs(table, spectra, ids[0])
        _, _, grads, _ = O.loss_and_grads(sd, graph, feat, tgt, OD)
        f = _flat(grads, D).numpy().astype(np.float64)
        acc = f if acc is None else acc + f
    ref = (acc / WORLD).astype(np.float32)
    # fp32 round-off only (the in-process oracle may use a different torch thread count)
    assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1e-30)


def test_global_batches_keep_the_uniform_global_batch_and_balance_atoms():
    from eims_b200.dist import global_batches
    table = synth_molecules(8 * 128 * 6 + 91, max_atoms=64, seed=8)
    sizes = np.diff(table.node_ptr)
    W, B = 8, 128
    bal = [global_batches(sizes, W, r, B, epoch=4, seed=2) for r in range(W)]
    ddp = [global_batches(sizes, W, r, B, epoch=4, seed=2, balance=False) for r in range(W)]
    perm = np.random.Generator(np.random.PCG64([2, 4])).permutation(len(sizes))
    assert all(b.shape == (6, B) and b.dtype == np.int32 for b in bal + ddp)
    for k in range(6):
        chunk = np.sort(perm[k * W * B:(k + 1) * W * B])
        # the SET of molecules in the global batch is the reference's uniform draw, whatever the assignment
        assert np.array_equal(np.sort(np.concatenate([b[k] for b in bal])), chunk)
        assert np.array_equal(np.sort(np.concatenate([b[k] for b in ddp])), chunk)
        assert np.array_equal(ddp[3][k], perm[k * W * B:(k + 1) * W * B][3::W])      # DistributedSampler order
        wb = np.array([sizes[b[k]].sum() for b in bal])
        wd = np.array([sizes[b[k]].sum() for b in ddp])
        assert wb.max() - wb.min() <= 0.003 * wb.mean()                              # atoms per rank within 0.3 %
        assert wd.max() - wd.min() > 5 * (wb.max() - wb.min())
    assert np.array_equal(global_batches(sizes, 1, 0, B, 0, 2), global_batches(sizes, 1, 0, B, 0, 2, balance=False))
    assert global_batches(sizes, W, 0, B, 0, 2, steps=2).shape == (2, B)
    with pytest.raises(ValueError):
        global_batches(sizes[:100], W, 0, B, 0)
