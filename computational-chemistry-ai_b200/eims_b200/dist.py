"""Data-parallel training of the GCN EI-MS path: one process per GPU, molecules sharded by
rank, one gradient all-reduce per optimiser step.

The reference has no multi-GPU code at all (a single process on `cuda:0`,
/root/reference/templates/ms-pred-gcn-eims-cupy.py:63), so the contract is the one
PyTorch DDP would give the script unchanged:

  * every rank starts from rank 0's weights (`broadcast_params`);
  * rank r trains on molecules perm[r::world] of the epoch's shuffled order, `batch` per rank
    (`shard_epoch`), all ranks taking the same number of steps;
  * gradients are summed over ranks in the flat fp32 buffer (NCCL over NVLink on the GPU,
    gloo in the CPU tests) and the 1/world scaling is applied inside the AdamW kernel
    (`eims_step.grad_scale`), so every rank applies the same update to its own copy;
  * BatchNorm batch statistics and running buffers stay rank-local (no SyncBN); rank 0's
    running buffers are the ones a checkpoint keeps.

Nothing here touches the data path of a step: K1 .. backward run on rank-local molecules with
no collective; the all-reduce sits between backward and AdamW.  `GradReducer` splits the flat
buffer into two buckets - the head (whose gradients are final before the GCN layers are
differentiated) and the GCN / BatchNorm tensors - so that the larger bucket's all-reduce
overlaps the rest of the backward pass on a side stream.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_epoch(num_mols: int, world: int, rank: int, batch: int, epoch: int, seed: int = 0, drop_last: bool = True):
    """Molecule ids of every step of `rank` in `epoch`: int32 [steps, batch].

    One permutation per epoch, shared by all ranks (same seed), strided by rank; the tail that
    does not fill a batch on every rank is dropped so that all ranks take the same number of
    steps (a rank that ran out of batches would deadlock the all-reduce)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    perm = np.random.Generator(np.random.PCG64([seed, epoch])).permutation(num_mols)
    steps = num_mols // (world * batch)
    if steps == 0 or not drop_last:
        raise ValueError(f"{num_mols} molecules do not fill one batch of {batch} on each of {world} ranks")
    mine = perm[rank::world][: steps * batch]
    return mine.astype(np.int32).reshape(steps, batch)


def head_split(offsets, num_gcn_layers: int) -> int:
    """Flat-buffer offset where the spectrum_predictor tensors start (parameters() order:
    2L GraphConv tensors, 2L BatchNorm tensors, then the 10 head tensors)."""
    return int(offsets[4 * num_gcn_layers])


class GradReducer:
    """Sum (not mean) of the flat gradient buffer over the ranks of `group`.

    `reduce(grads)` is the plain one-shot form.  With `overlap=True` on CUDA the caller marks
    the point of the backward pass where the head gradients are final (`head_ready()`); that
    bucket is reduced on a side stream while the GCN layers are differentiated, and
    `finish()` reduces the remaining bucket and joins the streams before AdamW."""

    def __init__(self, offsets, num_gcn_layers: int, group=None, overlap: bool = True):
        self.group = group
        self.split = head_split(offsets, num_gcn_layers)
        self.numel = int(offsets[-1])
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = overlap
        self._side = None
        self._pending = None

    def reduce(self, grads: torch.Tensor):
        if self.world > 1:
            dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=self.group)
        return grads

    # -- two-bucket overlapped form (CUDA) ---------------------------------------------
    def head_ready(self, grads: torch.Tensor):
        if self.world == 1:
            return
        if not (self.overlap and grads.is_cuda):
            return
        if self._side is None:
            self._side = torch.cuda.Stream(grads.device)
        cur = torch.cuda.current_stream(grads.device)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            dist.all_reduce(grads[self.split:], op=dist.ReduceOp.SUM, group=self.group)
        self._pending = grads

    def finish(self, grads: torch.Tensor):
        if self.world == 1:
            return grads
        if self._pending is None:
            return self.reduce(grads)
        dist.all_reduce(grads[: self.split], op=dist.ReduceOp.SUM, group=self.group)
        torch.cuda.current_stream(grads.device).wait_stream(self._side)
        self._pending = None
        return grads


def broadcast_params(fp, src: int = 0, group=None):
    """Identical initial weights and BatchNorm buffers on every rank (DDP's constructor)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(fp.params, src, group=group)
        dist.broadcast(fp.bn_running, src, group=group)


def train_step_dp(plan, ds, ids, fp, step, reducer: GradReducer, metrics=None, loss_kind="mse"):
    """One data-parallel optimiser step on this rank's batch: K1 + forward + loss + backward,
    gradient all-reduce (head bucket overlapped with the GCN backward), AdamW with
    grad_scale = 1/world."""
    if reducer.world == 1:
        plan.train_step(ds, ids, fp, step, metrics, loss_kind)
        return
    fp.ensure_adam()
    plan.train_step_split(ds, ids, fp, step, metrics, loss_kind, on_head_grads=lambda: reducer.head_ready(fp.grads))
    reducer.finish(fp.grads)
    plan.adamw(fp, step)
