"""The reference-facing call with HOST buffers: a batch of molecules (packed MolTable slice +
target spectra, as dense rows or as the peak lists they are binned from) in pinned host memory -> one H2D copy -> K1 + forward + loss +
backward + AdamW -> loss / cosine read back to the host.

This is what `collate_fn` + `.to(device)` + one `train_model` iteration do in the reference
(GCN:292-297, 411-439); bench.py times it as `e2e`.  Two device slots are used so the
upload of batch i+1 and its K1 batch build (both on the copy stream; the plan double-buffers
the batch tables) overlap the compute of batch i.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import Dataset, check
from .synth import MolTable


def _align(n, a=256):
    return (n + a - 1) // a * a


class PackedHostBatch:
    """One batch packed into a single pinned byte buffer:
    [node_ptr i64 | bond_ptr i64 | bond_begin i32 | bond_end i32 | feat f32 | targets f32]
    or, with peaks=(peak_ptr, mz, intensity) of the batch instead of dense targets,
    [... | feat f32 | peak_ptr i64 | peak_mz f32/f64 | peak_inten f32]  (binned on the device inside
    the loss kernel: ~1 KB instead of 4*max_mz bytes per molecule over PCIe)."""

    def __init__(self, table: MolTable, targets: np.ndarray | None, pin=True, peaks=None):
        B, N, nb = table.num_mols, int(table.node_ptr[-1]), int(table.bond_ptr[-1])
        F = table.feat.shape[1] if table.feat.ndim == 2 else 6
        parts = [("node_ptr", table.node_ptr.astype(np.int64)), ("bond_ptr", table.bond_ptr.astype(np.int64)),
                 ("bond_begin", table.bond_begin.astype(np.int32)), ("bond_end", table.bond_end.astype(np.int32)),
                 ("feat", np.ascontiguousarray(table.feat, np.float32).reshape(-1))]
        self.mz_is_f64 = 0
        if targets is not None:
            parts.append(("targets", np.ascontiguousarray(targets, np.float32).reshape(-1)))
        elif peaks is not None:
            pptr, mz, inten = peaks
            mz = np.ascontiguousarray(mz)
            if mz.dtype != np.float32:
                mz = mz.astype(np.float64)
            self.mz_is_f64 = int(mz.dtype == np.float64)
            pad = lambda a: a if a.size else np.zeros(1, a.dtype)
            parts += [("peak_ptr", np.ascontiguousarray(pptr, np.int64)), ("peak_mz", pad(mz)),
                      ("peak_inten", pad(np.ascontiguousarray(inten, np.float32)))]
        self.offsets, off = {}, 0
        for name, a in parts:
            self.offsets[name] = off
            off = _align(off + a.nbytes)
        self.nbytes = off
        self.buf = torch.empty(off, dtype=torch.uint8, pin_memory=pin)
        view = self.buf.numpy()
        for name, a in parts:
            o = self.offsets[name]
            view[o:o + a.nbytes] = a.view(np.uint8).reshape(-1)
        self.num_graphs, self.num_nodes, self.num_edges, self.feat_dim = B, N, 2 * nb, F
        self.has_targets = targets is not None
        self.has_peaks = targets is None and peaks is not None


class HostBatchRunner:
    """Runs training / inference steps whose inputs start in (pinned) host memory."""

    def __init__(self, plan, fp, capacity_bytes, n_slots=2):
        self.plan, self.fp = plan, fp
        dev = plan.device
        self.slots = [torch.empty(capacity_bytes, dtype=torch.uint8, device=dev) for _ in range(n_slots)]
        self.copy_stream = torch.cuda.Stream(dev)
        self.copied = [torch.cuda.Event() for _ in range(n_slots)]
        self.consumed = [None] * n_slots
        self.built = [False] * n_slots
        self.metrics = torch.zeros(8, dtype=torch.float32, device=dev)
        self.host_metrics = torch.zeros(8, dtype=torch.float32, pin_memory=True)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._i = 0

    def _dataset(self, slot, hb: PackedHostBatch) -> Dataset:
        base = self.slots[slot].data_ptr()
        o = hb.offsets
        pk = None
        if hb.has_peaks:
            pk = _lib.Peaks(base + o["peak_ptr"], base + o["peak_mz"], base + o["peak_inten"], hb.mz_is_f64, hb.num_graphs)
        ds = Dataset(base + o["node_ptr"], base + o["bond_ptr"], base + o["feat"], base + o["bond_begin"],
                     base + o["bond_end"], (base + o["targets"]) if hb.has_targets else None, hb.num_graphs,
                     C.pointer(pk) if pk is not None else None)
        ds._peaks_keepalive = pk
        return ds

    def upload(self, hb: PackedHostBatch, build: bool = False):
        """Enqueue the H2D copy of `hb` (and, with build=True, its K1 batch build) on the copy
        stream; returns the slot index.  With build=True uploads and steps must alternate."""
        slot = self._i % len(self.slots)
        self._i += 1
        if hb.nbytes > self.slots[slot].numel():
            raise _lib.EimsError(_lib.ERR_CAPACITY, "host batch larger than the staging slot")
        if len(self.slots) != 2 and build:
            raise ValueError("building ahead needs exactly two slots (the plan has two sets of batch tables)")
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[slot] is not None:
                self.copy_stream.wait_event(self.consumed[slot])  # the step that read this slot is done
            self.slots[slot][:hb.nbytes].copy_(self.buf_of(hb), non_blocking=True)
            if build:
                # K1 on the copy stream: the table set it overwrites was last read by the step that
                # consumed this slot, which the wait above already covers
                ds = self._dataset(slot, hb)
                check(self.plan.lib.eims_batch_build(self.plan.h, C.byref(ds), None, hb.num_graphs,
                                                     C.c_void_p(self.copy_stream.cuda_stream)))
            self.built[slot] = build
            self.copied[slot].record(self.copy_stream)
        self.h2d_bytes += hb.nbytes
        return slot

    @staticmethod
    def buf_of(hb):
        return hb.buf[:hb.nbytes]

    def train_step(self, slot, hb: PackedHostBatch, step, loss_kind="mse", optimizer=True):
        """K1 + forward + loss + backward + AdamW on the uploaded batch; the step's loss and
        cosine are copied back to pinned host memory (asynchronously; read with `result`)."""
        cur = torch.cuda.current_stream(self.plan.device)
        cur.wait_event(self.copied[slot])
        ds = self._dataset(slot, hb)
        self.fp.ensure_adam()
        fp = self.fp
        if self.built[slot]:
            # targets as peak lists: point the plan at this slot's arrays (the struct is copied)
            check(self.plan.lib.eims_plan_set_peak_targets(self.plan.h, ds.peaks if hb.has_peaks else None))
            self.plan._peak_targets = None
            check(self.plan.lib.eims_train_step_built(self.plan.h, C.c_void_p(ds.targets), None, _lib.ptr(fp.params),
                                                      _lib.ptr(fp.grads), _lib.ptr(fp.adam_m) if optimizer else None,
                                                      _lib.ptr(fp.adam_v) if optimizer else None, _lib.ptr(fp.bn_running),
                                                      _lib.LOSS[loss_kind], C.byref(step), _lib.ptr(self.metrics),
                                                      self.plan.stream))
        else:
            check(self.plan.lib.eims_train_step(self.plan.h, C.byref(ds), None, hb.num_graphs, _lib.ptr(fp.params),
                                                _lib.ptr(fp.grads), _lib.ptr(fp.adam_m) if optimizer else None,
                                                _lib.ptr(fp.adam_v) if optimizer else None,
                                                _lib.ptr(fp.bn_running), _lib.LOSS[loss_kind], C.byref(step),
                                                _lib.ptr(self.metrics), self.plan.stream))
        ev = torch.cuda.Event()
        ev.record(cur)
        self.consumed[slot] = ev
        self.host_metrics.copy_(self.metrics, non_blocking=True)
        self.d2h_bytes += self.metrics.numel() * 4
        for l in range(self.plan.d.num_gcn_layers):
            fp.num_batches_tracked[l] += 1

    def infer(self, slot, hb: PackedHostBatch, out_host: torch.Tensor):
        cur = torch.cuda.current_stream(self.plan.device)
        cur.wait_event(self.copied[slot])
        ds = self._dataset(slot, hb)
        fp = self.fp
        check(self.plan.lib.eims_infer_batch(self.plan.h, C.byref(ds), None, hb.num_graphs, _lib.ptr(fp.params),
                                             _lib.ptr(fp.bn_running), None, self.plan.stream))
        ev = torch.cuda.Event()
        ev.record(cur)
        self.consumed[slot] = ev
        prob = self.plan.buffer("prob", torch.float32, (hb.num_graphs, self.plan.d.max_mz))
        out_host[:hb.num_graphs].copy_(prob, non_blocking=True)
        self.d2h_bytes += prob.numel() * 4

    def result(self):
        """(loss, cosine) of the last finished training step (host sync)."""
        torch.cuda.current_stream(self.plan.device).synchronize()
        return float(self.host_metrics[4]), float(self.host_metrics[5])
