"""Slab-sharded Voronoi cell construction: one process per GPU, torch.distributed for the plumbing.

The path shards by spatial slab: cells depend only on particles within their
termination radius, and grid cell ids are x-major (celery.rs:323-324), so a slab of grid x-planes
plus a halo of `h` planes on each side is everything a rank needs.

    bounds      all-reduce(min/max) of 6 scalars          -> identical grid parameters everywhere
    histogram   all-reduce(sum) of particles per x-plane  -> equal-count slab cuts
    exchange    ONE all-to-all(v) of 32-byte {x, y, z, id}   -> every rank receives its owned planes
                records                                       plus the halo planes (ghost particles)
    compute     tess_diagram_initialize_slab + tess_compute_all on the local planes
    verify      any cell whose search reached a plane outside the halo is flagged
                (TESS_STATUS_HALO_INSUFFICIENT); the exchange is redone with a wider halo.

Every rank bins with the same global bounds / cpd and orders candidates inside a grid cell by
global particle id, so per-cell results are bit-identical to the single-GPU run.

A `SlabPlan` (bounds, slab cuts, per-destination exchange counts) comes back with every result.  Handing it
to the next call tells the pipeline that the particle set is unchanged: the three planning collectives and
their host round trips are skipped, the pack runs without its counting pass, and the only host
synchronisation left in the step is the final flag (halo too thin | counts differ from the plan), one
all-reduce.  A plan that does not match the particles is detected by that flag and the step is redone
without it.

The compute backend is injected (`SlabBackend`): the product backend drives the CUDA library on
torch CUDA tensors; the CPU tests inject a numpy backend to exercise this host logic under gloo.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


def cells_per_dimension(n_global: int) -> int:
    """celery.rs:161-162 with libm cbrt (math.cbrt is glibc's); saturating cast."""
    v = math.cbrt(n_global / 1.25)
    return (int(v) if v > 0 else 0) + 1


def slab_cuts(plane_counts: np.ndarray, n_ranks: int) -> List[int]:
    """Plane indices c[0]=0 <= ... <= c[n_ranks]=cpd such that every slab [c[g], c[g+1]) holds
    about the same number of particles (prefix sum of the per-plane histogram)."""
    cpd = len(plane_counts)
    total = int(plane_counts.sum())
    pre = np.concatenate([[0], np.cumsum(plane_counts.astype(np.int64))])
    cuts = [0]
    for g in range(1, n_ranks):
        target = total * g / n_ranks
        c = int(np.searchsorted(pre, target, side="left"))
        # choose the nearer of the two neighbouring plane boundaries
        if c > 0 and abs(pre[c - 1] - target) <= abs(pre[min(c, cpd)] - target):
            c -= 1
        c = max(cuts[-1], min(c, cpd))
        cuts.append(c)
    cuts.append(cpd)
    # every rank must own at least one plane when there are enough planes
    if cpd >= n_ranks:
        for g in range(1, n_ranks):
            cuts[g] = max(cuts[g], cuts[g - 1] + 1)
        for g in range(n_ranks - 1, 0, -1):
            cuts[g] = min(cuts[g], cuts[g + 1] - 1)
    return cuts


def receive_ranges(cuts: Sequence[int], halo: int) -> Tuple[List[int], List[int]]:
    cpd = cuts[-1]
    lo = [max(0, cuts[g] - halo) for g in range(len(cuts) - 1)]
    hi = [min(cpd, cuts[g + 1] + halo) for g in range(len(cuts) - 1)]
    return lo, hi


def row_cuts(row_counts: np.ndarray, n_ranks: int) -> List[int]:
    """Slab cuts finer than a plane: keys k = x*cpd + y of grid rows, c[0]=0 <= ... <= c[n_ranks]=cpd^2, every slab
    [c[g], c[g+1]) holding about the same number of particles.  Cell ids are x-major (celery.rs:323-324), so a run of rows
    is a run of the sorted order; whole planes hold 50 k cells at 10M points, rows 250."""
    return slab_cuts(row_counts, n_ranks)


def receive_ranges_rows(cuts: Sequence[int], cpd: int, halo: int) -> Tuple[List[int], List[int]]:
    """Planes a rank must hold for row cuts: every plane it owns a row of, plus `halo` planes on each side."""
    lo = [max(0, cuts[g] // cpd - halo) for g in range(len(cuts) - 1)]
    hi = [min(cpd, -(-cuts[g + 1] // cpd) + halo) for g in range(len(cuts) - 1)]
    return lo, hi


@dataclass
class SlabPlan:
    """What a step learns about an unchanged particle set (compute_sharded(plan=...))."""

    bounds6: np.ndarray
    cuts: List[int]          # plane indices, or grid-row keys x*cpd + y when `rows`
    halo: int
    send_counts: List[int]
    recv_counts: List[int]
    n_local: int
    n_global: int
    rows: bool = False


@dataclass
class SlabResult:
    """Per-rank outcome: rows are the owned cells in the rank's grid order."""

    batch: object            # CellBatch (or whatever the backend returns)
    own: Tuple[int, ...]      # owned planes (lo, hi), or (lo_plane, lo_row, hi_plane, hi_row) with row cuts
    local: Tuple[int, int]
    n_owned: int
    halo: int
    n_received: int
    rounds: int
    plan: Optional[SlabPlan] = None
    halo_ok: bool = True     # False: max_rounds / the whole grid was reached with cells still flagged HALO_INSUFFICIENT


class SlabBackend:
    """What the host logic needs from a compute backend.  Arrays are torch tensors on the
    backend's device."""

    device = None

    def bounds(self, xyz):  # -> tensor[6] f64: x_min,x_max,y_min,y_max,z_min,z_max
        raise NotImplementedError

    def plane_histogram(self, xyz, bounds6: np.ndarray, n_global: int):  # -> tensor[cpd] int64
        raise NotImplementedError

    def pack(self, xyz, id_base: int, bounds6, n_global, lo, hi):  # -> (counts list[int], xyz_packed, ids_packed)
        raise NotImplementedError

    def compute(self, xyz, ids, box, bounds6, n_global, own, local, opts):  # -> (batch, n_owned, any_halo_flag: bool or 1-element int32 tensor)
        raise NotImplementedError

    # Optional fast path (one record all-to-all, plan reuse).  A backend without it gets the two-array exchange.
    def pack_records(self, xyz, id_base: int, bounds6, n_global, lo, hi, planned_counts=None):  # -> (counts list[int], records (m, 4) f64, counts_dev int64[R])
        raise NotImplementedError

    def compute_records(self, rec, box, bounds6, n_global, own, local, opts):  # -> (batch, n_owned, flag tensor int32[1])
        raise NotImplementedError

    has_records = False

    # Optional: particles per grid row (x, y) -> tensor[cpd*cpd] int64; with it slabs are cut at row, not plane, granularity
    def row_histogram(self, xyz, bounds6: np.ndarray, n_global: int):
        raise NotImplementedError

    has_rows = False


class CudaSlabBackend(SlabBackend):
    """Product backend: libtess_b200 on torch CUDA tensors of the current device."""

    def __init__(self, device_index: int):
        import torch

        from . import _lib
        from .interface import Diagram

        self._torch, self._lib, self._Diagram = torch, _lib, Diagram
        self.device_index = device_index
        self.device = torch.device("cuda", device_index)
        self._diagram = None
        self._rec_buf = None

    def _stream(self) -> int:
        return self._torch.cuda.current_stream(self.device).cuda_stream

    def bounds(self, xyz):
        out = self._torch.empty(6, dtype=self._torch.float64, device=self.device)
        self._lib.check(self._lib.lib().tess_bounds(xyz.data_ptr(), xyz.shape[0], out.data_ptr(), self._stream()))
        return out

    def plane_histogram(self, xyz, bounds6, n_global):
        cpd = cells_per_dimension(n_global)
        out = self._torch.empty(cpd, dtype=self._torch.int64, device=self.device)
        b = np.ascontiguousarray(bounds6, dtype=np.float64)
        self._lib.check(self._lib.lib().tess_plane_histogram(xyz.data_ptr(), xyz.shape[0], b.ctypes.data, n_global, out.data_ptr(), self._stream()))
        return out

    def pack(self, xyz, id_base, bounds6, n_global, lo, hi):
        torch = self._torch
        n, R = xyz.shape[0], len(lo)
        b = np.ascontiguousarray(bounds6, dtype=np.float64)
        plo, phi = np.ascontiguousarray(lo, dtype=np.uint32), np.ascontiguousarray(hi, dtype=np.uint32)
        counts = torch.empty(R, dtype=torch.int64, device=self.device)
        cap = int(n * 1.5) + 1024
        while True:
            oxyz = torch.empty((cap, 3), dtype=torch.float64, device=self.device)
            oids = torch.empty(cap, dtype=torch.int64, device=self.device)
            rc = self._lib.lib().tess_pack_for_slabs(xyz.data_ptr(), None, id_base, n, b.ctypes.data, n_global, R, plo.ctypes.data, phi.ctypes.data,
                                                     counts.data_ptr(), oxyz.data_ptr(), oids.data_ptr(), cap, self._stream())
            c = [int(v) for v in counts.cpu().tolist()]
            if rc == 0:
                tot = sum(c)
                return c, oxyz[:tot], oids[:tot]
            if rc != -4:
                self._lib.check(rc)
            cap = sum(c) + 1024

    has_records = True
    has_rows = True

    def row_histogram(self, xyz, bounds6, n_global):
        cpd = cells_per_dimension(n_global)
        out = self._torch.empty(cpd * cpd, dtype=self._torch.int64, device=self.device)
        b = np.ascontiguousarray(bounds6, dtype=np.float64)
        self._lib.check(self._lib.lib().tess_row_histogram(xyz.data_ptr(), xyz.shape[0], b.ctypes.data, n_global, out.data_ptr(), self._stream()))
        return out

    def pack_records(self, xyz, id_base, bounds6, n_global, lo, hi, planned_counts=None):
        torch = self._torch
        n, R = xyz.shape[0], len(lo)
        b = np.ascontiguousarray(bounds6, dtype=np.float64)
        plo, phi = np.ascontiguousarray(lo, dtype=np.uint32), np.ascontiguousarray(hi, dtype=np.uint32)
        counts_dev = torch.empty(R, dtype=torch.int64, device=self.device)
        hc = np.zeros(R, dtype=np.uint64)
        plan = None if planned_counts is None else np.ascontiguousarray(planned_counts, dtype=np.uint64)
        cap = (int(plan.sum()) if plan is not None else int(n * 1.5)) + 1024
        while True:
            if self._rec_buf is None or self._rec_buf.shape[0] < cap:
                self._rec_buf = torch.empty((cap, 4), dtype=torch.float64, device=self.device)
            rc = self._lib.lib().tess_pack_records(xyz.data_ptr(), None, id_base, n, b.ctypes.data, n_global, R, plo.ctypes.data, phi.ctypes.data,
                                                   None if plan is None else plan.ctypes.data, hc.ctypes.data, counts_dev.data_ptr(), self._rec_buf.data_ptr(),
                                                   self._rec_buf.shape[0], self._stream())
            c = [int(v) for v in hc]
            if rc == 0:
                return c, self._rec_buf[: sum(c)], counts_dev
            if rc != -4:
                self._lib.check(rc)
            cap = sum(c) + 1024

    def compute_records(self, rec, box, bounds6, n_global, own, local, opts):
        return self._compute(None, None, rec, box, bounds6, n_global, own, local, opts)

    def compute(self, xyz, ids, box, bounds6, n_global, own, local, opts):
        batch, n_owned, flag = self._compute(xyz, ids, None, box, bounds6, n_global, own, local, opts)
        return batch, n_owned, bool(flag.item())

    def _compute(self, xyz, ids, rec, box, bounds6, n_global, own, local, opts):
        from .interface import Polyhedron

        if self._diagram is None:
            self._diagram = self._Diagram(self.device_index)
        d = self._diagram
        d.clear()
        s = self._stream()
        if rec is not None:
            d.add_records_device(rec.data_ptr(), rec.shape[0], stream=s)
        else:
            d.add_particles_device(xyz.data_ptr(), xyz.shape[0], ids_ptr=ids.data_ptr(), stream=s)
        d.initialize_slab(Polyhedron(*box), bounds6, n_global, own, local, stream=s)
        opts = dict(opts)
        sink = opts.pop("host_sink", None)
        if sink is None:
            batch = d.compute_all_cells(stream=s, **opts)
        else:  # (volumes, face_offsets, neighbors, areas, status[, n_chunks]): results stream to the rank's host arrays
            batch = d.compute_all_cells_to_host(*sink[:5], n_chunks=(sink[5] if len(sink) > 5 else 0), stream=s, **opts)
        flag = self._torch.zeros(1, dtype=self._torch.int32, device=self.device)
        if batch.n_cells:
            # status words stay on the device, and so does the answer: one tiny reduction tells whether any halo was too thin
            st = _as_tensor(self._torch, batch.device_views()["status"], batch.n_cells, self._torch.int32, self.device)
            flag = ((st & self._lib.STATUS_HALO_INSUFFICIENT) != 0).any().to(self._torch.int32).reshape(1)
        return batch, batch.n_cells, flag


def _as_tensor(torch, ptr: int, n: int, dtype, device):
    """Zero-copy torch view of library-owned device memory (valid while the result lives)."""
    class _Holder:
        pass

    itemsize = torch.empty(0, dtype=dtype).element_size()
    h = _Holder()
    h.__cuda_array_interface__ = {
        "shape": (n,), "typestr": {4: "<i4", 8: "<i8"}[itemsize] if dtype in (torch.int32, torch.int64) else "<f8",
        "data": (ptr, False), "version": 2, "strides": None,
    }
    return torch.as_tensor(h, device=device)


def compute_sharded(backend: SlabBackend, xyz_local, id_base: int, n_global: int, box: Sequence[float], dist=None, halo: int = 4,
                    max_rounds: int = 4, opts: Optional[dict] = None, bounds6: Optional[np.ndarray] = None, plan: Optional[SlabPlan] = None) -> SlabResult:
    """Run the sharded hot path on this rank.

    xyz_local : (n_local, 3) f64 tensor on the backend's device — an arbitrary subset of the global
                particle set (global ids id_base .. id_base+n_local-1).
    dist      : torch.distributed (initialised) or None for a single process.
    plan      : the `plan` of an earlier result over the SAME particles (all ranks pass one, or none does).
    """
    import os
    import time

    import torch

    opts = dict(opts or {})
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    multi = dist is not None and world > 1
    n_local = int(xyz_local.shape[0])
    # TESS_SHARD_TRACE=1: synchronise after every phase and print where the step's time goes (rank 0; development aid)
    trace = [] if os.environ.get("TESS_SHARD_TRACE") else None
    t_last = [time.perf_counter()]

    def mark(name):
        if trace is None:
            return
        if hasattr(torch, "cuda") and torch.cuda.is_available() and getattr(backend, "device", None) is not None and backend.device.type == "cuda":
            torch.cuda.synchronize(backend.device)
        now = time.perf_counter()
        trace.append((name, (now - t_last[0]) * 1e3))
        t_last[0] = now

    if plan is not None and (plan.n_local != n_local or plan.n_global != n_global or len(plan.cuts) != world + 1 or not backend.has_records):
        raise ValueError("compute_sharded: the plan was made for another particle set / world size")
    cpd = cells_per_dimension(n_global)

    if plan is None:
        # ---- global bounds: CeleryBounds::new (celery.rs:81-125) over ALL particles: one all-reduce (max of {-min, max})
        if bounds6 is None:
            b = backend.bounds(xyz_local)
            if multi:
                mm = torch.cat([-b[0::2], b[1::2]])
                dist.all_reduce(mm, op=dist.ReduceOp.MAX)
                b = torch.stack([-mm[:3], mm[3:]], dim=1).reshape(-1)
            bounds6 = b.cpu().numpy().astype(np.float64)
        # ---- equal-count slab cuts from the per-row (else per-plane) histogram ----------------------
        rows = bool(getattr(backend, "has_rows", False)) and world > 1 and cpd >= 2
        hist = backend.row_histogram(xyz_local, bounds6, n_global) if rows else backend.plane_histogram(xyz_local, bounds6, n_global)
        if multi:
            dist.all_reduce(hist, op=dist.ReduceOp.SUM)
        cuts = row_cuts(hist.cpu().numpy(), world) if rows else slab_cuts(hist.cpu().numpy(), world)
        mark("plan (bounds, histogram, cuts)")
    else:
        bounds6, cuts, halo, rows = plan.bounds6, plan.cuts, plan.halo, plan.rows

    rounds = 0
    while True:
        rounds += 1
        if rows:
            lo, hi = receive_ranges_rows(cuts, cpd, halo)
            own = (cuts[rank] // cpd, cuts[rank] % cpd, cuts[rank + 1] // cpd, cuts[rank + 1] % cpd)  # (plane, row) .. (plane, row)
        else:
            lo, hi = receive_ranges(cuts, halo)
            own = (cuts[rank], cuts[rank + 1])
        local = (lo[rank], hi[rank])
        mismatch = None
        if backend.has_records:
            # ---- ghost-particle exchange: ONE all-to-all(v) of 32-byte records ---------------------
            planned = plan.send_counts if plan is not None else None
            send_counts, srec, counts_dev = backend.pack_records(xyz_local, id_base, bounds6, n_global, lo, hi, planned)
            mark("pack")
            if multi:
                if plan is not None:
                    recv_counts = plan.recv_counts
                    mismatch = (counts_dev != torch.tensor(send_counts, dtype=torch.int64, device=counts_dev.device)).any().to(torch.int32).reshape(1)
                else:
                    sc = torch.tensor(send_counts, dtype=torch.int64, device=srec.device)
                    rc = torch.empty_like(sc)
                    dist.all_to_all_single(rc, sc)
                    recv_counts = [int(v) for v in rc.cpu().tolist()]
                rrec = torch.empty((sum(recv_counts), 4), dtype=srec.dtype, device=srec.device)
                dist.all_to_all_single(rrec, srec, output_split_sizes=recv_counts, input_split_sizes=send_counts)
            else:
                rrec, recv_counts = srec, list(send_counts)
            mark("exchange")
            batch, n_owned, flag = backend.compute_records(rrec, box, bounds6, n_global, own, local, opts)
            mark("bin + clip + outputs")
            n_received = int(rrec.shape[0])
            flag = flag.to(torch.int32).reshape(1)
        else:
            send_counts, sxyz, sids = backend.pack(xyz_local, id_base, bounds6, n_global, lo, hi)
            if multi:
                sc = torch.tensor(send_counts, dtype=torch.int64, device=sxyz.device)
                rc = torch.empty_like(sc)
                dist.all_to_all_single(rc, sc)
                recv_counts = [int(v) for v in rc.cpu().tolist()]
                rxyz = torch.empty((sum(recv_counts), 3), dtype=sxyz.dtype, device=sxyz.device)
                rids = torch.empty(sum(recv_counts), dtype=sids.dtype, device=sids.device)
                dist.all_to_all_single(rxyz, sxyz, output_split_sizes=recv_counts, input_split_sizes=send_counts)
                dist.all_to_all_single(rids, sids, output_split_sizes=recv_counts, input_split_sizes=send_counts)
            else:
                rxyz, rids, recv_counts = sxyz, sids, list(send_counts)
            batch, n_owned, flagged = backend.compute(rxyz, rids, box, bounds6, n_global, own, local, opts)
            n_received = int(rxyz.shape[0])
            flag = torch.tensor([1 if flagged else 0], dtype=torch.int32, device=rxyz.device)
        # ---- was any halo too thin (bit 0)?  did the particles differ from the plan (bit 1)?  One all-reduce, one read.
        if mismatch is not None:
            flag = flag | (mismatch << 1)
        if multi:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        f = int(flag.item())
        mark("flag all-reduce")
        if trace is not None and rank == 0:
            print("[shard trace] " + " | ".join(f"{k} {v:.3f} ms" for k, v in trace) + f" | clip kernel {batch.timings()['clip_ms']:.3f} redo {batch.timings()['redo_ms']:.3f} outputs {batch.timings()['outputs_ms']:.3f} ms", flush=True)
        if f & 2:
            # the caller's promise did not hold: plan again from scratch
            if hasattr(batch, "close"):
                batch.close()
            return compute_sharded(backend, xyz_local, id_base, n_global, box, dist=dist, halo=halo, max_rounds=max_rounds, opts=opts, plan=None)
        done = (f & 1) == 0
        if done or rounds >= max_rounds or halo >= cpd:
            new_plan = SlabPlan(bounds6=bounds6, cuts=list(cuts), halo=halo, send_counts=list(send_counts), recv_counts=list(recv_counts), n_local=n_local, n_global=n_global,
                                rows=rows)
            return SlabResult(batch=batch, own=own, local=local, n_owned=n_owned, halo=halo, n_received=n_received, rounds=rounds, plan=new_plan, halo_ok=done)
        halo = cpd if rounds + 1 >= max_rounds else min(cpd, 2 * halo)  # the last round takes every plane: it cannot be flagged
        plan = None
