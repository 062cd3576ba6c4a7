"""world_size-2 gloo test of the slab-sharding host logic (the-tessellator_b200/distributed.py).

The compute backend is injected: here a numpy backend that bins like celery.rs and, instead of
clipping, returns the (id, plane) bookkeeping the host logic must get right — which particles
reach which rank (owned + ghost planes), identical grid parameters on every rank, halo widening
when a rank reports HALO_INSUFFICIENT.  The clipping of slab diagrams itself is covered on the
GPU (tests/test_gpu_parity.py::test_slab_decomposition_is_bit_identical).
"""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _axis_index(v, vmin, vmax, inv, cpd):
    """celery.rs:269-314 in numpy (saturating cast, clamp, upper edge)."""
    t = (v - vmin) * inv
    i = np.where(t > 0, np.minimum(t, 2.0 ** 62), 0.0).astype(np.int64)
    i = np.minimum(i, cpd - 1)
    return np.where(v >= vmax, cpd - 1, i)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_global, seed, thin_halo, out, records=False, rows=False):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    D = importlib.import_module("the-tessellator_b200.distributed")
    gen = importlib.import_module("the-tessellator_b200.generators")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    class NumpyBackend(D.SlabBackend):
        def __init__(self):
            self.calls = []

        def bounds(self, xyz):
            a = xyz.numpy()
            return torch.tensor([a[:, 0].min(), a[:, 0].max(), a[:, 1].min(), a[:, 1].max(), a[:, 2].min(), a[:, 2].max()], dtype=torch.float64)

        def _planes(self, xyz, b, n_global):
            cpd = D.cells_per_dimension(n_global)
            inv = cpd / (b[1] - b[0])
            return _axis_index(xyz.numpy()[:, 0], b[0], b[1], inv, cpd), cpd

        def plane_histogram(self, xyz, b, n_global):
            gx, cpd = self._planes(xyz, b, n_global)
            return torch.from_numpy(np.bincount(gx, minlength=cpd).astype(np.int64))

        def pack(self, xyz, id_base, b, n_global, lo, hi):
            gx, _ = self._planes(xyz, b, n_global)
            ids = id_base + np.arange(len(gx), dtype=np.int64)
            counts, px, pi = [], [], []
            for g in range(len(lo)):
                m = (gx >= lo[g]) & (gx < hi[g])
                counts.append(int(m.sum()))
                px.append(xyz.numpy()[m])
                pi.append(ids[m])
            return counts, torch.from_numpy(np.concatenate(px)), torch.from_numpy(np.concatenate(pi))

        def row_histogram(self, xyz, b, n_global):
            gx, cpd = self._planes(xyz, b, n_global)
            gy = _axis_index(xyz.numpy()[:, 1], b[2], b[3], cpd / (b[3] - b[2]), cpd)
            return torch.from_numpy(np.bincount(gx * cpd + gy, minlength=cpd * cpd).astype(np.int64))

        def compute(self, xyz, ids, box, b, n_global, own, local, opts):
            if len(own) == 4:  # row cuts: (lo_plane, lo_row, hi_plane, hi_row)
                gx, cpd = self._planes(xyz, b, n_global)
                gy = _axis_index(xyz.numpy()[:, 1], b[2], b[3], cpd / (b[3] - b[2]), cpd)
                key = gx * cpd + gy
                k0, k1 = own[0] * cpd + own[1], own[2] * cpd + own[3]
                if not getattr(self, "planned", False):
                    assert np.all((gx >= local[0]) & (gx < local[1])), "received a particle outside the local planes"
                owned = (key >= k0) & (key < k1)
                self.calls.append(dict(own=own, local=local, n=len(gx)))
                need = 3  # pretend cells need 3 planes of halo beyond every plane the rank owns a row of
                p0, p1 = own[0], -(-k1 // cpd)
                flagged = (p0 > 0 and p0 - local[0] < need) or (p1 < cpd and local[1] - p1 < need)
                batch = dict(ids=ids.numpy()[owned], recv_ids=ids.numpy(), bounds=np.array(b), own=own, local=local)
                return batch, int(owned.sum()), flagged
            gx, _ = self._planes(xyz, b, n_global)
            if not getattr(self, "planned", False):  # (records laid out by a stale plan are garbage: the step is discarded and redone)
                assert np.all((gx >= local[0]) & (gx < local[1])), "received a particle outside the local planes"
            owned = (gx >= own[0]) & (gx < own[1])
            halo_lo, halo_hi = own[0] - local[0], local[1] - own[1]
            self.calls.append(dict(own=own, local=local, n=len(gx)))
            # pretend cells need 3 planes of halo: flag when an interior side has fewer
            cpd = D.cells_per_dimension(n_global)
            need = 3
            flagged = (own[0] > 0 and halo_lo < need) or (own[1] < cpd and halo_hi < need)
            batch = dict(ids=ids.numpy()[owned], recv_ids=ids.numpy(), bounds=np.array(b), own=own, local=local)
            return batch, int(owned.sum()), flagged

    class RecordBackend(NumpyBackend):
        """The one-all-to-all path: 32-byte records, exchange counts known from a plan."""

        has_records = True

        def pack_records(self, xyz, id_base, b, n_global, lo, hi, planned_counts=None):
            counts, px, pi = self.pack(xyz, id_base, b, n_global, lo, hi)
            rec = torch.cat([px, pi.view(torch.float64).reshape(-1, 1)], dim=1)
            self.planned = planned_counts is not None
            if planned_counts is not None:  # the layout follows the plan: segments cut or padded to the planned counts
                seg, o = [], 0
                for have, want in zip(counts, planned_counts):
                    part = rec[o:o + min(have, want)]
                    seg.append(torch.cat([part, torch.zeros((want - part.shape[0], 4), dtype=torch.float64)]))
                    o += have
                rec = torch.cat(seg)
            return (list(planned_counts) if planned_counts is not None else counts), rec, torch.tensor(counts, dtype=torch.int64)

        def compute_records(self, rec, box, b, n_global, own, local, opts):
            batch, n_owned, flagged = self.compute(rec[:, :3].contiguous(), rec[:, 3].contiguous().view(torch.int64), box, b, n_global, own, local, opts)
            return batch, n_owned, torch.tensor([1 if flagged else 0], dtype=torch.int32)

    per = n_global // world
    start = rank * per
    n_local = per if rank < world - 1 else n_global - start
    # ranks hold arbitrary (index-contiguous, spatially random) subsets of one global stream
    xyz = torch.from_numpy(gen.uniform(n_local, seed, start=start))
    be = RecordBackend() if records else NumpyBackend()
    be.has_rows = rows
    res = D.compute_sharded(be, xyz, start, n_global, [0, 0, 0, 1, 1, 1], dist=dist, halo=1 if thin_halo else 4)
    if records:
        # the same particles again, with the plan: no planning collectives, same exchange, same result
        first = res
        n_before = len(be.calls)
        res = D.compute_sharded(be, xyz, start, n_global, [0, 0, 0, 1, 1, 1], dist=dist, plan=first.plan)
        assert be.planned and len(be.calls) == n_before + 1 and res.rounds == 1 and res.halo == first.halo
        assert np.array_equal(res.batch["recv_ids"], first.batch["recv_ids"]) and np.array_equal(res.batch["ids"], first.batch["ids"])
        # other particles with the old plan (a broken promise): detected through the counts, planned again
        xyz2 = torch.from_numpy(gen.uniform(n_local, seed + 1, start=start))
        again = D.compute_sharded(be, xyz2, start, n_global, [0, 0, 0, 1, 1, 1], dist=dist, plan=first.plan)
        fresh = D.compute_sharded(be, xyz2, start, n_global, [0, 0, 0, 1, 1, 1], dist=dist)
        assert np.array_equal(again.batch["recv_ids"], fresh.batch["recv_ids"]) and again.plan.send_counts == fresh.plan.send_counts
        assert res.halo_ok and again.halo_ok
    np.savez(out.format(rank=rank), ids=res.batch["ids"], recv_ids=res.batch["recv_ids"], bounds=res.batch["bounds"], own=np.array(res.own),
             local=np.array(res.local), halo=res.halo, rounds=res.rounds, n_owned=res.n_owned, n_calls=len(be.calls))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_row_cuts(tmp_path, gen):
    """Slab cuts finer than a plane (rows of the x-major grid): the two ranks share a plane, each owns its rows of it, both
    hold the whole plane plus the halo, every particle is owned exactly once and the split is balanced to a ROW's worth."""
    import torch.multiprocessing as mp

    D = importlib.import_module("the-tessellator_b200.distributed")
    world, n, seed = 2, 20_000, 78
    out = str(tmp_path / "rank{rank}.npz")
    mp.spawn(_worker, args=(world, _free_port(), n, seed, False, out, True, True), nprocs=world, join=True)
    pts = gen.uniform(n, seed)
    cpd = D.cells_per_dimension(n)
    b = np.array([pts[:, 0].min(), pts[:, 0].max(), pts[:, 1].min(), pts[:, 1].max(), pts[:, 2].min(), pts[:, 2].max()])
    gx = _axis_index(pts[:, 0], b[0], b[1], cpd / (b[1] - b[0]), cpd)
    gy = _axis_index(pts[:, 1], b[2], b[3], cpd / (b[3] - b[2]), cpd)
    key = gx * cpd + gy
    r = [np.load(out.format(rank=k)) for k in range(world)]
    k = [(int(o[0]) * cpd + int(o[1]), int(o[2]) * cpd + int(o[3])) for o in (r[0]["own"], r[1]["own"])]
    assert k[0][0] == 0 and k[0][1] == k[1][0] and k[1][1] == cpd * cpd
    assert k[0][1] % cpd != 0  # (with 20k random points the balanced cut does not fall on a plane boundary)
    assert abs(int(r[0]["n_owned"]) - n // 2) <= np.bincount(key).max()
    owned = np.concatenate([r[j]["ids"] for j in range(world)])
    assert np.array_equal(np.sort(owned), np.arange(n))
    for j in range(world):
        assert np.all((key[r[j]["ids"]] >= k[j][0]) & (key[r[j]["ids"]] < k[j][1]))
        lo, hi = r[j]["local"]
        assert np.array_equal(np.sort(r[j]["recv_ids"]), np.nonzero((gx >= lo) & (gx < hi))[0])
        assert lo == max(0, k[j][0] // cpd - 4) and hi == min(cpd, -(-k[j][1] // cpd) + 4)


@pytest.mark.parametrize("records", [False, True])
@pytest.mark.parametrize("thin_halo", [False, True])
def test_two_rank_slab_exchange(tmp_path, gen, thin_halo, records):
    import torch.multiprocessing as mp

    D = importlib.import_module("the-tessellator_b200.distributed")
    world, n, seed = 2, 20_000, 77
    out = str(tmp_path / "rank{rank}.npz")
    mp.spawn(_worker, args=(world, _free_port(), n, seed, thin_halo, out, records), nprocs=world, join=True)
    pts = gen.uniform(n, seed)
    cpd = D.cells_per_dimension(n)
    b = np.array([pts[:, 0].min(), pts[:, 0].max(), pts[:, 1].min(), pts[:, 1].max(), pts[:, 2].min(), pts[:, 2].max()])
    gx = _axis_index(pts[:, 0], b[0], b[1], cpd / (b[1] - b[0]), cpd)
    r = [np.load(out.format(rank=k)) for k in range(world)]
    # identical global grid parameters on every rank
    for k in range(world):
        assert np.array_equal(r[k]["bounds"], b)
    # slabs tile the planes, balanced to within one plane's worth of particles
    assert r[0]["own"][0] == 0 and r[0]["own"][1] == r[1]["own"][0] and r[1]["own"][1] == cpd
    assert abs(int(r[0]["n_owned"]) - n // 2) <= np.bincount(gx).max()
    # every particle is owned exactly once, by the rank holding its plane
    owned = np.concatenate([r[k]["ids"] for k in range(world)])
    assert np.array_equal(np.sort(owned), np.arange(n))
    for k in range(world):
        assert np.all((gx[r[k]["ids"]] >= r[k]["own"][0]) & (gx[r[k]["ids"]] < r[k]["own"][1]))
        lo, hi = r[k]["local"]
        expect = np.nonzero((gx >= lo) & (gx < hi))[0]
        assert np.array_equal(np.sort(r[k]["recv_ids"]), expect)  # owned + ghost planes, nothing else
        h = int(r[k]["halo"])
        assert lo == max(0, r[k]["own"][0] - h) and hi == min(cpd, r[k]["own"][1] + h)
    if thin_halo and records:  # the saved result is the planned second step: one round at the halo the first one found
        assert all(int(r[k]["rounds"]) == 1 and int(r[k]["halo"]) == 4 for k in range(world))
    elif thin_halo:  # halo 1 < 3 needed: widened 1 -> 2 -> 4, both ranks in lock-step
        assert all(int(r[k]["rounds"]) == 3 and int(r[k]["halo"]) == 4 for k in range(world))
    else:
        assert all(int(r[k]["rounds"]) == 1 and int(r[k]["halo"]) == 4 for k in range(world))


def test_slab_cuts_and_ranges():
    D = importlib.import_module("the-tessellator_b200.distributed")
    assert D.cells_per_dimension(10_000_000) == 201 and D.cells_per_dimension(10_000) == 21 and D.cells_per_dimension(1) == 1
    c = D.slab_cuts(np.full(201, 100), 8)
    assert c[0] == 0 and c[-1] == 201 and all(24 <= c[i + 1] - c[i] <= 26 for i in range(8))
    # clustered histogram: cuts follow the mass, not the width
    hist = np.ones(100, dtype=np.int64)
    hist[40:50] = 1000
    c = D.slab_cuts(hist, 4)
    assert c == sorted(c) and c[0] == 0 and c[-1] == 100
    loads = [hist[c[i]:c[i + 1]].sum() for i in range(4)]
    assert max(loads) <= 2 * (hist.sum() / 4) and all(c[i + 1] > c[i] for i in range(4))
    lo, hi = D.receive_ranges([0, 25, 50, 75, 100], 4)
    assert lo == [0, 21, 46, 71] and hi == [29, 54, 79, 100]
    # fewer planes than ranks: empty slabs are allowed, the cuts stay monotone
    c = D.slab_cuts(np.array([5, 5]), 4)
    assert c[0] == 0 and c[-1] == 2 and c == sorted(c)


def test_row_cuts_and_ranges():
    """Slab cuts at grid-row granularity: keys x*cpd + y; the planes a rank must hold are every plane it owns a row of plus
    the halo; whole-plane cuts are the special case own_*_row == 0."""
    D = importlib.import_module("the-tessellator_b200.distributed")
    cpd = 10
    rows = np.full(cpd * cpd, 7, dtype=np.int64)
    c = D.row_cuts(rows, 8)
    assert c[0] == 0 and c[-1] == cpd * cpd and c == sorted(c)
    loads = [int(rows[c[g]:c[g + 1]].sum()) for g in range(8)]
    assert max(loads) - min(loads) <= 7  # balanced to one row's worth (plane cuts: 10 rows' worth)
    lo, hi = D.receive_ranges_rows(c, cpd, 2)
    for g in range(8):
        first_plane, last_plane = c[g] // cpd, (c[g + 1] - 1) // cpd
        assert lo[g] == max(0, first_plane - 2) and hi[g] == min(cpd, last_plane + 1 + 2)
    # a cut on a plane boundary needs no extra plane
    lo, hi = D.receive_ranges_rows([0, 50, 100], cpd, 1)
    assert lo == [0, 4] and hi == [6, 10]
    # an uneven histogram: the cut follows the mass
    rows = np.ones(cpd * cpd, dtype=np.int64)
    rows[33] = 1000
    c = D.row_cuts(rows, 2)
    assert c[1] in (33, 34)
