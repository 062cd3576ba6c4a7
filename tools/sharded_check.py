#!/usr/bin/env python
"""Multi-GPU correctness check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py [n_points] [kind]

Every rank holds an index-contiguous (spatially random) part of one seeded point set, the slab
pipeline of the-tessellator_b200/distributed.py runs, and rank 0 compares every cell — volume, face
areas and neighbour list, bit for bit — with a single-GPU run over the whole set.
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T = importlib.import_module("the-tessellator_b200")
D = importlib.import_module("the-tessellator_b200.distributed")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    kind = sys.argv[2] if len(sys.argv) > 2 else "uniform"
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    gen = T.generators
    pts = gen.uniform(n, 3) if kind == "uniform" else gen.clustered(n, 4)
    per = n // world
    start = rank * per
    cnt = per if rank < world - 1 else n - start
    xyz = torch.from_numpy(pts[start:start + cnt]).to(dev)
    be = D.CudaSlabBackend(lr)
    res = D.compute_sharded(be, xyz, start, n, [0, 0, 0, 1, 1, 1], dist=dist, halo=4, opts=dict(outputs=7))
    assert res.halo_ok
    b = res.batch
    first = [a.copy() for a in (b.cell_ids, b.volumes, b.face_offsets, b.neighbors, b.areas, b.status)]
    # the same particles again with the first step's plan (no planning collectives, one record all-to-all)
    res2 = D.compute_sharded(be, xyz, start, n, [0, 0, 0, 1, 1, 1], dist=dist, opts=dict(outputs=7), plan=res.plan)
    b2 = res2.batch
    same = all(np.array_equal(x, y) for x, y in zip(first, (b2.cell_ids, b2.volumes, b2.face_offsets, b2.neighbors, b2.areas, b2.status))) and res2.rounds == 1
    t_same = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(t_same, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"planned step bit-identical = {bool(t_same.item())}", flush=True)
    b = b2
    ids = torch.from_numpy(b.cell_ids.copy()).to(dev)
    vol = torch.from_numpy(b.volumes.copy()).to(dev)
    cnts = torch.from_numpy(np.diff(b.face_offsets)).to(dev)
    nbr = torch.from_numpy(b.neighbors.copy()).to(dev)
    area = torch.from_numpy(b.areas.copy()).to(dev)
    st = int((b.status != 0).sum())
    print(f"[rank {rank}] own planes {res.own} local {res.local} owned {res.n_owned} received {res.n_received} halo {res.halo} rounds {res.rounds} flagged {st}", flush=True)

    def gather(t):
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=dev))
        m = int(max(s.item() for s in sizes))
        pad = torch.zeros(m, dtype=t.dtype, device=dev)
        pad[: t.shape[0]] = t
        out = [torch.zeros(m, dtype=t.dtype, device=dev) for _ in range(world)]
        dist.all_gather(out, pad)
        return [o[: int(s.item())].cpu().numpy() for o, s in zip(out, sizes)]

    g_ids, g_vol, g_cnt, g_nbr, g_area = gather(ids), gather(vol), gather(cnts), gather(nbr), gather(area)
    ok = True
    if rank == 0:
        d = T.Diagram(lr)
        d.add_particles(pts)
        d.initialize(T.Polyhedron(0, 0, 0, 1, 1, 1))
        w = d.compute_all_cells(outputs=7)
        fo = w.face_offsets
        seen = np.zeros(n, bool)
        for r in range(world):
            i = g_ids[r]
            seen[i] = True
            ok &= bool(np.array_equal(g_vol[r], w.volumes[i]))
            ok &= bool(np.array_equal(g_cnt[r], np.diff(fo)[i]))
            # face lists: concatenate the whole-domain slices in the rank's row order
            sel = np.concatenate([np.arange(fo[k], fo[k + 1]) for k in i[:: max(1, len(i) // 20000)]])
            off = np.concatenate([[0], np.cumsum(g_cnt[r])])
            mine = np.concatenate([np.arange(off[k], off[k + 1]) for k in range(0, len(i), max(1, len(i) // 20000))])
            ok &= bool(np.array_equal(g_nbr[r][mine], w.neighbors[sel])) and bool(np.array_equal(g_area[r][mine], w.areas[sel]))
        ok &= bool(seen.all())
        total = sum(float(v.sum()) for v in g_vol)
        print(f"sharded over {world} GPUs vs single GPU: bit-identical = {ok}; cells {int(seen.sum())}/{n}; sum of volumes {total!r}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
