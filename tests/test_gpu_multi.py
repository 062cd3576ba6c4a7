"""Tests that need TWO GPUs in one box (skipped on a one-GPU box; run them with `gpurun --gpus 2`).

* the slab pipeline over real NCCL (the-tessellator_b200/distributed.py, one process per GPU under torchrun) against the
  single-GPU run, bit for bit — including a second step that reuses the first one's SlabPlan;
* two diagrams on two devices in ONE process: the kernels' launch configuration (dynamic shared-memory opt-in, SM count,
  occupancy) is cached per device (clip.cu launch_cfg), and the large-cell configuration needs the opt-in on both.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BOX = [0, 0, 0, 1, 1, 1]


def _need_two(tess):
    if tess.device_count() < 2:
        pytest.skip("needs two GPUs in one box")


@pytest.mark.parametrize("kind", ["uniform", "clustered"])
def test_two_rank_nccl_run_equals_single_gpu(tess, kind):
    _need_two(tess)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(ROOT, "tools", "sharded_check.py"), "400000", kind]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "bit-identical = True" in p.stdout and "planned step bit-identical = True" in p.stdout, p.stdout[-3000:]


def test_two_devices_in_one_process(tess, gen, ob):
    _need_two(tess)
    u = gen.uniform(300, 54)
    th, ph = np.arccos(2 * u[:, 0] - 1), 2 * np.pi * u[:, 1]
    shell = 0.5 + 0.3 * np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)
    bg = gen.uniform(2000, 55)
    pts = np.concatenate([[[0.5, 0.5, 0.5]], shell, bg[np.linalg.norm(bg - 0.5, axis=1) > 0.35]])  # cell 0 needs the large configuration
    r = ob.Diagram(pts, box=BOX).compute_cells(mode=ob.MODE_SECURITY)
    assert len(r.cell_neighbors(0)) > 100
    import threading

    out = {}

    def run(dev):
        d = tess.Diagram(dev)
        d.add_particles(pts)
        d.initialize(tess.Polyhedron(*BOX))
        out[dev] = d.compute_all_cells(outputs=7)

    # device 1 FIRST (a per-process cache would configure only that one), then device 0, then both at once from two host threads
    run(1)
    run(0)
    for dev in (0, 1):
        helpers.assert_cells_identical(out[dev], r, what=f"device {dev}")
        assert np.all(out[dev].status == 0)
    ts = [threading.Thread(target=run, args=(dev,)) for dev in (0, 1)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for dev in (0, 1):
        helpers.assert_cells_identical(out[dev], r, what=f"device {dev}, concurrent")
