"""The C++ mirror of the reference API (include/tess.hpp) driven by a small C++ program; its printed
numbers are checked against the CPU oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_harness_matches_oracle(tess, gen, ob, tmp_path):
    exe = str(tmp_path / "test_interface")
    libdir = os.path.dirname(tess._lib.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_interface.cpp"),
                           "-o", exe, "-L", libdir, "-ltess_b200", f"-Wl,-rpath,{libdir}"])
    n = 3000
    out = subprocess.run([exe, str(n)], capture_output=True, text=True, check=True).stdout
    pts = gen.uniform(n, 62)
    od = ob.Diagram(pts, box=[0, 0, 0, 1, 1, 1])
    r = od.compute_cells(mode=ob.MODE_SECURITY)
    seen = 0
    for line in out.splitlines():
        m = re.match(r"cell (\d+) volume (\S+) faces (.*)", line)
        if m:
            i, v = int(m.group(1)), float(m.group(2))
            faces = sorted((int(a), float(b)) for a, b in (t.split(":") for t in m.group(3).split()))
            exp = sorted(zip(r.cell_neighbors(i).tolist(), r.cell_areas(i).tolist()))
            assert abs(v - r.volumes[i]) <= 1e-12 * r.volumes[i]
            assert [f[0] for f in faces] == [e[0] for e in exp]
            assert all(abs(f[1] - e[1]) <= 1e-12 * e[1] for f, e in zip(faces, exp))
            seen += 1
    assert seen == 6
    assert abs(float(re.search(r"total (\S+)", out).group(1)) - 1.0) <= 1e-12
    q = od.compute_cell_at_point(0.31, 0.62, 0.44)
    mq = re.search(r"query volume (\S+) nfaces (\d+)", out)
    assert abs(float(mq.group(1)) - q.volumes[0]) <= 1e-12 * q.volumes[0] and int(mq.group(2)) == len(q.neighbors)
    assert "expected error -5" in out
    me = re.search(r"expand steps (\d+) sweep (\d+) equal (\d) cursor (\d+) cells_in_radius (\d+)", out)
    es = od.expanding_search(0.31, 0.62, 0.44)
    want = es.expand(0.01, 200)
    assert int(me.group(3)) == 1 and int(me.group(1)) == int(me.group(2)) == len(want) > 0
    assert int(me.group(5)) == len(od.find_cells_in_radius(0.31, 0.62, 0.44, 0.1))
