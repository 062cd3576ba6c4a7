"""Binning passes (tess_diagram_initialize) over N uniform points: the target of the ncu capture of K1-K4."""
import importlib
import sys
import traceback

sys.path.insert(0, ".")
T = importlib.import_module("the-tessellator_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
pts = T.generators.uniform(n, 3)
try:
    d = T.Diagram(0)
    d.add_particles(pts)
    for k in range(2):
        d.initialize(T.Polyhedron(0, 0, 0, 1, 1, 1))
        print("initialized", n, "pass", k, flush=True)
    d.close()
except Exception:
    traceback.print_exc()
