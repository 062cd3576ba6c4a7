"""the-tessellator_b200 — B200-native per-particle Voronoi cell construction.

One hot path of mcomstock/the-tessellator, rebuilt for sm_100a: GPU counting-sort binning
(celery.rs) + thread-per-cell / warp-per-cell half-space clipping (polyhedron.rs) behind the reference's
Diagram / Cell / VoronoiFace API (interface.rs).  See DESIGN.md.

The directory name contains a hyphen; import it with
    importlib.import_module("the-tessellator_b200")
"""
from . import _lib, generators  # noqa: F401
from ._lib import TessError, build  # noqa: F401
from .interface import Cell, CellBatch, Diagram, ExpandingSearch, Polyhedron, VoronoiFace, device_count, set_main_tier  # noqa: F401

__all__ = ["Diagram", "Cell", "VoronoiFace", "Polyhedron", "CellBatch", "ExpandingSearch", "TessError", "build", "device_count", "set_main_tier", "generators"]
