"""Convert the reference's shipped DAT meshes into compact binary fixtures (run in the build
container, where /root/reference exists; the GPU box has no /root/reference).

    python tests/golden/make_mesh_fixtures.py

Writes tests/golden/mesh_channel.npz (cases/PoiseuilleFlow2D/ChannelMesh.dat) and
tests/golden/mesh_cylinder3.npz (cases/Cylinder2D/CylinderMesh3.dat): vertices (N,2) f64, cells (C,3) u32.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from gpupfem2_b200.mesh import load_dat  # noqa: E402

REF = os.environ.get("PFEM2_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
for name, rel in (("channel", "cases/PoiseuilleFlow2D/ChannelMesh.dat"),
                  ("cylinder3", "cases/Cylinder2D/CylinderMesh3.dat")):
    m = load_dat(os.path.join(REF, rel))
    np.savez_compressed(os.path.join(OUT, f"mesh_{name}.npz"), vertices=m.vertices, cells=m.cells)
    print(name, m.n_nodes, m.n_cells)
