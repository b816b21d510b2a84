"""Binary case / dump files exchanged with oracle/ref_harness.cu (test + baseline infrastructure).

A *case file* carries everything the reference needs to run the isolated particle step on the GPU
box (mesh, one-ring, frozen nodal field, step protocol); a *dump* is the particle + projected nodal
state the reference produced.  Layouts are documented in oracle/ref_harness.cu.
"""
from __future__ import annotations

import numpy as np

MAGIC = 0x50464D32


def write_case(path, mesh, fx, fy, level, substeps, dt, nsteps, dump_steps):
    hdr = np.array([MAGIC, mesh.n_nodes, mesh.n_cells, mesh.nbr_indices.size, level, substeps, nsteps,
                    len(dump_steps)], dtype=np.int64)
    with open(path, "wb") as f:
        hdr.tofile(f)
        np.array([dt], dtype=np.float64).tofile(f)
        np.asarray(dump_steps, dtype=np.int64).tofile(f)
        np.ascontiguousarray(mesh.vertices, dtype=np.float64).tofile(f)
        np.ascontiguousarray(mesh.cells, dtype=np.uint32).tofile(f)
        np.ascontiguousarray(mesh.nbr_offsets, dtype=np.int32).tofile(f)
        np.ascontiguousarray(mesh.nbr_indices, dtype=np.int32).tofile(f)
        np.ascontiguousarray(fx, dtype=np.float64).tofile(f)
        np.ascontiguousarray(fy, dtype=np.float64).tofile(f)


def read_dump(path):
    """-> dict(x, y, l0, l1, l2, vx, vy, cell, id, wx, wy)."""
    with open(path, "rb") as f:
        n, nn = (int(v) for v in np.fromfile(f, dtype=np.int64, count=2))
        out = {}
        for k in ("x", "y", "l0", "l1", "l2", "vx", "vy"):
            out[k] = np.fromfile(f, dtype=np.float64, count=n)
        out["cell"] = np.fromfile(f, dtype=np.uint32, count=n)
        out["id"] = np.fromfile(f, dtype=np.uint32, count=n)
        out["wx"] = np.fromfile(f, dtype=np.float64, count=nn)
        out["wy"] = np.fromfile(f, dtype=np.float64, count=nn)
    return out


def canonical_order(state):
    """Permutation sorting particles by (cell, x, y) -- SURVEY N6: array order is scheduling dependent."""
    return np.lexsort((state["y"], state["x"], state["cell"]))
