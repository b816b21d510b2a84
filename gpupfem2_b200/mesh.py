"""Host-side mesh containers and generators for the PFEM-2 particle step.

The particle path only *reads* the mesh: vertices, cells, inverse Jacobians and the
vertex-sharing one-ring CSR, exactly the arrays the reference's ``Mesh2D`` getters expose
(reference src/mesh_2d.cuh:18-44).  This module builds those arrays as numpy arrays; the handler
uploads them.  Nothing here runs on the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class HostMesh:
    """Same content as the reference ``Mesh2D`` getters (src/mesh_2d.cuh:18-44), on the host."""

    vertices: np.ndarray  # (N, 2) float64   -- getVertices(), Point2
    cells: np.ndarray  # (C, 3) uint32    -- getCells(), uint3, vertex order as in the file
    nbr_offsets: np.ndarray | None = None  # (C+1,) int32 -- getCellNeighborsOffsets()
    nbr_indices: np.ndarray | None = None  # (nnz,) int32 -- getCellNeighborIndices(), ascending per cell
    inv_jacobi: np.ndarray | None = None  # (C, 4) float64 -- getInvJacobi(), Matrix2x2 row-major
    meta: dict = field(default_factory=dict)

    @property
    def n_nodes(self) -> int:
        return int(self.vertices.shape[0])

    @property
    def n_cells(self) -> int:
        return int(self.cells.shape[0])


def load_dat(path: str, scale: float = 1.0) -> HostMesh:
    """Read the reference's DAT text format (reference src/mesh_2d.cu:36-96).

    Line 1: ``numVertices numEntities``; then ``id x y z`` per vertex; then ``id type ...`` per entity
    where type 203 is a triangle with three 1-based vertex ids and every other entity is skipped.
    """
    with open(path) as f:
        tok = f.read().split()
    nv, _ne = int(tok[0]), int(tok[1])
    pos = 2
    verts = np.empty((nv, 2), dtype=np.float64)
    for i in range(nv):
        verts[i, 0] = scale * float(tok[pos + 1])
        verts[i, 1] = scale * float(tok[pos + 2])
        pos += 4
    cells = []
    while pos + 1 < len(tok):
        typ = int(tok[pos + 1])
        if typ == 203:
            cells.append((int(tok[pos + 2]) - 1, int(tok[pos + 3]) - 1, int(tok[pos + 4]) - 1))
            pos += 5
        else:
            pos += 4
    return HostMesh(verts, np.asarray(cells, dtype=np.uint32).reshape(-1, 3), meta={"source": path})


def structured_channel(nx: int, ny: int, lx: float, ly: float, colmajor: bool = True) -> HostMesh:
    """Structured triangulated channel [0,lx]x[0,ly].

    nx x ny quads, each split along the same diagonal into two counter-clockwise triangles
    (a,b,c) and (a,c,d).  ``colmajor`` numbers nodes and quads x-major (node = i*(ny+1)+j,
    quad = i*ny+j) so that a strip of the channel in x is a contiguous cell-index range (the
    multi-GPU partition); otherwise row-major (node = j*(nx+1)+i, quad = j*nx+i).
    Identical arithmetic to oracle/ref_harness.cu:channel() (x = i*hx, y = j*hy).
    """
    hx, hy = lx / nx, ly / ny
    ii, jj = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="ij")

    def nid(i, j):
        return (i * (ny + 1) + j) if colmajor else (j * (nx + 1) + i)

    verts = np.empty(((nx + 1) * (ny + 1), 2), dtype=np.float64)
    verts[nid(ii, jj).ravel(), 0] = (ii * hx).ravel()
    verts[nid(ii, jj).ravel(), 1] = (jj * hy).ravel()
    qi, qj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    q = ((qi * ny + qj) if colmajor else (qj * nx + qi)).ravel()
    a, b = nid(qi, qj).ravel(), nid(qi + 1, qj).ravel()
    c, d = nid(qi + 1, qj + 1).ravel(), nid(qi, qj + 1).ravel()
    cells = np.empty((2 * nx * ny, 3), dtype=np.uint32)
    cells[2 * q] = np.stack([a, b, c], axis=1)
    cells[2 * q + 1] = np.stack([a, c, d], axis=1)
    return HostMesh(verts, cells, meta={"nx": nx, "ny": ny, "lx": lx, "ly": ly, "colmajor": colmajor})


def poiseuille_field(mesh: HostMesh, umax: float, height: float) -> tuple[np.ndarray, np.ndarray]:
    """u = 4 U y (H - y) / H^2, v = 0 at the nodes (same operation order as ref_harness.cu)."""
    y = mesh.vertices[:, 1]
    fx = 4.0 * umax * y * (height - y) / (height * height)
    return np.ascontiguousarray(fx), np.zeros_like(fx)


def vortex_field(mesh: HostMesh, umean: float, amp: float, wavelength: float) -> tuple[np.ndarray, np.ndarray]:
    """Mean flow plus a lattice of Taylor-Green vortices; divergence free."""
    k = 2.0 * np.pi / wavelength
    x, y = mesh.vertices[:, 0], mesh.vertices[:, 1]
    fx = umean + amp * np.sin(k * x) * np.cos(k * y)
    fy = -amp * np.cos(k * x) * np.sin(k * y)
    return np.ascontiguousarray(fx), np.ascontiguousarray(fy)


def write_dat(path: str, mesh: HostMesh) -> None:
    """Write the reference's DAT text format (vertices + type-203 triangles, 1-based ids) so that the reference's
    own Mesh2D::loadMeshFromFile (src/mesh_2d.cu:36-96) reads the mesh back bit for bit (%.14e like the shipped files)."""
    with open(path, "w") as f:
        f.write(f"{mesh.n_nodes} {mesh.n_cells}\n")
        for i, (x, y) in enumerate(mesh.vertices, start=1):
            f.write(f"{i} {x:.14e} {y:.14e} {0.0:.14e}\n")
        for k, (a, b, c) in enumerate(mesh.cells.astype(np.int64), start=1):
            f.write(f"{k} 203 {a + 1} {b + 1} {c + 1} \n")
