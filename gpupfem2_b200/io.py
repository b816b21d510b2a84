"""Mesh and particle-state I/O around the particle step.

The reference reads meshes from DAT text (src/mesh_2d.cu:36-96), has no checkpoint / restart at all and exports particles
as ASCII VTK after materialising its 96-byte AoS on the host (src/data_export.cu:97-171).  This module adds

* ``read_dat_fast``            -- the same DAT format through one vectorised parse (a 16M-triangle DAT is ~1 GB of text),
* ``save_mesh`` / ``load_mesh`` -- binary mesh (npz: vertices, cells and, when present, the one-ring CSR / inverse Jacobians),
* ``save_checkpoint`` / ``load_checkpoint`` / ``restore`` -- particle state (the nine per-particle arrays of
  ``pfem2_download`` / ``pfem2_upload``) plus the handler's division level and a fingerprint of the mesh it belongs to,
* ``device_columns``           -- zero-copy device views of the position / velocity / cell / id columns of the sorted 64-byte
  record array (no AoS materialisation),
* ``write_particles_vtu``      -- the particle file of ``DataExport::exportParticlesToVTK`` either as the reference's ASCII
  text (same fields, same number formatting) or with raw appended binary arrays.

Nothing here is on the hot path; only numpy is required (torch only for ``device_columns``).
"""
from __future__ import annotations

import base64
import hashlib
import json
import os
import struct

import numpy as np

from .mesh import HostMesh

FIELDS = ("x", "y", "l0", "l1", "l2", "vx", "vy", "cell", "id")
CHECKPOINT_VERSION = 1


# ------------------------------------------------------------------------------------------------
# meshes
# ------------------------------------------------------------------------------------------------
def read_dat_fast(path: str, scale: float = 1.0) -> HostMesh:
    """Reference DAT text (src/mesh_2d.cu:36-96) -> HostMesh, vectorised.

    Line 1: ``numVertices numEntities``; ``id x y z`` per vertex; ``id type v...`` per entity, where only type 203
    (triangle, three 1-based vertex ids) is kept, exactly like ``Mesh2D::loadMeshFromFile``.  Entities of other types may
    carry a different number of ids, so the entity block is parsed line-wise but without a Python loop over tokens."""
    with open(path, "rb") as f:
        head = f.readline().split()
        nv, _ne = int(head[0]), int(head[1])
        rest = f.read()
    lines = rest.split(b"\n")
    vtxt = b" ".join(lines[:nv])
    v = np.array(vtxt.split(), dtype=np.float64).reshape(nv, 4)
    verts = np.ascontiguousarray(v[:, 1:3]) * scale if scale != 1.0 else np.ascontiguousarray(v[:, 1:3])
    ent = [ln for ln in lines[nv:] if ln.strip()]
    tri = [ln for ln in ent if ln.split(None, 2)[1] == b"203"]
    if tri:
        t = np.array(b" ".join(tri).split(), dtype=np.int64).reshape(len(tri), 5)
        cells = (t[:, 2:5] - 1).astype(np.uint32)
    else:
        cells = np.empty((0, 3), dtype=np.uint32)
    return HostMesh(verts, np.ascontiguousarray(cells), meta={"source": path})


def mesh_fingerprint(mesh) -> str:
    """sha256 over (n_nodes, n_cells, first / last 4096 cells and vertices): cheap identity check for checkpoints."""
    verts = np.ascontiguousarray(np.asarray(mesh.vertices, dtype=np.float64))
    cells = np.ascontiguousarray(np.asarray(mesh.cells, dtype=np.uint32))
    h = hashlib.sha256()
    h.update(struct.pack("<qq", verts.shape[0], cells.shape[0]))
    for a in (verts, cells):
        h.update(a[:4096].tobytes())
        h.update(a[-4096:].tobytes())
    return h.hexdigest()


def save_mesh(path: str, mesh: HostMesh) -> None:
    arrays = {"vertices": np.asarray(mesh.vertices, dtype=np.float64), "cells": np.asarray(mesh.cells, dtype=np.uint32)}
    for k in ("nbr_offsets", "nbr_indices", "inv_jacobi"):
        if getattr(mesh, k) is not None:
            arrays[k] = np.asarray(getattr(mesh, k))
    arrays["meta"] = np.frombuffer(json.dumps({k: v for k, v in mesh.meta.items() if isinstance(v, (int, float, str, bool))}).encode(),
                                   dtype=np.uint8)
    np.savez(path, **arrays)


def load_mesh(path: str) -> HostMesh:
    d = np.load(path)
    meta = json.loads(bytes(d["meta"]).decode()) if "meta" in d.files else {}
    return HostMesh(d["vertices"], d["cells"], d["nbr_offsets"] if "nbr_offsets" in d.files else None,
                    d["nbr_indices"] if "nbr_indices" in d.files else None, d["inv_jacobi"] if "inv_jacobi" in d.files else None, meta)


# ------------------------------------------------------------------------------------------------
# particle state: checkpoint / restart
# ------------------------------------------------------------------------------------------------
def save_checkpoint(path: str, state: dict, *, level: int, mesh=None, step: int = 0, time: float = 0.0, extra: dict | None = None) -> None:
    """state: the dict of ``ParticleHandler2D.download()`` (x y l0 l1 l2 vx vy float64, cell id uint32), any order."""
    n = int(np.asarray(state["x"]).shape[0])
    arrays = {}
    for k in FIELDS:
        a = np.asarray(state[k])
        if a.shape != (n,):
            raise ValueError(f"field {k} has shape {a.shape}, expected ({n},)")
        arrays[k] = a.astype(np.uint32 if k in ("cell", "id") else np.float64, copy=False)
    meta = {"version": CHECKPOINT_VERSION, "count": n, "level": int(level), "step": int(step), "time": float(time),
            "mesh": mesh_fingerprint(mesh) if mesh is not None else None, "extra": extra or {}}
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    tmp = path + ".tmp.npz"
    np.savez(tmp, **arrays)
    os.replace(tmp, path)  # a crash never leaves a truncated checkpoint under the final name


def load_checkpoint(path: str, mesh=None) -> tuple[dict, dict]:
    """-> (state, meta); raises if the file was written for another mesh."""
    d = np.load(path)
    meta = json.loads(bytes(d["meta"]).decode())
    if meta.get("version") != CHECKPOINT_VERSION:
        raise ValueError(f"unsupported checkpoint version {meta.get('version')}")
    if mesh is not None and meta.get("mesh") is not None and meta["mesh"] != mesh_fingerprint(mesh):
        raise ValueError("checkpoint belongs to a different mesh")
    state = {k: np.ascontiguousarray(d[k]) for k in FIELDS}
    if state["x"].shape[0] != meta["count"]:
        raise ValueError("checkpoint is inconsistent: count does not match the arrays")
    return state, meta


def checkpoint_handler(path: str, handler, mesh=None, **kw) -> None:
    """Download the state of a (CUDA or oracle) handler and write it."""
    level = kw.pop("level", None)
    if level is None:
        level = int(round(float(getattr(handler, "particles_per_cell", 0)) ** 0.5)) or int(getattr(handler, "level", 0))
    save_checkpoint(path, handler.download(), level=level, mesh=mesh, **kw)


def restore(path: str, handler, mesh=None) -> dict:
    """Load a checkpoint into a freshly created handler (``upload`` re-sorts by cell); returns the metadata."""
    state, meta = load_checkpoint(path, mesh)
    handler.upload(state)
    return meta


# ------------------------------------------------------------------------------------------------
# particle export
# ------------------------------------------------------------------------------------------------
def device_columns(handler) -> dict:
    """Zero-copy DEVICE views into the sorted 64-byte record array of a CUDA handler:
    pos (n, 2) f64, lab (n, 2) f64, l2 (n,) f64, vel (n, 2) f64, cell (n,) int32, id (n,) int32.
    Valid until the next mutating call.  This is the device-side export of SURVEY §8f row 4: no 96-byte AoS is built."""
    import ctypes as C

    import torch

    from .handler import _wrap_device

    n = handler.get_particle_count()
    p = C.c_void_p()
    handler._check(handler._L.pfem2_device_records(handler._h, C.byref(p)), "device_records")
    if n == 0:
        z = torch.empty((0, 2), dtype=torch.float64, device=handler.mesh.device)
        return {"pos": z, "lab": z, "l2": z[:, 0], "vel": z, "cell": z[:, 0].to(torch.int32), "id": z[:, 0].to(torch.int32)}
    rec = _wrap_device(p.value, (n, 8), "<f8", handler.mesh.device)       # {x, y | L0, L1 | L2, (cell, id) | vx, vy}
    ints = rec.view(torch.int32).view(n, 16)
    return {"pos": rec[:, 0:2], "lab": rec[:, 2:4], "l2": rec[:, 4], "vel": rec[:, 6:8], "cell": ints[:, 10], "id": ints[:, 11]}


def _fmt(v: float) -> str:
    return "%g" % v  # C++ ostream default formatting of a double (precision 6), as in data_export.cu


def write_particles_vtu(path: str, x, y, vx, vy, *, binary: bool = False) -> None:
    """The file of ``DataExport::exportParticlesToVTK`` (src/data_export.cu:97-171): one vertex cell per particle,
    Points (x, y, 0) and a 3-component ``velocity`` point array, declared Float32 like the reference.

    binary=False reproduces the reference's ASCII layout (readable by the same tools and by tests/insitu_compare.parse_vtu);
    binary=True writes the same arrays as base64-encoded inline binary (header = one UInt32 byte count per array)."""
    x, y, vx, vy = (np.asarray(a, dtype=np.float64).ravel() for a in (x, y, vx, vy))
    n = x.shape[0]
    if not (y.shape[0] == vx.shape[0] == vy.shape[0] == n):
        raise ValueError("x, y, vx, vy must have the same length")

    def enc(a: np.ndarray) -> str:
        raw = np.ascontiguousarray(a).tobytes()
        return base64.b64encode(struct.pack("<I", len(raw))).decode() + base64.b64encode(raw).decode()

    with open(path, "w") as f:
        w = f.write
        w('<?xml version="1.0" ?> \n')
        w('<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">\n')
        w("  <UnstructuredGrid>\n")
        w(f'    <Piece NumberOfPoints="{n}" NumberOfCells="{n}">\n')
        w("      <Points>\n")
        if binary:
            pts = np.zeros((n, 3), dtype=np.float32)
            pts[:, 0], pts[:, 1] = x, y
            w('        <DataArray type="Float32" NumberOfComponents="3" format="binary">\n')
            w("          " + enc(pts) + "\n")
        else:
            w('        <DataArray type="Float32" NumberOfComponents="3" Format="ascii">\n')
            w("".join(f"          {_fmt(a)} {_fmt(b)} 0.0\n" for a, b in zip(x, y)))
        w("        </DataArray>\n      </Points>\n      <Cells>\n")
        ids = np.arange(n, dtype=np.int32)
        for name, arr in (("connectivity", ids), ("offsets", ids + 1), ("types", np.ones(n, dtype=np.int32))):
            if binary:
                w(f'        <DataArray type="Int32" Name="{name}" format="binary">\n          {enc(arr)}\n        </DataArray>\n')
            else:
                w(f'        <DataArray type="Int32" Name="{name}" Format="ascii">\n        ')
                w("".join(f"  {v}" for v in arr))
                w("\n        </DataArray>\n")
        w("      </Cells>\n")
        w('      <PointData Scalars="scalars">\n')
        if binary:
            vel = np.zeros((n, 3), dtype=np.float32)
            vel[:, 0], vel[:, 1] = vx, vy
            w('        <DataArray type="Float32" Name="velocity" NumberOfComponents="3" format="binary">\n')
            w("          " + enc(vel) + "\n")
        else:
            w('        <DataArray type="Float32" Name="velocity" NumberOfComponents="3" Format="ascii">\n')
            w("".join(f"          {_fmt(a)} {_fmt(b)} 0.0\n" for a, b in zip(vx, vy)))
        w("        </DataArray>\n      </PointData>\n    </Piece>\n  </UnstructuredGrid>\n</VTKFile>\n")


def read_particles_vtu(path: str) -> dict:
    """Read back a file written by ``write_particles_vtu`` (either flavour) or by the reference: {points (n,3), velocity (n,3)}."""
    import re

    txt = open(path).read()
    out = {}
    for m in re.finditer(r"<DataArray([^>]*)>(.*?)</DataArray>", txt, flags=re.S):
        attrs, body = m.group(1), m.group(2).strip()
        name = re.search(r'Name="([^"]+)"', attrs)
        key = name.group(1) if name else "points"
        typ = re.search(r'type="([^"]+)"', attrs).group(1)
        dt = {"Float32": np.float32, "Int32": np.int32, "Float64": np.float64}[typ]
        if re.search(r'[Ff]ormat="binary"', attrs):
            raw = base64.b64decode(body[8:])  # 8 base64 characters = the 4-byte length header (padded)
            a = np.frombuffer(raw, dtype=dt)
        else:
            a = np.array(body.split(), dtype=np.float64).astype(dt)
        nc = re.search(r'NumberOfComponents="(\d+)"', attrs)
        out[key] = a.reshape(-1, int(nc.group(1))) if nc else a
    return out
