"""ctypes binding of the CPU oracle (oracle/pfem2_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpfem2_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
            os.path.join(_HERE, "pfem2_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, _dp, _up, _dp, _ip, _ip, C.c_int, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_particles_per_cell.argtypes = [C.c_void_p]
        L.orc_subcell_centers.argtypes = [C.c_void_p, _dp]
        L.orc_seed.argtypes = [C.c_void_p]
        L.orc_init_velocity.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_advect.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_int]
        L.orc_project.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_move.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_int]
        L.orc_check_distribution.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_int]
        L.orc_project_accumulate.argtypes = [C.c_void_p, _dp]
        L.orc_correct.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.orc_count.argtypes = [C.c_void_p]
        L.orc_last_stats.argtypes = [C.c_void_p, _ip, _ip]
        L.orc_download.argtypes = [C.c_void_p] + [_dp] * 7 + [_up] * 2
        L.orc_upload.argtypes = [C.c_void_p, C.c_int] + [_dp] * 7 + [_up] * 2
        L.orc_inv_jacobi.argtypes = [C.c_int, _dp, _up, _dp]
        L.orc_one_ring.argtypes = [C.c_int, C.c_int, _up, _ip, _ip]
        L.orc_to_local.argtypes = [_dp, _dp, C.c_double, C.c_double, _dp]
        L.orc_inside.argtypes = [_dp]
        L.orc_subcell.argtypes = [_dp, C.c_int]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _u(a):
    return a.ctypes.data_as(_up) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def inv_jacobi(vertices, cells):
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    c = np.ascontiguousarray(cells, dtype=np.uint32)
    out = np.empty((c.shape[0], 4), dtype=np.float64)
    lib().orc_inv_jacobi(c.shape[0], _d(v), _u(c), _d(out))
    return out


def one_ring(n_nodes, cells):
    c = np.ascontiguousarray(cells, dtype=np.uint32)
    off = np.empty(c.shape[0] + 1, dtype=np.int32)
    lib().orc_one_ring(n_nodes, c.shape[0], _u(c), _i(off), None)
    idx = np.empty(int(off[-1]), dtype=np.int32)
    lib().orc_one_ring(n_nodes, c.shape[0], _u(c), _i(off), _i(idx))
    return off, idx


def complete_mesh(mesh):
    """Fill nbr_offsets / nbr_indices / inv_jacobi of a HostMesh with the oracle's restatements."""
    if mesh.nbr_offsets is None:
        mesh.nbr_offsets, mesh.nbr_indices = one_ring(mesh.n_nodes, mesh.cells)
    if mesh.inv_jacobi is None:
        mesh.inv_jacobi = inv_jacobi(mesh.vertices, mesh.cells)
    return mesh


class OracleHandler:
    """Mirror of the reference ParticleHandler2D interface (particle_handler_2d.cuh:9-54) on the CPU."""

    FIELDS = ("x", "y", "l0", "l1", "l2", "vx", "vy", "cell", "id")

    def __init__(self, mesh, cell_division_level, max_level=4, subcell_mode=0):
        complete_mesh(mesh)
        self.mesh = mesh
        self._keep = [np.ascontiguousarray(mesh.vertices, dtype=np.float64),
                      np.ascontiguousarray(mesh.cells, dtype=np.uint32),
                      np.ascontiguousarray(mesh.inv_jacobi, dtype=np.float64),
                      np.ascontiguousarray(mesh.nbr_offsets, dtype=np.int32),
                      np.ascontiguousarray(mesh.nbr_indices, dtype=np.int32)]
        v, c, j, o, i = self._keep
        self._h = C.c_void_p(lib().orc_create(mesh.n_nodes, mesh.n_cells, _d(v), _u(c), _d(j), _i(o), _i(i),
                                              cell_division_level, max_level, subcell_mode))
        self.particles_per_cell = lib().orc_particles_per_cell(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def subcell_centers(self):
        out = np.empty((self.particles_per_cell, 3), dtype=np.float64)
        lib().orc_subcell_centers(self._h, _d(out))
        return out

    def seed_particles(self):
        return lib().orc_seed(self._h)

    def init_particle_velocity(self, fx, fy):
        lib().orc_init_velocity(self._h, _d(fx), _d(fy))

    def advect_particles(self, fx, fy, dt, substeps):
        return lib().orc_advect(self._h, _d(fx), _d(fy), dt, substeps)

    def move(self, fx, fy, dt, substeps):
        return lib().orc_move(self._h, _d(fx), _d(fy), dt, substeps)

    def check_distribution(self, fx, fy, own_lo, own_hi):
        return lib().orc_check_distribution(self._h, _d(fx), _d(fy), own_lo, own_hi)

    def project_accumulate(self):
        acc = np.zeros((self.mesh.n_nodes, 3), dtype=np.float64)
        lib().orc_project_accumulate(self._h, _d(acc))
        return acc

    def project_velocity_onto_grid(self, wx, wy):
        lib().orc_project(self._h, _d(wx), _d(wy))

    def correct_particle_velocity(self, fx, fy, ox, oy):
        lib().orc_correct(self._h, _d(fx), _d(fy), _d(ox), _d(oy))

    def particle_count(self):
        return lib().orc_count(self._h)

    def last_stats(self):
        lost = np.zeros(16, dtype=np.int32)
        added = C.c_int(0)
        lib().orc_last_stats(self._h, _i(lost), C.byref(added))
        return lost, added.value

    def download(self):
        n = self.particle_count()
        out = {k: np.empty(n, dtype=np.float64) for k in self.FIELDS[:7]}
        out["cell"] = np.empty(n, dtype=np.uint32)
        out["id"] = np.empty(n, dtype=np.uint32)
        lib().orc_download(self._h, *[_d(out[k]) for k in self.FIELDS[:7]], _u(out["cell"]), _u(out["id"]))
        return out

    def upload(self, state):
        n = int(state["x"].shape[0])
        a = [np.ascontiguousarray(state[k], dtype=np.float64) for k in self.FIELDS[:7]]
        c = np.ascontiguousarray(state["cell"], dtype=np.uint32)
        i = np.ascontiguousarray(state["id"], dtype=np.uint32)
        lib().orc_upload(self._h, n, *[_d(v) for v in a], _u(c), _u(i))

    def step(self, fx, fy, wx, wy, dt, substeps):
        """One isolated particle step (ref_harness.cu protocol): advect(F) ; project(W) ; correct(F, W)."""
        n = self.advect_particles(fx, fy, dt, substeps)
        self.project_velocity_onto_grid(wx, wy)
        self.correct_particle_velocity(fx, fy, wx, wy)
        return n
