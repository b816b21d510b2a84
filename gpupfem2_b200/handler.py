"""Host-side mirror of the reference's ``ParticleHandler2D`` over the C ABI (include/pfem2_b200.h).

Same methods, argument meaning and call order as reference src/particles/particle_handler_2d.cuh:9-54
(``seedParticles``, ``initParticleVelocity``, ``advectParticles``, ``projectVelocityOntoGrid``,
``correctParticleVelocity``, ``getParticles``, ``getParticleCount``), spelled in snake_case.  Nodal
velocities are pairs of CUDA float64 torch tensors (the reference's ``deviceVector<double*>`` of two
component arrays).  torch is plumbing only: device memory and streams; all compute is in the CUDA library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .mesh import HostMesh


class Pfem2Error(RuntimeError):
    pass


class DeviceMesh:
    """The reference Mesh2D's device arrays (src/mesh_2d.cuh:18-44) as torch CUDA tensors."""

    cell_base = node_base = 0        # offsets of a mesh slice into the global numbering (partitioned multi-GPU runs)
    n_cells_global = n_nodes_global = None

    def __init__(self, mesh: HostMesh, device="cuda:0", build_missing=True):
        self.device = torch.device(device)
        self.n_nodes, self.n_cells = mesh.n_nodes, mesh.n_cells
        self.vertices = torch.as_tensor(np.ascontiguousarray(mesh.vertices, dtype=np.float64)).to(self.device)
        self.cells = torch.as_tensor(np.ascontiguousarray(mesh.cells, dtype=np.uint32).view(np.int32)).to(self.device)
        L = _lib.load()
        if mesh.inv_jacobi is not None:
            self.inv_jacobi = torch.as_tensor(np.ascontiguousarray(mesh.inv_jacobi, dtype=np.float64)).to(self.device)
        elif build_missing:
            self.inv_jacobi = torch.empty((self.n_cells, 4), dtype=torch.float64, device=self.device)
            with torch.cuda.device(self.device):
                rc = L.pfem2_mesh_inv_jacobi(self.n_cells, self.vertices.data_ptr(), self.cells.data_ptr(),
                                             self.inv_jacobi.data_ptr(), None)
            if rc:
                raise Pfem2Error(L.pfem2_last_error(None).decode())
        if mesh.nbr_offsets is not None:
            self.nbr_offsets = torch.as_tensor(np.ascontiguousarray(mesh.nbr_offsets, dtype=np.int32)).to(self.device)
            self.nbr_indices = torch.as_tensor(np.ascontiguousarray(mesh.nbr_indices, dtype=np.int32)).to(self.device)
        elif build_missing:
            self.nbr_offsets, self.nbr_indices = device_one_ring(self.n_nodes, self.cells)
        torch.cuda.synchronize(self.device)

    @classmethod
    def from_tensors(cls, vertices: torch.Tensor, cells: torch.Tensor):
        """Build from device tensors (vertices (N,2) float64, cells (C,3) int32 bit-pattern of uint32); inverse
        Jacobians and the one-ring are computed on the device by the library."""
        self = cls.__new__(cls)
        self.device = vertices.device
        self.n_nodes, self.n_cells = int(vertices.shape[0]), int(cells.shape[0])
        self.vertices, self.cells = vertices.contiguous(), cells.contiguous()
        L = _lib.load()
        self.inv_jacobi = torch.empty((self.n_cells, 4), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = L.pfem2_mesh_inv_jacobi(self.n_cells, self.vertices.data_ptr(), self.cells.data_ptr(),
                                         self.inv_jacobi.data_ptr(), None)
        if rc:
            raise Pfem2Error(L.pfem2_last_error(None).decode())
        self.nbr_offsets, self.nbr_indices = device_one_ring(self.n_nodes, self.cells)
        torch.cuda.synchronize(self.device)
        return self

    def view(self) -> _lib.MeshView:
        return _lib.MeshView(self.n_nodes, self.n_cells, self.vertices.data_ptr(), self.cells.data_ptr(),
                             self.inv_jacobi.data_ptr(), self.nbr_offsets.data_ptr(), self.nbr_indices.data_ptr())


def device_structured_channel(nx, ny, lx, ly, colmajor=True, device="cuda:0", col_lo=None, col_hi=None) -> "DeviceMesh":
    """mesh.structured_channel generated directly in HBM (same numbering and the same x = i*hx, y = j*hy
    arithmetic, so the vertex bits agree with the host generator and with oracle/ref_harness.cu).

    col_lo / col_hi (colmajor only): only the quad columns [col_lo, col_hi) of the global channel are generated -- the mesh
    slice of one strip of a partitioned multi-GPU run (owned columns + halo).  Cell and node ids are relative to the slice
    (`cell_base`, `node_base` give the offsets into the global numbering, `n_cells_global` the global size); coordinates use
    the GLOBAL column index, so their bits are those of the global mesh."""
    dev = torch.device(device)
    hx, hy = lx / nx, ly / ny
    i0, i1 = (0, nx) if col_lo is None else (int(col_lo), int(col_hi))
    if (i0, i1) != (0, nx) and not colmajor:
        raise ValueError("column slices need the column-major numbering")
    if not 0 <= i0 < i1 <= nx:
        raise ValueError("bad column range")
    i = torch.arange(i0, i1 + 1, device=dev, dtype=torch.int64)
    j = torch.arange(ny + 1, device=dev, dtype=torch.int64)
    ni = i1 - i0
    if colmajor:  # node = i*(ny+1)+j
        x = (i.to(torch.float64) * hx)[:, None].expand(ni + 1, ny + 1)
        y = (j.to(torch.float64) * hy)[None, :].expand(ni + 1, ny + 1)
    else:  # node = j*(nx+1)+i
        x = (i.to(torch.float64) * hx)[None, :].expand(ny + 1, nx + 1)
        y = (j.to(torch.float64) * hy)[:, None].expand(ny + 1, nx + 1)
    vertices = torch.stack([x.reshape(-1), y.reshape(-1)], dim=1).contiguous()

    def nid(a, b):  # (a = column index relative to the slice)
        return a * (ny + 1) + b if colmajor else b * (nx + 1) + a

    qi = torch.arange(ni, device=dev, dtype=torch.int64)
    qj = torch.arange(ny, device=dev, dtype=torch.int64)
    if colmajor:  # quad = i*ny+j
        I, J = qi[:, None].expand(ni, ny).reshape(-1), qj[None, :].expand(ni, ny).reshape(-1)
    else:  # quad = j*nx+i
        I, J = qi[None, :].expand(ny, nx).reshape(-1), qj[:, None].expand(ny, nx).reshape(-1)
    a, b, c, d = nid(I, J), nid(I + 1, J), nid(I + 1, J + 1), nid(I, J + 1)
    cells = torch.stack([a, b, c, a, c, d], dim=1).reshape(-1, 3).to(torch.int32).contiguous()
    dm = DeviceMesh.from_tensors(vertices, cells)
    dm.cell_base, dm.node_base = 2 * ny * i0, (ny + 1) * i0
    dm.n_cells_global, dm.n_nodes_global = 2 * nx * ny, (nx + 1) * (ny + 1)
    return dm


def device_mesh_slice(vertices, cells, cell_lo: int, cell_hi: int, device="cuda:0") -> "DeviceMesh":
    """The cells [cell_lo, cell_hi) of a global mesh (host arrays or tensors: vertices (N, 2) float64, cells (C, 3)) as a
    DeviceMesh of its own: node ids relative to the smallest node id the slice touches.  For numberings in which a contiguous
    cell range touches a contiguous-enough node range (banded numberings: sort the cells by centroid x and number the nodes
    accordingly); inverse Jacobians and the one-ring are rebuilt on the device, cell by cell the same bits as the global mesh's."""
    v = torch.as_tensor(np.ascontiguousarray(vertices, dtype=np.float64)) if not isinstance(vertices, torch.Tensor) else vertices
    c = torch.as_tensor(np.ascontiguousarray(cells).astype(np.int64)) if not isinstance(cells, torch.Tensor) else cells.to(torch.int64)
    sl = c[int(cell_lo):int(cell_hi)]
    n_lo, n_hi = int(sl.min()), int(sl.max()) + 1
    dm = DeviceMesh.from_tensors(v[n_lo:n_hi].to(device).contiguous(), (sl - n_lo).to(torch.int32).to(device).contiguous())
    dm.cell_base, dm.node_base = int(cell_lo), n_lo
    dm.n_cells_global, dm.n_nodes_global = int(c.shape[0]), int(v.shape[0])
    return dm


def mesh_band(mesh: "DeviceMesh") -> int:
    """max |neighbour - cell| over the one-ring lists: how far a particle's cell INDEX can move in one substep."""
    L = _lib.load()
    out = C.c_int(0)
    with torch.cuda.device(mesh.device):
        rc = L.pfem2_mesh_band(mesh.n_cells, mesh.nbr_offsets.data_ptr(), mesh.nbr_indices.data_ptr(), C.byref(out), None)
    if rc:
        raise Pfem2Error(L.pfem2_last_error(None).decode())
    return out.value


def device_one_ring(n_nodes: int, cells: torch.Tensor):
    """Vertex-sharing one-ring CSR built on the device (replaces Mesh2D::fillCellNeighborIndices, mesh_2d.cu:107-139)."""
    L = _lib.load()
    n_cells = cells.shape[0]
    off = torch.empty(n_cells + 1, dtype=torch.int32, device=cells.device)
    nnz = C.c_int(0)
    with torch.cuda.device(cells.device):
        rc = L.pfem2_mesh_one_ring(n_nodes, n_cells, cells.data_ptr(), off.data_ptr(), None, C.byref(nnz), None)
        if rc:
            raise Pfem2Error(L.pfem2_last_error(None).decode())
        idx = torch.empty(max(nnz.value, 1), dtype=torch.int32, device=cells.device)
        rc = L.pfem2_mesh_one_ring(n_nodes, n_cells, cells.data_ptr(), off.data_ptr(), idx.data_ptr(), C.byref(nnz), None)
        if rc:
            raise Pfem2Error(L.pfem2_last_error(None).decode())
    return off, idx[: nnz.value]


class ParticleHandler2D:
    FIELDS = ("x", "y", "l0", "l1", "l2", "vx", "vy", "cell", "id")

    def __init__(self, mesh: DeviceMesh, cell_division_level: int, *, subcell_mode=0, max_division_level=4,
                 capacity_factor=1.5, verbose=False, exact_search=False, stable_order=False, defer_correct=True, host_pipeline=0,
                 lazy_sort=True, graph_advect=0):
        self._L = _lib.load()
        self.mesh = mesh  # borrowed for the handler's lifetime, like the reference's `const Mesh2D *`
        opt = _lib.Options()
        self._L.pfem2_default_options(C.byref(opt))
        opt.subcell_mode = subcell_mode
        opt.max_division_level = max_division_level
        opt.capacity_factor = capacity_factor
        opt.device = mesh.device.index if mesh.device.index is not None else 0
        opt.verbose = 1 if verbose else 0
        opt.exact_search = 1 if exact_search else 0
        opt.stable_order = 1 if stable_order else 0
        opt.defer_correct = 1 if defer_correct else 0
        opt.host_pipeline = int(host_pipeline)
        opt.lazy_sort = 1 if lazy_sort else 0
        opt.graph_advect = int(graph_advect)
        self._h = C.c_void_p()
        view = mesh.view()
        rc = self._L.pfem2_create(C.byref(self._h), C.byref(view), cell_division_level, C.byref(opt))
        if rc:
            raise Pfem2Error(f"pfem2_create: {self._L.pfem2_last_error(None).decode()}")
        lvl = max(min(cell_division_level, max_division_level), 1)
        self.particles_per_cell = lvl * lvl

    def close(self):
        if getattr(self, "_h", None):
            self._L.pfem2_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc:
            raise Pfem2Error(f"{what}: {self._L.pfem2_last_error(self._h).decode()} (code {rc})")

    # --- reference interface -------------------------------------------------------------------
    def seed_particles(self):
        self._check(self._L.pfem2_seed(self._h), "seedParticles")

    def init_particle_velocity(self, vel):
        self._check(self._L.pfem2_init_velocity(self._h, vel[0].data_ptr(), vel[1].data_ptr()), "initParticleVelocity")

    def advect_particles(self, vel, time_step: float, particle_substeps: int):
        self._check(self._L.pfem2_advect(self._h, vel[0].data_ptr(), vel[1].data_ptr(), time_step, particle_substeps),
                    "advectParticles")

    def project_velocity_onto_grid(self, vel):
        """Overwrites vel[0], vel[1] with the projected nodal velocity (like the reference)."""
        self._check(self._L.pfem2_project(self._h, vel[0].data_ptr(), vel[1].data_ptr()), "projectVelocityOntoGrid")

    def project_velocity_onto_grid_dual(self, vel, vel_copy):
        """Extension: projectVelocityOntoGrid(vel) that also writes the result into vel_copy (the cases' copy_d2d into the
        "old" solution right after the call, cases/Cylinder2D/main.cu:804-805)."""
        self._check(self._L.pfem2_project_dual(self._h, vel[0].data_ptr(), vel[1].data_ptr(), vel_copy[0].data_ptr(),
                                               vel_copy[1].data_ptr()), "projectVelocityOntoGrid")

    def correct_particle_velocity(self, vel, vel_old):
        self._check(self._L.pfem2_correct(self._h, vel[0].data_ptr(), vel[1].data_ptr(), vel_old[0].data_ptr(),
                                          vel_old[1].data_ptr()), "correctParticleVelocity")

    def get_particle_count(self) -> int:
        n = C.c_int(0)
        self._check(self._L.pfem2_particle_count(self._h, C.byref(n)), "getParticleCount")
        return n.value

    def get_particles(self) -> torch.Tensor:
        """The reference's 96-byte AoS Particle2D records as a (count, 12) float64 view (device memory owned by
        the handler, valid until the next mutating call)."""
        p, n = C.c_void_p(), C.c_int(0)
        self._check(self._L.pfem2_export_aos(self._h, C.byref(p), C.byref(n)), "getParticles")
        if not n.value:
            return torch.empty((0, 12), dtype=torch.float64, device=self.mesh.device)
        return _wrap_device(p.value, (n.value, 12), "<f8", self.mesh.device)

    # --- pointer-table flavour (deviceVector<double*>::data) --------------------------------------
    def advect_particles_ptrs(self, table: torch.Tensor, time_step, particle_substeps):
        self._check(self._L.pfem2_advect_ptrs(self._h, table.data_ptr(), time_step, particle_substeps), "advectParticles")

    def project_velocity_onto_grid_ptrs(self, table: torch.Tensor):
        self._check(self._L.pfem2_project_ptrs(self._h, table.data_ptr()), "projectVelocityOntoGrid")

    def project_velocity_onto_grid_dual_ptrs(self, table: torch.Tensor, table_copy: torch.Tensor):
        self._check(self._L.pfem2_project_dual_ptrs(self._h, table.data_ptr(), table_copy.data_ptr()), "projectVelocityOntoGrid")

    def correct_particle_velocity_ptrs(self, table: torch.Tensor, table_old: torch.Tensor):
        self._check(self._L.pfem2_correct_ptrs(self._h, table.data_ptr(), table_old.data_ptr()), "correctParticleVelocity")

    def init_particle_velocity_ptrs(self, table: torch.Tensor):
        self._check(self._L.pfem2_init_velocity_ptrs(self._h, table.data_ptr()), "initParticleVelocity")

    # --- extras --------------------------------------------------------------------------------
    def step(self, frozen, work, dt, substeps):
        """Isolated particle step (oracle/ref_harness.cu protocol): advect(F) ; project(W) ; correct(F, W)."""
        self.advect_particles(frozen, dt, substeps)
        self.project_velocity_onto_grid(work)
        self.correct_particle_velocity(frozen, work)

    def step_host(self, fx: np.ndarray, fy: np.ndarray, wx: np.ndarray, wy: np.ndarray, dt, substeps) -> int:
        """Same step with HOST nodal buffers (pinned torch tensors or numpy): the end-to-end C-ABI call."""
        n = C.c_int(0)
        self._check(self._L.pfem2_step_host(self._h, _hp(fx), _hp(fy), _hp(wx), _hp(wy), dt, substeps, C.byref(n)), "step_host")
        return n.value

    def set_profiling(self, enabled: bool):
        self._check(self._L.pfem2_set_profiling(self._h, 1 if enabled else 0), "set_profiling")

    def phase_times(self, reset=True) -> dict:
        """{phase: (milliseconds, launch groups)} accumulated since the last reset, from CUDA events."""
        n = len(_lib.PHASES)
        ms = (C.c_double * n)()
        calls = (C.c_longlong * n)()
        self._check(self._L.pfem2_get_phase_times(self._h, ms, calls, 1 if reset else 0), "phase_times")
        return {name: (ms[k], calls[k]) for k, name in enumerate(_lib.PHASES)}

    def stats(self) -> dict:
        s = _lib.Stats()
        self._check(self._L.pfem2_get_stats(self._h, C.byref(s)), "get_stats")
        return {k: getattr(s, k) for k, _ in _lib.Stats._fields_}

    def download(self) -> dict:
        n = self.get_particle_count()
        out = {k: np.empty(n, dtype=np.float64) for k in self.FIELDS[:7]}
        out["cell"] = np.empty(n, dtype=np.uint32)
        out["id"] = np.empty(n, dtype=np.uint32)
        self._check(self._L.pfem2_download(self._h, *[out[k].ctypes.data for k in self.FIELDS]), "download")
        return out

    def upload(self, state: dict):
        n = int(state["x"].shape[0])
        a = [np.ascontiguousarray(state[k], dtype=np.float64) for k in self.FIELDS[:7]]
        a.append(np.ascontiguousarray(state["cell"], dtype=np.uint32))
        a.append(np.ascontiguousarray(state.get("id", np.zeros(n, dtype=np.uint32)), dtype=np.uint32))
        self._check(self._L.pfem2_upload(self._h, n, *[v.ctypes.data for v in a]), "upload")

    def device_records(self) -> torch.Tensor:
        """Zero-copy (count, 8) float64 view of the 64-byte records in the sorted (physical) order: {x, y | L0, L1 | L2, (cell, id)
        | vx, vy}.  Materialises a pending permutation and applies a pending correction first; valid until the next mutating call."""
        n = self.get_particle_count()
        p = C.c_void_p()
        self._check(self._L.pfem2_device_records(self._h, C.byref(p)), "device_records")
        if not n:
            return torch.empty((0, 8), dtype=torch.float64, device=self.mesh.device)
        return _wrap_device(p.value, (n, 8), "<f8", self.mesh.device)

    def state_checksum(self) -> torch.Tensor:
        """Order-independent checksum of the particle SET, evaluated on the device: int64 tensor
        [count, sum bits(x), sum bits(y), sum bits(L0), sum bits(L1), sum bits(L2), sum cell] with wrapping sums.  Two runs hold
        the same particles in the same cells at bit-identical positions iff these agree (up to 2^-64 collisions); summing the
        tensors of the strips of a multi-GPU run gives the checksum of the global state.  Velocities and ids are left out: the
        last bits of the velocities depend on the summation order of the projection, ids of re-seeded particles are slot numbers."""
        rec = self.device_records().view(torch.int64)
        out = torch.zeros(7, dtype=torch.int64, device=self.mesh.device)
        out[0] = rec.shape[0]
        if rec.shape[0]:
            out[1:6] = rec[:, :5].sum(dim=0)
            out[6] = (rec[:, 5] & 0xFFFFFFFF).sum()
        return out

    def node_ranges(self, substeps: int):
        """(in_lo, in_hi, own_lo, own_hi): the nodes whose velocity an advect reads / whose projection this handle owns."""
        out = (C.c_int * 4)()
        self._check(self._L.pfem2_node_ranges(self._h, substeps, out), "node_ranges")
        return tuple(int(v) for v in out)

    def cell_starts(self) -> torch.Tensor:
        p = C.c_void_p()
        self._check(self._L.pfem2_cell_starts(self._h, C.byref(p)), "cell_starts")
        return _wrap_device(p.value, (self.mesh.n_cells + 1,), "<i4", self.mesh.device)


def _hp(a):
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    return a.ctypes.data


class _DevView:
    """Borrowed device memory exposed through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _wrap_device(ptr, shape, typestr, device):
    return torch.as_tensor(_DevView(ptr, shape, typestr), device=device)


def kernel_launches() -> int:
    return int(_lib.load().pfem2_kernel_launches())
