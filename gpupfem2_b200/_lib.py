"""ctypes loader of libpfem2_b200.so (the C ABI in include/pfem2_b200.h).

There is no CPU fallback: if the CUDA library is missing or fails to load, importing the handler
raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C gpupfem2_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PFEM2_LIB_PATH: A/B measurements of kernel variants built side by side (tools/); the product loads the in-tree library
LIB_PATH = os.environ.get("PFEM2_LIB_PATH") or os.path.join(_HERE, "libpfem2_b200.so")

PFEM2_OK, PFEM2_EINVAL, PFEM2_ECUDA, PFEM2_ECAPACITY, PFEM2_ESTATE = 0, 1, 2, 3, 4

# every symbol include/pfem2_b200.h declares (tests check the library exports all of them)
SYMBOLS = (
    "pfem2_default_options", "pfem2_last_error", "pfem2_version", "pfem2_create", "pfem2_destroy", "pfem2_seed",
    "pfem2_init_velocity", "pfem2_init_velocity_ptrs", "pfem2_advect", "pfem2_advect_ptrs", "pfem2_project",
    "pfem2_project_ptrs", "pfem2_correct", "pfem2_correct_ptrs", "pfem2_particle_count", "pfem2_get_stats",
    "pfem2_export_aos", "pfem2_step_host", "pfem2_download", "pfem2_upload", "pfem2_device_records", "pfem2_cell_starts",
    "pfem2_mesh_inv_jacobi", "pfem2_mesh_one_ring", "pfem2_sort_pairs", "pfem2_kernel_launches", "pfem2_set_profiling",
    "pfem2_get_phase_times", "pfem2_set_owned_cells", "pfem2_advect_move", "pfem2_emigrants_count", "pfem2_emigrants_pack",
    "pfem2_immigrants_append", "pfem2_advect_finish", "pfem2_project_accumulate", "pfem2_project_finalize",
    "pfem2_set_rank_bounds", "pfem2_emigrants_pack_neighbours", "pfem2_immigrants_append_device",
    "pfem2_p2p_inbox_create", "pfem2_p2p_connect", "pfem2_emigrants_send_p2p", "pfem2_immigrants_recv_p2p",
    "pfem2_project_halo_p2p", "pfem2_p2p_last_sent", "pfem2_project_dual", "pfem2_project_dual_ptrs", "pfem2_node_ranges", "pfem2_set_global_cell_offset", "pfem2_mesh_band", "pfem2_step_host_p2p", "pfem2_advect_p2p", "pfem2_project_p2p",
)


class MeshView(C.Structure):
    _fields_ = [("n_nodes", C.c_int), ("n_cells", C.c_int), ("d_vertices", C.c_void_p), ("d_cells", C.c_void_p),
                ("d_inv_jacobi", C.c_void_p), ("d_nbr_offsets", C.c_void_p), ("d_nbr_indices", C.c_void_p)]


class Options(C.Structure):
    _fields_ = [("struct_size", C.c_int), ("subcell_mode", C.c_int), ("max_division_level", C.c_int),
                ("capacity_factor", C.c_double), ("stream", C.c_void_p), ("device", C.c_int), ("verbose", C.c_int), ("exact_search", C.c_int), ("stable_order", C.c_int), ("reserved_scatter_tma", C.c_int), ("defer_correct", C.c_int), ("reserved_lane_per_record", C.c_int), ("host_pipeline", C.c_int), ("reserved_fuse_project", C.c_int), ("lazy_sort", C.c_int), ("graph_advect", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("count", C.c_int), ("lost", C.c_int), ("added", C.c_int), ("movers", C.c_int), ("capacity", C.c_int),
                ("overflow", C.c_int)]


PHASES = ("advect_locate", "sort", "reorder", "project_cells", "project_nodes", "correct")

_lib = None


def load():
    """Load the CUDA library; raises (never falls back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built: the PFEM-2 particle step has no CPU fallback; "
                           "run __graft_entry__.build() or make -C gpupfem2_b200/csrc")
    L = C.CDLL(LIB_PATH)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    L.pfem2_default_options.argtypes = [C.POINTER(Options)]
    L.pfem2_default_options.restype = None
    L.pfem2_last_error.argtypes = [vp]
    L.pfem2_last_error.restype = C.c_char_p
    L.pfem2_version.restype = C.c_char_p
    L.pfem2_create.argtypes = [C.POINTER(vp), C.POINTER(MeshView), i, C.POINTER(Options)]
    L.pfem2_destroy.argtypes = [vp]
    L.pfem2_seed.argtypes = [vp]
    L.pfem2_init_velocity.argtypes = [vp, vp, vp]
    L.pfem2_init_velocity_ptrs.argtypes = [vp, vp]
    L.pfem2_advect.argtypes = [vp, vp, vp, d, i]
    L.pfem2_advect_ptrs.argtypes = [vp, vp, d, i]
    L.pfem2_project.argtypes = [vp, vp, vp]
    L.pfem2_project_ptrs.argtypes = [vp, vp]
    L.pfem2_project_dual.argtypes = [vp, vp, vp, vp, vp]
    L.pfem2_project_dual_ptrs.argtypes = [vp, vp, vp]
    L.pfem2_correct.argtypes = [vp, vp, vp, vp, vp]
    L.pfem2_correct_ptrs.argtypes = [vp, vp, vp]
    L.pfem2_particle_count.argtypes = [vp, C.POINTER(i)]
    L.pfem2_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.pfem2_export_aos.argtypes = [vp, C.POINTER(vp), C.POINTER(i)]
    L.pfem2_step_host.argtypes = [vp, vp, vp, vp, vp, d, i, C.POINTER(i)]
    L.pfem2_advect_p2p.argtypes = [vp, i, vp, vp, d, i]
    L.pfem2_project_p2p.argtypes = [vp, vp, vp, vp]
    L.pfem2_step_host_p2p.argtypes = [vp, i, vp, vp, vp, vp, d, i, C.POINTER(i)]
    L.pfem2_download.argtypes = [vp] + [vp] * 9
    L.pfem2_upload.argtypes = [vp, i] + [vp] * 9
    L.pfem2_device_records.argtypes = [vp, C.POINTER(vp)]
    L.pfem2_cell_starts.argtypes = [vp, C.POINTER(vp)]
    L.pfem2_mesh_inv_jacobi.argtypes = [i, vp, vp, vp, vp]
    L.pfem2_mesh_one_ring.argtypes = [i, i, vp, vp, vp, C.POINTER(i), vp]
    L.pfem2_sort_pairs.argtypes = [i, i, vp, vp, vp, vp, C.POINTER(i), vp]
    L.pfem2_kernel_launches.restype = C.c_longlong
    L.pfem2_set_owned_cells.argtypes = [vp, i, i]
    L.pfem2_advect_move.argtypes = [vp, vp, vp, d, i]
    L.pfem2_emigrants_count.argtypes = [vp, C.POINTER(i), i, C.POINTER(i)]
    L.pfem2_emigrants_pack.argtypes = [vp, vp, C.c_longlong]
    L.pfem2_immigrants_append.argtypes = [vp, vp, i]
    L.pfem2_advect_finish.argtypes = [vp, vp, vp]
    L.pfem2_set_rank_bounds.argtypes = [vp, C.POINTER(i), i]
    L.pfem2_emigrants_pack_neighbours.argtypes = [vp, i, vp, vp, i]
    L.pfem2_immigrants_append_device.argtypes = [vp, vp, i, i]
    L.pfem2_p2p_inbox_create.argtypes = [vp, i, i, i, C.POINTER(i), vp]
    L.pfem2_p2p_connect.argtypes = [vp, i, vp]
    L.pfem2_emigrants_send_p2p.argtypes = [vp, i]
    L.pfem2_immigrants_recv_p2p.argtypes = [vp]
    L.pfem2_project_halo_p2p.argtypes = [vp, vp]
    L.pfem2_p2p_last_sent.argtypes = [vp, C.POINTER(i)]
    L.pfem2_project_accumulate.argtypes = [vp, vp]
    L.pfem2_project_finalize.argtypes = [vp, vp, vp, vp]
    L.pfem2_node_ranges.argtypes = [vp, i, C.POINTER(i)]
    L.pfem2_set_global_cell_offset.argtypes = [vp, i]
    L.pfem2_mesh_band.argtypes = [i, vp, vp, C.POINTER(i), vp]
    L.pfem2_set_profiling.argtypes = [vp, i]
    L.pfem2_get_phase_times.argtypes = [vp, C.POINTER(d), C.POINTER(C.c_longlong), i]
    _lib = L
    return L
