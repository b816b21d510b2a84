"""Definitions of the parity cases shared by the golden-vector generator and the tests.

Each case = mesh + frozen nodal field F + (level, substeps, dt, nsteps, dump_steps), run with the
isolated step protocol of oracle/ref_harness.cu:  advect(F) ; project(W) ; correct(F, W).
No RNG anywhere: seeding is the reference's deterministic sub-cell centres.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from gpupfem2_b200.mesh import HostMesh, poiseuille_field, structured_channel, vortex_field

HERE = os.path.dirname(os.path.abspath(__file__))


@dataclass
class Case:
    name: str
    mesh: HostMesh
    fx: np.ndarray
    fy: np.ndarray
    level: int
    substeps: int
    dt: float
    nsteps: int
    dump_steps: tuple
    full_state: bool = True  # False: golden keeps (cell, x, y, vx, vy) only


def _fixture_mesh(name):
    d = np.load(os.path.join(HERE, f"mesh_{name}.npz"))
    return HostMesh(d["vertices"], d["cells"], meta={"fixture": name})


def _tiny(colmajor):
    return structured_channel(12, 6, 2.0, 1.0, colmajor=colmajor)


def boundary_nodes(mesh):
    """Boolean mask of nodes lying on an edge that belongs to exactly one triangle."""
    c = mesh.cells.astype(np.int64)
    e = np.concatenate([c[:, [0, 1]], c[:, [1, 2]], c[:, [2, 0]]])
    e.sort(axis=1)
    key = e[:, 0] * mesh.n_nodes + e[:, 1]
    uniq, cnt = np.unique(key, return_counts=True)
    b = uniq[cnt == 1]
    mask = np.zeros(mesh.n_nodes, dtype=bool)
    mask[b // mesh.n_nodes] = True
    mask[b % mesh.n_nodes] = True
    return mask


def _mix(mesh, umax, height, amp, wavelength):
    px, py = poiseuille_field(mesh, umax, height)
    vx, vy = vortex_field(mesh, 0.0, amp, wavelength)
    return np.ascontiguousarray(px + vx), np.ascontiguousarray(py + vy)


def build_case(name: str) -> Case:
    """Flow direction note: the reference's kDeleteParticles (particle_handler_2d.cu:164-171) races when a
    to-be-deleted particle sits in the last n array slots and then keeps a deleted particle
    and drops a valid one.  With x-major numbering the outflow end x = lx holds the LAST cells, i.e. the
    array tail, so the pinning cases run the flow towards x = 0 (negative umax) or in a closed box, where
    the reference follows its own rule and its output is deterministic."""
    if name == "tiny_l2":  # row-major numbering, flow towards x = 0
        m = _tiny(False)
        fx, fy = _mix(m, -0.5, 1.0, 0.2, 1.0)
        return Case(name, m, fx, fy, 2, 3, 0.2, 40, (0, 1, 2, 5, 10, 20, 40))
    if name in ("tiny_l1", "tiny_l3", "tiny_l4"):
        m = _tiny(True)
        fx, fy = _mix(m, -0.5, 1.0, 0.2, 1.0)
        return Case(name, m, fx, fy, int(name[-1]), 3, 0.2, 20, (0, 1, 5, 20))
    if name == "tiny_box":  # closed box: vortices only, no normal velocity on any wall -> no deletions
        m = _tiny(True)
        fx, fy = vortex_field(m, 0.0, 0.5, 1.0)
        return Case(name, m, fx, fy, 2, 3, 0.2, 30, (1, 10, 30))
    if name == "tiny_fast":  # CFL ~ 0.6 per substep, S = 1: jumps beyond the one-ring are deleted inside the
        # domain.  Kept below the reference's fixed side buffers (P0/10 delete slots, :307-308).
        m = _tiny(True)
        fx, fy = _mix(m, -0.5, 1.0, 0.3, 1.0)
        return Case(name, m, fx, fy, 2, 1, 0.2, 10, tuple(range(0, 11)))
    if name == "channel_l2":  # BASELINE.json config 1 (isolated mode): shipped mesh, level 2, dt 0.01, S 3
        m = _fixture_mesh("channel")
        y = m.vertices[:, 1]
        fx = np.ascontiguousarray(0.5 * y * (1.0 - y))
        return Case(name, m, fx, np.zeros_like(fx), 2, 3, 0.01, 50, (1, 50), full_state=False)
    if name in ("channel_fast", "channel_fast_rev"):  # 20x the velocity plus vortices: many cell changes
        m = _fixture_mesh("channel")
        y = m.vertices[:, 1]
        sgn = -1.0 if name.endswith("rev") else 1.0
        vx, vy = vortex_field(m, 0.0, 0.5, 0.5)
        fx = np.ascontiguousarray(sgn * 10.0 * y * (1.0 - y) + vx)
        return Case(name, m, fx, np.ascontiguousarray(vy), 2, 3, 0.01, 30, (30,), full_state=False)
    if name in ("cyl3_l2", "cyl3_l2_rev"):  # BASELINE.json config 2 mesh (isolated mode): channel profile, body ignored
        m = _fixture_mesh("cylinder3")
        fx, fy = poiseuille_field(m, -1.5 if name.endswith("rev") else 1.5, 0.41)
        return Case(name, m, fx, fy, 2, 3, 0.001, 60, (60,), full_state=False)
    if name == "cyl3_box":  # Cylinder mesh, vortices with zero velocity on every boundary node: nothing can leave
        m = _fixture_mesh("cylinder3")
        vx, vy = vortex_field(m, 0.0, 1.5, 0.2)
        keep = ~boundary_nodes(m)
        return Case(name, m, np.ascontiguousarray(vx * keep), np.ascontiguousarray(vy * keep), 2, 3, 0.001, 60, (60,),
                    full_state=False)
    raise KeyError(name)


ALL_CASES = ("tiny_l2", "tiny_l1", "tiny_l3", "tiny_l4", "tiny_box", "tiny_fast", "channel_l2", "channel_fast",
             "channel_fast_rev", "cyl3_l2", "cyl3_l2_rev", "cyl3_box")
