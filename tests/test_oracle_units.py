"""Unit tests of the oracle's element-wise restatements and mesh helpers (CPU)."""
import numpy as np
import pytest

import cases
from gpupfem2_b200.mesh import HostMesh, structured_channel


def brute_force_ring(cells, c):
    """Definition of Mesh2D::fillCellNeighborIndices (mesh_2d.cu:112-123): share at least one vertex."""
    sh = np.isin(cells, cells[c]).any(axis=1)
    sh[c] = False
    return np.nonzero(sh)[0]


@pytest.mark.parametrize("mesh_name", ["tiny", "channel"])
def test_one_ring_matches_definition(oracle, mesh_name):
    m = structured_channel(7, 5, 1.4, 1.0, colmajor=True) if mesh_name == "tiny" else cases._fixture_mesh("channel")
    off, idx = oracle.one_ring(m.n_nodes, m.cells)
    cells = m.cells.astype(np.int64)
    step = 1 if mesh_name == "tiny" else 53
    for c in range(0, m.n_cells, step):
        ref = brute_force_ring(cells, c)
        got = idx[off[c]:off[c + 1]]
        assert np.array_equal(ref, got)
        assert np.all(np.diff(got) > 0)  # ascending, the order the locate tie-break depends on


def test_subcell_centres(oracle):
    m = oracle.complete_mesh(structured_channel(2, 2, 1.0, 1.0))
    for level in (1, 2, 3, 4):
        h = oracle.OracleHandler(m, level)
        cen = h.subcell_centers()
        assert cen.shape == (level * level, 3)
        assert np.allclose(cen.sum(axis=1), 1.0, atol=1e-15)
        assert np.all(cen > 0)
        # every centre falls into its own sub-cell (so a freshly seeded cell is fully occupied)
        subs = [oracle.lib().orc_subcell(oracle._d(np.ascontiguousarray(cen[s])), level) for s in range(level * level)]
        assert subs == list(range(level * level))
    # clamped to MAX_CELL_DIVISION_LEVEL = 4 (particle_handler_2d.cu:241) and to >= 1
    assert oracle.OracleHandler(m, 9).particles_per_cell == 16
    assert oracle.OracleHandler(m, 0).particles_per_cell == 1
    assert oracle.OracleHandler(m, 6, max_level=8).particles_per_cell == 36


def test_inside_tolerance_band(oracle):
    L = oracle.lib()

    def inside(a, b):
        v = np.array([a, b, 1.0 - a - b])
        return bool(L.orc_inside(oracle._d(v)))

    assert inside(0.2, 0.3)
    assert inside(-1.9e-6, 0.3)       # within DOUBLE_MIN = 2e-6 (constants.h:5)
    assert not inside(-2.1e-6, 0.3)
    assert inside(1.0, 0.0)
    assert not inside(1.0000021, -1e-7)
    v = np.array([np.nan, 0.1, 0.1])
    assert bool(L.orc_inside(oracle._d(v)))  # NaN compares false everywhere -> "inside" (geometry.cuh:37-46)


def test_subcell_spill_in_tolerance_band(oracle):
    """for Ly in [-2e-6, 0] the row index is n, so the flat index leaves the cell's own counters."""
    L = oracle.lib()
    for level in (1, 2, 4):
        v = np.array([0.3, -1e-6, 0.7 + 1e-6])
        assert L.orc_subcell(oracle._d(v), level) >= level * level
        v = np.array([0.3, 1e-6, 0.7 - 1e-6])
        assert 0 <= L.orc_subcell(oracle._d(v), level) < level * level


def test_inv_jacobi_is_inverse(oracle):
    m = cases._fixture_mesh("channel")
    J = oracle.inv_jacobi(m.vertices, m.cells).reshape(-1, 2, 2)
    v = m.vertices
    c = m.cells.astype(np.int64)
    A = np.stack([v[c[:, 0]] - v[c[:, 2]], v[c[:, 1]] - v[c[:, 2]]], axis=1)  # rows v31, v32 (mesh_2d.cu:26-31)
    eye = np.einsum("cij,cjk->cik", A, J)
    assert np.allclose(eye, np.eye(2)[None], atol=1e-9)


def test_to_local_roundtrip(oracle):
    m = oracle.complete_mesh(structured_channel(3, 2, 1.5, 1.0))
    h = oracle.OracleHandler(m, 3)
    h.seed_particles()
    s = h.download()
    L = oracle.lib()
    out = np.zeros(3)
    for p in range(0, s["x"].shape[0], 5):
        c = int(s["cell"][p])
        v3 = np.ascontiguousarray(m.vertices[m.cells[c, 2]])
        L.orc_to_local(oracle._d(np.ascontiguousarray(m.inv_jacobi[c])), oracle._d(v3), s["x"][p], s["y"][p], oracle._d(out))
        assert np.allclose(out, [s["l0"][p], s["l1"][p], s["l2"][p]], atol=1e-14)
        assert L.orc_inside(oracle._d(out))


def test_empty_field_and_zero_velocity(oracle):
    """Zero nodal field: nothing moves, nothing is lost or added, projection of zero velocity is zero."""
    m = oracle.complete_mesh(structured_channel(4, 3, 1.0, 1.0))
    h = oracle.OracleHandler(m, 2)
    n0 = h.seed_particles()
    z = np.zeros(m.n_nodes)
    before = h.download()
    wx, wy = np.ones(m.n_nodes), np.ones(m.n_nodes)
    assert h.step(z, z, wx, wy, 0.1, 3) == n0
    after = h.download()
    for k in ("x", "y", "cell"):
        assert np.array_equal(before[k], after[k])
    assert np.all(wx == 0) and np.all(wy == 0)
    lost, added = h.last_stats()
    assert lost.sum() == 0 and added == 0
