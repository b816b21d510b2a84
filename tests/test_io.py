"""Mesh / particle-state I/O (gpupfem2_b200/io.py): file formats on the CPU with the oracle,
restart and device-side export on the GPU."""
import os

import numpy as np
import pytest

import cases
from gpupfem2_b200 import io as pio
from gpupfem2_b200.mesh import HostMesh, load_dat, write_dat
from helpers import assert_states_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def golden_mesh(name="channel"):
    d = np.load(os.path.join(ROOT, "tests", "golden", f"mesh_{name}.npz"))
    return HostMesh(d["vertices"], d["cells"])


def test_fast_dat_reader_equals_reference_format_reader(tmp_path):
    m = golden_mesh("cylinder3")
    p = str(tmp_path / "mesh.dat")
    write_dat(p, m)
    with open(p, "a") as f:  # entities of other types are skipped, as in Mesh2D::loadMeshFromFile
        f.write("99999 102 1 2 \n")
    a, b = pio.read_dat_fast(p), load_dat(p)
    assert np.array_equal(a.cells, b.cells) and np.array_equal(a.cells, m.cells)
    assert np.array_equal(a.vertices, b.vertices)
    assert np.allclose(a.vertices, m.vertices, rtol=0, atol=1e-13)  # the DAT text carries 15 significant digits


def test_binary_mesh_roundtrip(tmp_path, oracle):
    m = oracle.complete_mesh(cases._tiny(True))
    p = str(tmp_path / "mesh.npz")
    pio.save_mesh(p, m)
    r = pio.load_mesh(p)
    for k in ("vertices", "cells", "nbr_offsets", "nbr_indices", "inv_jacobi"):
        assert np.array_equal(getattr(m, k), getattr(r, k)), k
    assert pio.mesh_fingerprint(m) == pio.mesh_fingerprint(r) != pio.mesh_fingerprint(golden_mesh())


def test_checkpoint_restart_continues_bit_identically_oracle(tmp_path, oracle):
    c = cases.build_case("tiny_l3")
    oracle.complete_mesh(c.mesh)
    a = oracle.OracleHandler(c.mesh, c.level)
    a.seed_particles()
    a.init_particle_velocity(c.fx, c.fy)
    wx, wy = np.zeros_like(c.fx), np.zeros_like(c.fx)
    for _ in range(5):
        a.step(c.fx, c.fy, wx, wy, c.dt, c.substeps)
    p = str(tmp_path / "state.npz")
    pio.checkpoint_handler(p, a, mesh=c.mesh, step=5, time=5 * c.dt)
    b = oracle.OracleHandler(c.mesh, c.level)
    meta = pio.restore(p, b, mesh=c.mesh)
    assert meta["step"] == 5 and meta["count"] == a.particle_count() and meta["level"] == c.level
    wx2, wy2 = np.zeros_like(c.fx), np.zeros_like(c.fx)
    for _ in range(5):
        na = a.step(c.fx, c.fy, wx, wy, c.dt, c.substeps)
        nb = b.step(c.fx, c.fy, wx2, wy2, c.dt, c.substeps)
        assert na == nb
    assert_states_equal(a.download(), b.download(), "restart", exact_vel=True)
    with pytest.raises(ValueError):
        pio.load_checkpoint(p, mesh=golden_mesh())  # another mesh
    st, _ = pio.load_checkpoint(p)
    st["x"] = st["x"][:-1]
    with pytest.raises(ValueError):
        pio.save_checkpoint(str(tmp_path / "bad.npz"), st, level=c.level)  # ragged state


def test_particle_vtu_ascii_and_binary(tmp_path):
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from insitu_compare import parse_vtu

    rng = np.random.default_rng(7)
    x, y, vx, vy = rng.random(37), rng.random(37), rng.standard_normal(37), rng.standard_normal(37)
    pa, pb = str(tmp_path / "a.vtu"), str(tmp_path / "b.vtu")
    pio.write_particles_vtu(pa, x, y, vx, vy)
    pio.write_particles_vtu(pb, x, y, vx, vy, binary=True)
    ra, rb = pio.read_particles_vtu(pa), pio.read_particles_vtu(pb)
    assert ra["points"].shape == rb["points"].shape == (37, 3)
    assert np.allclose(ra["points"][:, 0], x, rtol=1e-5) and np.allclose(ra["velocity"][:, 1], vy, rtol=1e-5, atol=1e-7)  # %g
    assert np.array_equal(rb["points"][:, 1], y.astype(np.float32)) and np.array_equal(rb["velocity"][:, 0], vx.astype(np.float32))
    assert np.array_equal(rb["offsets"], np.arange(1, 38)) and np.all(rb["types"] == 1)
    ref_style = parse_vtu(pa)  # the parser used on the reference's own files
    assert ref_style["velocity"].shape[0] == 3 * 37 and ref_style["connectivity"].shape[0] == 37
    pio.write_particles_vtu(str(tmp_path / "empty.vtu"), [], [], [], [])  # empty input


@pytest.mark.gpu
def test_cuda_checkpoint_restart_and_device_columns(tmp_path, oracle):
    torch = pytest.importorskip("torch")
    from gpupfem2_b200 import handler as gpu

    c = cases.build_case("cyl3_l2")
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    f = (torch.as_tensor(c.fx).cuda(), torch.as_tensor(c.fy).cuda())
    w = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    w2 = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    a = gpu.ParticleHandler2D(dm, c.level, stable_order=True)
    a.seed_particles()
    a.init_particle_velocity(f)
    for _ in range(4):
        a.step(f, w, c.dt, c.substeps)
    p = str(tmp_path / "state.npz")
    pio.checkpoint_handler(p, a, mesh=c.mesh, level=c.level, step=4)
    # device-side export: zero-copy column views of the sorted record array agree with the host download
    cols, st = pio.device_columns(a), a.download()
    assert np.array_equal(cols["pos"][:, 0].cpu().numpy(), st["x"]) and np.array_equal(cols["vel"][:, 1].cpu().numpy(), st["vy"])
    assert np.array_equal(cols["cell"].cpu().numpy().view(np.uint32), st["cell"]) and np.array_equal(cols["l2"].cpu().numpy(), st["l2"])
    assert np.all(np.diff(st["cell"].astype(np.int64)) >= 0)
    pio.write_particles_vtu(str(tmp_path / "p.vtu"), cols["pos"][:, 0].cpu(), cols["pos"][:, 1].cpu(), cols["vel"][:, 0].cpu(),
                            cols["vel"][:, 1].cpu(), binary=True)
    assert pio.read_particles_vtu(str(tmp_path / "p.vtu"))["points"].shape == (st["x"].shape[0], 3)
    b = gpu.ParticleHandler2D(dm, c.level, stable_order=True)
    pio.restore(p, b, mesh=c.mesh)
    for _ in range(4):
        a.step(f, w, c.dt, c.substeps)
        b.step(f, w2, c.dt, c.substeps)
        assert a.get_particle_count() == b.get_particle_count()
    assert_states_equal(a.download(), b.download(), "cuda restart")
    assert np.allclose(w[0].cpu().numpy(), w2[0].cpu().numpy(), rtol=1e-12, atol=0)
