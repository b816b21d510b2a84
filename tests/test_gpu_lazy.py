"""Parity tests aimed at the lazy re-sort (pfem2_options.lazy_sort, the default; gpupfem2_b200/csrc/pfem2_move.cuh k_move_gather,
pfem2_resort.cuh k_rank / k_reseed_lazy / k_materialize): both shared-memory tile layouts of the gathered move pass,
every reader of the physical order in the permuted state, growth, restart, the chunked pfem2_step_host.

Same bar as test_gpu_parity.py: the reference's CUDA dumps and the CPU oracle; owner cells / positions / local coordinates /
seed-remove sets bit-exact, velocities and projected nodal fields within 1e-12 relative."""
import numpy as np
import pytest

import cases
from helpers import REL_TOL, assert_state_matches_golden, assert_states_equal, load_golden, rel_inf
from test_gpu_parity import GOLDEN_CASES, dev_field, run_both

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    from gpupfem2_b200 import handler

    return handler


@pytest.fixture(params=["swizzle64", "linear"])
def tile_layout(request, monkeypatch):
    """Both shared-memory layouts of the lazy move pass: 64-byte swizzled tiles (default) and linear tiles (the fallback,
    PFEM2_LAZY_SWIZZLE=0, read at pfem2_create)."""
    monkeypatch.setenv("PFEM2_LAZY_SWIZZLE", "1" if request.param == "swizzle64" else "0")
    return request.param


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_lazy_matches_reference_dumps(gpu, oracle, name, tile_layout):
    """The reference's own CUDA output.  Steps between two dumps chain permuted state -> permuted state; a dump (download)
    materialises the sorted order, after which the next advect starts from the identity permutation again."""
    c = cases.build_case(name)
    g = load_golden(name)
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    h = gpu.ParticleHandler2D(dm, c.level, lazy_sort=True)
    h.seed_particles()
    f, w = dev_field(c)
    h.init_particle_velocity(f)
    steps = [int(s) for s in g["steps"]]
    counts = g["counts"]
    for s in range(1, max(max(steps), len(counts)) + 1):
        h.step(f, w, c.dt, c.substeps)
        if s <= len(counts):
            assert h.get_particle_count() == counts[s - 1], f"{name}: count after step {s}"
        if s in steps:
            assert_state_matches_golden(h.download(), w[0].cpu().numpy(), w[1].cpu().numpy(), g, s, c.full_state, name)
    h.close()


@pytest.mark.parametrize("case,level,substeps,dt,nsteps,every", [
    ("tiny_out", 4, 3, 0.2, 20, 5),      # outflow + re-seeding, partial last tile
    ("tiny_fast", 2, 1, 0.25, 8, 4),     # CFL ~ 1.2 per substep: interior deletions (lost records stay in the dense array)
    ("tiny_l5", 5, 3, 0.2, 8, 4),        # levels above the reference's cap
    ("tiny_l6", 6, 3, 0.2, 8, 8),        # 36 / cell: 64-bit occupancy masks
    ("tiny_l8", 8, 3, 0.2, 6, 3),
])
def test_lazy_matches_oracle(gpu, oracle, case, level, substeps, dt, nsteps, every, tile_layout):
    m = cases._tiny(True)
    if case == "tiny_fast":
        fx, fy = cases._mix(m, 1.0, 1.0, 0.6, 1.0)
    elif case == "tiny_out":
        fx, fy = cases._mix(m, 0.5, 1.0, 0.2, 1.0)
    else:
        fx, fy = cases._mix(m, -0.5, 1.0, 0.2, 1.0)
    h, o = run_both(gpu, oracle, m, fx, fy, level, substeps, dt, nsteps, check_every=every, lazy_sort=True, max_division_level=8)
    if case == "tiny_fast":
        assert h.stats()["lost"] > 0


def test_lazy_on_the_shipped_cylinder_mesh(gpu, oracle):
    c = cases.build_case("cyl3_l2")
    run_both(gpu, oracle, c.mesh, c.fx, c.fy, c.level, c.substeps, c.dt, 12, check_every=6, lazy_sort=True)


def test_lazy_clamped_subcell_mode_and_exact_search(gpu, oracle):
    m = cases._tiny(False)
    fx, fy = cases._mix(m, 0.5, 1.0, 0.2, 1.0)
    run_both(gpu, oracle, m, fx, fy, 3, 3, 0.2, 10, check_every=5, subcell_mode=1, lazy_sort=True)
    run_both(gpu, oracle, m, fx, fy, 3, 3, 0.2, 10, check_every=5, exact_search=True, lazy_sort=True)


def test_lazy_capacity_growth(gpu, oracle):
    """The dense array also holds the particles lost in the pass and the re-seeds are appended behind it; growth has to
    materialise the sorted order first."""
    m = cases._tiny(True)
    fx, fy = cases._mix(m, -0.5, 1.0, 0.2, 1.0)
    h, o = run_both(gpu, oracle, m, fx, fy, 2, 3, 0.2, 40, check_every=10, capacity_factor=1.05, lazy_sort=True)
    assert h.stats()["capacity"] > int(1.05 * m.n_cells * 4) + 1


def test_lazy_eager_correction_and_every_reader_of_the_physical_order(gpu, oracle):
    """defer_correct=False materialises in every step (the eager kernel walks the physical order); getParticles, device
    records + segment table and a second projection are looked at in the permuted state."""
    c = cases.build_case("channel_fast_rev")
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    ha = gpu.ParticleHandler2D(dm, c.level)
    hb = gpu.ParticleHandler2D(dm, c.level, lazy_sort=True)
    hc = gpu.ParticleHandler2D(dm, c.level, lazy_sort=True, defer_correct=False)
    f, wa = dev_field(c)
    wb = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    wc = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    for h in (ha, hb, hc):
        h.seed_particles()
        h.init_particle_velocity(f)
    for s in range(6):
        ha.step(f, wa, c.dt, c.substeps)
        hb.step(f, wb, c.dt, c.substeps)
        hc.step(f, wc, c.dt, c.substeps)
        assert ha.get_particle_count() == hb.get_particle_count() == hc.get_particle_count(), f"step {s}"
        sa, sb, sc = ha.stats(), hb.stats(), hc.stats()
        for k in ("lost", "added", "movers"):
            assert sa[k] == sb[k] == sc[k], f"step {s}: {k}"
        for w in (wb, wc):
            assert rel_inf(w[0].cpu().numpy(), wa[0].cpu().numpy()) <= REL_TOL and rel_inf(w[1].cpu().numpy(), wa[1].cpu().numpy()) <= REL_TOL
    # a second projection in the permuted state (the deferred correction is flushed eagerly: materialises)
    w2 = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    ha.project_velocity_onto_grid(wa)
    hb.project_velocity_onto_grid(w2)
    assert rel_inf(w2[0].cpu().numpy(), wa[0].cpu().numpy()) <= REL_TOL
    hb.advect_particles(f, c.dt, c.substeps)
    ha.advect_particles(f, c.dt, c.substeps)
    # permuted state: the sorted view must be materialised for these readers
    starts = hb.cell_starts().cpu().numpy()
    raw = hb.get_particles().cpu().numpy().view(np.uint8).reshape(-1, 96)
    cells = raw[:, 80:84].copy().view(np.uint32).ravel()
    assert raw.shape[0] == hb.get_particle_count() == starts[-1]
    assert np.all(np.diff(cells.astype(np.int64)) >= 0), "getParticles() not sorted by cell after a lazy advect"
    assert np.array_equal(np.diff(starts), np.bincount(cells, minlength=c.mesh.n_cells))
    assert_states_equal(ha.download(), hb.download(), "lazy vs default")


def test_lazy_step_host_and_upload(gpu, oracle):
    """pfem2_step_host (one chunk: small mesh) and a checkpoint / restart through download + upload."""
    c = cases.build_case("tiny_l3")
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    ha = gpu.ParticleHandler2D(dm, c.level)
    hb = gpu.ParticleHandler2D(dm, c.level, lazy_sort=True)
    f, w = dev_field(c)
    for h in (ha, hb):
        h.seed_particles()
        h.init_particle_velocity(f)
    hwx, hwy = np.zeros_like(c.fx), np.zeros_like(c.fx)
    for s in range(6):
        ha.step(f, w, c.dt, c.substeps)
        n = hb.step_host(c.fx, c.fy, hwx, hwy, c.dt, c.substeps)
        assert n == ha.get_particle_count(), f"step {s}"
        assert rel_inf(hwx, w[0].cpu().numpy()) <= REL_TOL and rel_inf(hwy, w[1].cpu().numpy()) <= REL_TOL
    s = hb.download()
    assert_states_equal(ha.download(), s, "lazy step_host")
    hc = gpu.ParticleHandler2D(dm, c.level, lazy_sort=True)
    rng = np.random.default_rng(11)
    perm = rng.permutation(s["x"].shape[0])
    hc.upload({k: v[perm] for k, v in s.items()})
    w2 = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    for _ in range(3):
        ha.step(f, w, c.dt, c.substeps)
        hc.step(f, w2, c.dt, c.substeps)
    assert_states_equal(ha.download(), hc.download(), "lazy after restart")


@pytest.mark.parametrize("name,chunks", [("channel_fast", 3), ("cyl3_l2", 4), ("tiny_l4", 5)])
def test_lazy_pipelined_step_host_equals_the_three_calls(gpu, oracle, name, chunks):
    """The chunked pfem2_step_host under lazy_sort: the move pass runs chunk by chunk over WHOLE tiles of the sorted order (the
    pass is not in place, so a tile must not be shared by two launches), the projection chunk by chunk through the permutation."""
    c = cases.build_case(name)
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    ha = gpu.ParticleHandler2D(dm, c.level)
    hb = gpu.ParticleHandler2D(dm, c.level, host_pipeline=chunks, lazy_sort=True)
    f, w = dev_field(c)
    for h in (ha, hb):
        h.seed_particles()
        h.init_particle_velocity(f)
    hfx, hfy = torch.as_tensor(c.fx).pin_memory(), torch.as_tensor(c.fy).pin_memory()
    hwx, hwy = torch.zeros_like(hfx).pin_memory(), torch.zeros_like(hfx).pin_memory()
    for s in range(8):
        ha.step(f, w, c.dt, c.substeps)
        n = hb.step_host(hfx, hfy, hwx, hwy, c.dt, c.substeps)
        assert n == ha.get_particle_count(), f"step {s}"
        sa, sb = ha.stats(), hb.stats()
        assert (sa["lost"], sa["added"], sa["movers"]) == (sb["lost"], sb["added"], sb["movers"]), f"step {s}"
        assert rel_inf(w[0].cpu().numpy(), hwx.numpy()) <= REL_TOL and rel_inf(w[1].cpu().numpy(), hwy.numpy()) <= REL_TOL
    assert_states_equal(ha.download(), hb.download(), name)


def test_capacity_overflow_is_reported_not_written_out_of_bounds(gpu, oracle):
    """A step that needs more rows than the capacity policy can foresee (the policy extrapolates the growth of the previous
    step): every cell holds its whole population in ONE sub-cell, so the distribution check wants 15 new particles per cell on
    top of a full array.  The plan must raise the overflow flag, the rank / re-seed kernels must not touch memory behind the
    arrays (k_reseed_lazy padded src_new[] before looking at the flag), and the next call that synchronises the
    counters must fail with PFEM2_ECAPACITY."""
    from gpupfem2_b200.mesh import structured_channel

    m = oracle.complete_mesh(structured_channel(60, 40, 6.0, 4.0, colmajor=True))
    dm = gpu.DeviceMesh(m)
    h = gpu.ParticleHandler2D(dm, 4, capacity_factor=1.05)
    h.seed_particles()
    f = (torch.zeros(m.n_nodes, dtype=torch.float64, device="cuda"), torch.zeros(m.n_nodes, dtype=torch.float64, device="cuda"))
    h.init_particle_velocity(f)
    s = h.download()
    cap = h.stats()["capacity"]
    first = np.flatnonzero(np.r_[True, np.diff(s["cell"].astype(np.int64)) != 0])  # one particle (sub-cell 0) per cell
    assert first.size == m.n_cells
    pick = first[np.arange(cap) % first.size]
    h.upload({k: v[pick] for k, v in s.items()})
    assert h.get_particle_count() == cap
    with pytest.raises(gpu.Pfem2Error, match="capacity"):
        h.advect_particles(f, 1e-9, 3)
        h.get_particle_count()
    torch.cuda.synchronize()  # no sticky CUDA error (an out-of-bounds write would surface here or in the next test)
    h.close()
    # the same device state is fine with room to grow
    h = gpu.ParticleHandler2D(dm, 4, capacity_factor=3.5)
    h.seed_particles()
    h.init_particle_velocity(f)
    h.upload({k: v[pick] for k, v in s.items()})
    h.advect_particles(f, 1e-9, 3)
    assert h.get_particle_count() == cap + 15 * m.n_cells
    h.close()


@pytest.mark.parametrize("name", ["channel_l2", "cyl3_l2"])
def test_graphed_advect_equals_plain_launches(gpu, oracle, name):
    """pfem2_options.graph_advect: advectParticles replayed as ONE CUDA-graph launch per call (automatic on the small shipped meshes).
    Same particle set bit for bit, same counters and projected field as the plain enqueue sequence, across everything that
    invalidates or bypasses a graph: the first (physical-state) call, a changed time step (re-capture), a reader of the physical
    order in between (materialise -> plain call -> capture again), an eager correction, capacity growth, and both pointer flavours."""
    c = cases.build_case(name)
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    hg = gpu.ParticleHandler2D(dm, c.level, graph_advect=1, capacity_factor=1.06)
    hp = gpu.ParticleHandler2D(dm, c.level, graph_advect=-1, capacity_factor=1.06)
    f, wg = dev_field(c)
    wp = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    table = torch.tensor([f[0].data_ptr(), f[1].data_ptr()], dtype=torch.int64, device="cuda")
    for h in (hg, hp):
        h.seed_particles()
        h.init_particle_velocity(f)
    for s in range(24):
        dt = c.dt if s < 14 else 0.5 * c.dt  # a new time step: the graphs are captured again
        for h, w in ((hg, wg), (hp, wp)):
            if s % 5 == 4:
                h.advect_particles_ptrs(table, dt, c.substeps)  # deviceVector<double*>::data flavour
            else:
                h.advect_particles(f, dt, c.substeps)
            h.project_velocity_onto_grid(w)
            h.correct_particle_velocity(f, w)
        assert hg.get_particle_count() == hp.get_particle_count(), f"step {s}"
        sg, sp = hg.stats(), hp.stats()
        assert (sg["lost"], sg["added"], sg["movers"]) == (sp["lost"], sp["added"], sp["movers"]), f"step {s}"
        assert rel_inf(wg[0].cpu().numpy(), wp[0].cpu().numpy()) <= REL_TOL and rel_inf(wg[1].cpu().numpy(), wp[1].cpu().numpy()) <= REL_TOL
        if s == 8:  # a reader of the physical order: the next advect starts from the identity permutation, without a graph
            assert_states_equal(hg.download(), hp.download(), f"{name} step {s}")
        if s == 11:  # getParticles applies the deferred correction eagerly: the next graph is captured without it
            assert hg.get_particles().shape == hp.get_particles().shape
    assert hg.stats()["capacity"] == hp.stats()["capacity"]
    assert_states_equal(hg.download(), hp.download(), name)
    hg.close()
    hp.close()
