"""GPU parity tests proper: the CUDA path (through the C ABI) against
  (a) the reference's own CUDA output (golden dumps, tests/golden/ref_*.npz),
  (b) the CPU oracle on the same seeded inputs,
  (c) the live reference binary when oracle/_ref/ref_harness travelled to the box.
Bar: owner cells / positions / local coordinates / seed-remove sets bit-exact; velocities and projected
nodal fields within 1e-12 relative (helpers.REL_TOL)."""
import os
import subprocess

import numpy as np
import pytest

import cases
from helpers import REL_TOL, assert_state_matches_golden, assert_states_equal, load_golden, rel_inf

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests")
    from gpupfem2_b200 import handler

    return handler


def dev_field(c, dev="cuda:0"):
    f = (torch.as_tensor(c.fx).to(dev), torch.as_tensor(c.fy).to(dev))
    w = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    return f, w


GOLDEN_CASES = ["tiny_l1", "tiny_l2", "tiny_l3", "tiny_l4", "tiny_box", "tiny_fast", "channel_l2", "channel_fast_rev", "cyl3_box"]


RESORT_MODES = {"lazy": {}, "physical": {"lazy_sort": False}, "stable_order": {"stable_order": True}}


@pytest.mark.parametrize("mode", list(RESORT_MODES))
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cuda_matches_reference_dumps(gpu, oracle, name, mode):
    """The three ways the sorted order is kept: the lazy re-sort (default: permutation over a dense array, pfem2_move.cuh),
    the physical counting-sort scatter of every advect (lazy_sort=0) and the deterministic stable order (radix-sorted movers)."""
    c = cases.build_case(name)
    g = load_golden(name)
    oracle.complete_mesh(c.mesh)  # one-ring from the validated O(C) builder; invJ recomputed on the device below
    c.mesh.inv_jacobi = None
    dm = gpu.DeviceMesh(c.mesh)
    if "invj" in g:
        assert np.array_equal(dm.inv_jacobi.cpu().numpy(), g["invj"]), "device inverse Jacobians differ from the reference's"
    h = gpu.ParticleHandler2D(dm, c.level, **RESORT_MODES[mode])
    h.seed_particles()
    assert h.get_particle_count() == c.mesh.n_cells * c.level * c.level
    f, w = dev_field(c)
    h.init_particle_velocity(f)
    steps = [int(s) for s in g["steps"]]
    counts = g["counts"]
    if 0 in steps:
        assert_state_matches_golden(h.download(), None, None, g, 0, c.full_state, name)
    for s in range(1, max(max(steps), len(counts)) + 1):
        h.step(f, w, c.dt, c.substeps)
        if s <= len(counts):
            assert h.get_particle_count() == counts[s - 1], f"{name}: count after step {s}"
        if s in steps:
            assert_state_matches_golden(h.download(), w[0].cpu().numpy(), w[1].cpu().numpy(), g, s, c.full_state, name)
    h.close()


def run_both(gpu, oracle, mesh, fx, fy, level, substeps, dt, nsteps, check_every=1, **opts):
    oracle.complete_mesh(mesh)
    dm = gpu.DeviceMesh(mesh)
    h = gpu.ParticleHandler2D(dm, level, **opts)
    o = oracle.OracleHandler(mesh, level, max_level=opts.get("max_division_level", 4), subcell_mode=opts.get("subcell_mode", 0))
    opts = dict(opts)
    h.seed_particles()
    o.seed_particles()
    f = (torch.as_tensor(fx).cuda(), torch.as_tensor(fy).cuda())
    w = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    wx, wy = np.zeros_like(fx), np.zeros_like(fx)
    h.init_particle_velocity(f)
    o.init_particle_velocity(fx, fy)
    assert_states_equal(h.download(), o.download(), "after seed", exact_vel=True)
    for s in range(1, nsteps + 1):
        h.step(f, w, dt, substeps)
        n = o.step(fx, fy, wx, wy, dt, substeps)
        assert h.get_particle_count() == n, f"step {s}: count {h.get_particle_count()} vs oracle {n}"
        st = h.stats()
        lost, added = o.last_stats()
        assert st["lost"] == lost.sum() and st["added"] == added, f"step {s}: lost/added {st} vs {lost.sum()}/{added}"
        if s % check_every == 0 or s == nsteps:
            assert_states_equal(h.download(), o.download(), f"step {s}")
            assert rel_inf(w[0].cpu().numpy(), wx) <= REL_TOL and rel_inf(w[1].cpu().numpy(), wy) <= REL_TOL
    return h, o


@pytest.mark.parametrize("level", [5, 6, 8])
def test_extended_division_levels_match_oracle(gpu, oracle, level):
    """Levels above the reference's cap of 4 (BASELINE.json's 32/cell is bracketed by levels 5 and 6)."""
    m = cases._tiny(True)
    fx, fy = cases._mix(m, -0.5, 1.0, 0.2, 1.0)
    run_both(gpu, oracle, m, fx, fy, level, 3, 0.2, 8, max_division_level=8)


def test_clamped_subcell_mode_matches_oracle(gpu, oracle):
    m = cases._tiny(False)
    fx, fy = cases._mix(m, 0.5, 1.0, 0.2, 1.0)
    run_both(gpu, oracle, m, fx, fy, 3, 3, 0.2, 10, subcell_mode=1)
    run_both(gpu, oracle, m, fx, fy, 3, 3, 0.2, 10, subcell_mode=1, stable_order=True)


def test_eager_and_deferred_correction_give_identical_bits(gpu, oracle):
    """defer_correct folds correctParticleVelocity into the next advect pass; the particle velocities, the exported
    AoS view and the projected field must be bit-identical to the eager kernel at every point a caller can look."""
    c = cases.build_case("tiny_l3")
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    ha = gpu.ParticleHandler2D(dm, c.level, stable_order=True, defer_correct=True)
    hb = gpu.ParticleHandler2D(dm, c.level, stable_order=True, defer_correct=False)
    f, w = dev_field(c)
    w2 = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    for h in (ha, hb):
        h.seed_particles()
        h.init_particle_velocity(f)
    for s in range(8):
        ha.step(f, w, c.dt, c.substeps)
        hb.step(f, w2, c.dt, c.substeps)
        if s == 2:  # look at the particles right after a correct call: the deferred update must be flushed
            assert np.array_equal(ha.get_particles().cpu().numpy(), hb.get_particles().cpu().numpy())
        if s == 4:  # two corrections in a row, then a projection
            ha.correct_particle_velocity(f, w)
            hb.correct_particle_velocity(f, w2)
            ha.project_velocity_onto_grid(w)
            hb.project_velocity_onto_grid(w2)
            assert np.array_equal(w[0].cpu().numpy(), w2[0].cpu().numpy())
    a, b = ha.download(), hb.download()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(w[0].cpu().numpy(), w2[0].cpu().numpy()) and np.array_equal(w[1].cpu().numpy(), w2[1].cpu().numpy())


def test_high_cfl_interior_deletions_match_oracle(gpu, oracle):
    """CFL ~ 1.2 per substep: jumps beyond the one-ring are deleted inside the domain."""
    m = cases._tiny(True)
    fx, fy = cases._mix(m, 1.0, 1.0, 0.6, 1.0)
    h, o = run_both(gpu, oracle, m, fx, fy, 2, 1, 0.25, 8)
    assert h.stats()["lost"] > 0


def test_outflow_with_array_tail_deletions_follows_the_rule(gpu, oracle):
    """Flow towards the LAST cells: here the reference's delete race (N3) fires, the CUDA path must follow
    the rule {no accepting cell in own ∪ one-ring} exactly, like the oracle."""
    m = cases._tiny(True)
    fx, fy = cases._mix(m, 0.5, 1.0, 0.2, 1.0)
    run_both(gpu, oracle, m, fx, fy, 4, 3, 0.2, 20, check_every=5)


def test_shipped_cylinder_mesh_with_body_and_outflow(gpu, oracle):
    c = cases.build_case("cyl3_l2")
    run_both(gpu, oracle, c.mesh, c.fx, c.fy, c.level, c.substeps, c.dt, 12, check_every=6)


def test_sorted_storage_invariants(gpu, oracle):
    """After every advect the arrays are physically sorted by owning cell and cell_starts delimits the segments."""
    c = cases.build_case("channel_fast_rev")
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    h = gpu.ParticleHandler2D(dm, c.level)
    h.seed_particles()
    f, w = dev_field(c)
    h.init_particle_velocity(f)
    for _ in range(5):
        h.step(f, w, c.dt, c.substeps)
    s = h.download()
    assert np.all(np.diff(s["cell"].astype(np.int64)) >= 0), "cell keys not sorted"
    starts = h.cell_starts().cpu().numpy()
    assert starts[0] == 0 and starts[-1] == s["cell"].shape[0]
    counts = np.bincount(s["cell"], minlength=c.mesh.n_cells)
    assert np.array_equal(np.diff(starts), counts)
    # stable: within a cell, survivors precede this step's re-seeded particles, which sit at sub-cell centres
    h.close()


def test_reference_aos_layout(gpu, oracle):
    """getParticles(): 96-byte Particle2D records, ID@0 position@16 localPosition@32 velocity@64 cellID@80."""
    c = cases.build_case("tiny_l2")
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    h = gpu.ParticleHandler2D(dm, c.level)
    h.seed_particles()
    f, w = dev_field(c)
    h.init_particle_velocity(f)
    h.step(f, w, c.dt, c.substeps)
    s = h.download()
    raw = h.get_particles().cpu().numpy().view(np.uint8).reshape(-1, 96)
    assert raw.shape[0] == s["x"].shape[0]
    col = lambda off, dt: raw[:, off:off + np.dtype(dt).itemsize].copy().view(dt).ravel()
    assert np.array_equal(col(0, np.uint32), s["id"])
    assert np.array_equal(col(16, np.float64), s["x"]) and np.array_equal(col(24, np.float64), s["y"])
    assert np.array_equal(col(32, np.float64), s["l0"]) and np.array_equal(col(40, np.float64), s["l1"])
    assert np.array_equal(col(48, np.float64), s["l2"])
    assert np.array_equal(col(64, np.float64), s["vx"]) and np.array_equal(col(72, np.float64), s["vy"])
    assert np.array_equal(col(80, np.uint32), s["cell"])


def test_pointer_table_flavour_and_step_host(gpu, oracle):
    """deviceVector<double*>::data style arguments and the host-buffer step give the same bits as the plain calls."""
    c = cases.build_case("tiny_l3")
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    hs = [gpu.ParticleHandler2D(dm, c.level, stable_order=True) for _ in range(3)]
    f, w = dev_field(c)
    w2 = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    tf = torch.tensor([f[0].data_ptr(), f[1].data_ptr()], dtype=torch.int64, device="cuda")
    tw = torch.tensor([w2[0].data_ptr(), w2[1].data_ptr()], dtype=torch.int64, device="cuda")
    for h in hs:
        h.seed_particles()
    hs[0].init_particle_velocity(f)
    hs[1].init_particle_velocity_ptrs(tf)
    hs[2].init_particle_velocity(f)
    hwx, hwy = np.zeros_like(c.fx), np.zeros_like(c.fx)
    for _ in range(6):
        hs[0].step(f, w, c.dt, c.substeps)
        hs[1].advect_particles_ptrs(tf, c.dt, c.substeps)
        hs[1].project_velocity_onto_grid_ptrs(tw)
        hs[1].correct_particle_velocity_ptrs(tf, tw)
        n = hs[2].step_host(c.fx, c.fy, hwx, hwy, c.dt, c.substeps)
        assert n == hs[0].get_particle_count() == hs[1].get_particle_count()
    a, b, d = hs[0].download(), hs[1].download(), hs[2].download()
    for k in a:
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], d[k]), k  # same kernels, same order -> same bits
    assert np.array_equal(w[0].cpu().numpy(), w2[0].cpu().numpy()) and np.array_equal(w[0].cpu().numpy(), hwx)


def test_project_dual_writes_both_destinations(gpu, oracle):
    """Extension pfem2_project_dual(_ptrs): the projection plus the cases' copy into the "old" solution in one node pass."""
    c = cases.build_case("tiny_l3")
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    ha, hb = gpu.ParticleHandler2D(dm, c.level, stable_order=True), gpu.ParticleHandler2D(dm, c.level, stable_order=True)
    f, w = dev_field(c)
    w2, old2 = (torch.zeros_like(f[0]), torch.zeros_like(f[0])), (torch.full_like(f[0], 7.0), torch.full_like(f[0], 7.0))
    w3, old3 = (torch.zeros_like(f[0]), torch.zeros_like(f[0])), (torch.full_like(f[0], 7.0), torch.full_like(f[0], 7.0))
    t3 = torch.tensor([w3[0].data_ptr(), w3[1].data_ptr()], dtype=torch.int64, device="cuda")
    o3 = torch.tensor([old3[0].data_ptr(), old3[1].data_ptr()], dtype=torch.int64, device="cuda")
    for h in (ha, hb):
        h.seed_particles()
        h.init_particle_velocity(f)
    for _ in range(3):
        ha.advect_particles(f, c.dt, c.substeps)
        hb.advect_particles(f, c.dt, c.substeps)
        ha.project_velocity_onto_grid(w)
        hb.project_velocity_onto_grid_dual(w2, old2)
        hb.project_velocity_onto_grid_dual_ptrs(t3, o3)
        for k in range(2):
            ref = w[k].cpu().numpy()
            for other in (w2[k], old2[k], w3[k], old3[k]):
                assert np.array_equal(ref, other.cpu().numpy(), equal_nan=True)
        ha.correct_particle_velocity(f, w)
        hb.correct_particle_velocity(f, old2)
    assert_states_equal(ha.download(), hb.download(), "project_dual")


def test_download_upload_roundtrip_and_resort(gpu, oracle):
    """Checkpoint/restart: a shuffled upload is re-sorted by cell and continues to the same state."""
    c = cases.build_case("tiny_l2")
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    h1, h2 = gpu.ParticleHandler2D(dm, c.level), gpu.ParticleHandler2D(dm, c.level)
    f, w = dev_field(c)
    h1.seed_particles()
    h1.init_particle_velocity(f)
    for _ in range(4):
        h1.step(f, w, c.dt, c.substeps)
    s = h1.download()
    rng = np.random.default_rng(7)
    perm = rng.permutation(s["x"].shape[0])
    h2.upload({k: v[perm] for k, v in s.items()})
    t = h2.download()
    assert np.all(np.diff(t["cell"].astype(np.int64)) >= 0)
    assert_states_equal(s, t, "after upload", exact_vel=True)
    for k in s:  # upload is a stable sort by cell of what was uploaded
        assert np.array_equal(t[k], s[k][perm][np.argsort(s["cell"][perm], kind="stable")]), k
    w2 = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    for _ in range(3):
        h1.step(f, w, c.dt, c.substeps)
        h2.step(f, w2, c.dt, c.substeps)
    assert_states_equal(h1.download(), h2.download(), "after restart")


def test_device_one_ring_and_inv_jacobi_match_oracle(gpu, oracle):
    for m in (cases._fixture_mesh("channel"), cases._fixture_mesh("cylinder3"), cases._tiny(True)):
        off, idx = oracle.one_ring(m.n_nodes, m.cells)
        cells = torch.as_tensor(m.cells.view(np.int32)).cuda()
        doff, didx = gpu.device_one_ring(m.n_nodes, cells)
        assert np.array_equal(doff.cpu().numpy(), off) and np.array_equal(didx.cpu().numpy(), idx)
        dm = gpu.DeviceMesh(m)
        assert np.array_equal(dm.inv_jacobi.cpu().numpy(), oracle.inv_jacobi(m.vertices, m.cells))


@pytest.mark.parametrize("n,bits", [(0, 8), (1, 1), (255, 8), (4097, 13), (100003, 17), (1 << 20, 24), (300000, 32)])
def test_radix_sort_pairs(gpu, n, bits):
    """Stable LSD radix sort of (cell key, particle index) pairs against numpy's stable argsort."""
    import ctypes as C

    from gpupfem2_b200 import _lib

    rng = np.random.default_rng(n + bits)
    keys = rng.integers(0, 1 << bits, size=n, dtype=np.uint64).astype(np.uint32)
    if n > 10:
        keys[: n // 3] = keys[0]  # heavy duplicates: stability matters
    vals = np.arange(n, dtype=np.uint32)
    k = [torch.as_tensor(keys.view(np.int32)).cuda() if n else torch.empty(1, dtype=torch.int32, device="cuda"),
         torch.empty(max(n, 1), dtype=torch.int32, device="cuda")]
    v = [torch.as_tensor(vals.view(np.int32)).cuda() if n else torch.empty(1, dtype=torch.int32, device="cuda"),
         torch.empty(max(n, 1), dtype=torch.int32, device="cuda")]
    flip = C.c_int(0)
    rc = _lib.load().pfem2_sort_pairs(n, bits, k[0].data_ptr(), v[0].data_ptr(), k[1].data_ptr(), v[1].data_ptr(), C.byref(flip), None)
    assert rc == 0
    order = np.argsort(keys, kind="stable")
    got_k = k[flip.value].cpu().numpy().view(np.uint32)[:n]
    got_v = v[flip.value].cpu().numpy().view(np.uint32)[:n]
    assert np.array_equal(got_k, keys[order]) and np.array_equal(got_v, vals[order])


def test_capacity_growth(gpu, oracle):
    """Re-seeding only ever adds particles; storage must grow transparently."""
    m = cases._tiny(True)
    fx, fy = cases._mix(m, -0.5, 1.0, 0.2, 1.0)
    h, o = run_both(gpu, oracle, m, fx, fy, 2, 3, 0.2, 40, check_every=10, capacity_factor=1.05)
    assert h.stats()["capacity"] > int(1.05 * m.n_cells * 4) + 1


def test_live_reference_binary(gpu, oracle, tmp_path):
    """When the reference build travelled to the box (oracle/_ref/), run it here and compare directly."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_harness not built (needs /root/reference at build time)")
    from gpupfem2_b200.casefile import read_dump, write_case

    c = cases.build_case("tiny_box")
    oracle.complete_mesh(c.mesh)
    case_path = str(tmp_path / "case.bin")
    write_case(case_path, c.mesh, c.fx, c.fy, c.level, c.substeps, c.dt, 12, (12,))
    subprocess.run([exe, "dump", case_path, str(tmp_path / "ref")], check=True, stdout=subprocess.DEVNULL, timeout=300)
    ref = read_dump(str(tmp_path / "ref_step00012.bin"))
    c.mesh.inv_jacobi = np.fromfile(str(tmp_path / "ref_invj.bin"), dtype=np.float64).reshape(-1, 4)
    dm = gpu.DeviceMesh(c.mesh)
    h = gpu.ParticleHandler2D(dm, c.level)
    h.seed_particles()
    f, w = dev_field(c)
    h.init_particle_velocity(f)
    for _ in range(12):
        h.step(f, w, c.dt, c.substeps)
    mine = h.download()
    assert_states_equal(mine, ref, "vs live reference")
    assert rel_inf(w[0].cpu().numpy(), ref["wx"]) <= REL_TOL and rel_inf(w[1].cpu().numpy(), ref["wy"]) <= REL_TOL


def test_device_channel_generator_matches_host(gpu, oracle):
    from gpupfem2_b200.mesh import structured_channel

    for colmajor in (True, False):
        hm = oracle.complete_mesh(structured_channel(9, 5, 1.8, 1.0, colmajor=colmajor))
        dm = gpu.device_structured_channel(9, 5, 1.8, 1.0, colmajor=colmajor)
        assert np.array_equal(dm.vertices.cpu().numpy(), hm.vertices)
        assert np.array_equal(dm.cells.cpu().numpy().view(np.uint32), hm.cells)
        assert np.array_equal(dm.nbr_offsets.cpu().numpy(), hm.nbr_offsets)
        assert np.array_equal(dm.nbr_indices.cpu().numpy(), hm.nbr_indices)
        assert np.array_equal(dm.inv_jacobi.cpu().numpy(), hm.inv_jacobi)


@pytest.mark.parametrize("name", ["cyl3_l2", "channel_fast", "tiny_fast"])
def test_walk_fast_path_equals_ordered_scan(gpu, oracle, name):
    """The edge-walk locate (default) and the reference's ordered one-ring scan (exact_search) give identical
    bits on graded unstructured meshes with bodies, outflow and interior deletions."""
    c = cases.build_case(name)
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    ha = gpu.ParticleHandler2D(dm, c.level, stable_order=True)
    hb = gpu.ParticleHandler2D(dm, c.level, exact_search=True, stable_order=True)
    f, w = dev_field(c)
    w2 = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    for h in (ha, hb):
        h.seed_particles()
        h.init_particle_velocity(f)
    for s in range(10):
        ha.step(f, w, c.dt, c.substeps)
        hb.step(f, w2, c.dt, c.substeps)
    a, b = ha.download(), hb.download()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert ha.stats() == hb.stats()
    assert np.array_equal(w[0].cpu().numpy(), w2[0].cpu().numpy())


@pytest.mark.parametrize("name", ["cyl3_l2", "channel_fast", "tiny_fast", "tiny_l4"])
def test_lazy_resort_equals_physical_resort(gpu, oracle, name):
    """The default (records move once per step: gathered move pass through the permutation, rank pass, appended re-seeds)
    against the physical re-sort of every advect (in-place move pass, counting-sort scatter with four lanes per record):
    same particle set bit for bit and the same counters in every step, on meshes whose particle counts are not multiples of the
    32-record tile (partial last tile, tensor-map out-of-bounds rows) and with deletions, re-seeding and a pending deferred
    correction in every advect.  (Fast order on both sides: the slot order inside a cell is not deterministic, so the projected
    field is compared within the tolerance.)"""
    c = cases.build_case(name)
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    ha = gpu.ParticleHandler2D(dm, c.level)
    hb = gpu.ParticleHandler2D(dm, c.level, lazy_sort=False)
    f, w = dev_field(c)
    w2 = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    for h in (ha, hb):
        h.seed_particles()
        h.init_particle_velocity(f)
    for s in range(10):
        ha.step(f, w, c.dt, c.substeps)
        hb.step(f, w2, c.dt, c.substeps)
        assert ha.get_particle_count() == hb.get_particle_count()
        sa, sb = ha.stats(), hb.stats()
        assert (sa["lost"], sa["added"], sa["movers"]) == (sb["lost"], sb["added"], sb["movers"]), f"step {s}"
        assert rel_inf(w[0].cpu().numpy(), w2[0].cpu().numpy()) <= REL_TOL and rel_inf(w[1].cpu().numpy(), w2[1].cpu().numpy()) <= REL_TOL
    assert_states_equal(ha.download(), hb.download(), name)


def test_fast_order_quad_scatter_matches_oracle_after_capacity_growth(gpu, oracle):
    """Fast (atomic-order) path with the quad scatter across a capacity growth: the tensor maps of the record buffers
    are re-encoded for the new allocation; the canonicalised state still equals the oracle's."""
    m = cases._tiny(True)
    fx, fy = cases._mix(m, 0.5, 1.0, 0.35, 1.0)
    run_both(gpu, oracle, m, fx, fy, 4, 3, 0.2, 14, check_every=7, capacity_factor=1.02)


@pytest.mark.parametrize("name,chunks", [("channel_fast", 3), ("cyl3_l2", 4), ("tiny_l4", 5)])
def test_pipelined_step_host_equals_the_three_calls(gpu, oracle, name, chunks):
    """pfem2_step_host in its pipelined form (pfem2_options.host_pipeline: chunked move pass behind the sliced upload,
    chunked projection ahead of the sliced download, dependencies derived from the mesh numbering) against the plain
    advect / project / correct calls: same particle set bit for bit, same counters, projected field within the
    tolerance (fast order: the slot order inside a cell, hence the last bits of the nodal sums, is not deterministic).
    channel_fast has a banded numbering (real overlap), cyl3 an unstructured one (dependencies degrade gracefully)."""
    c = cases.build_case(name)
    oracle.complete_mesh(c.mesh)
    dm = gpu.DeviceMesh(c.mesh)
    ha = gpu.ParticleHandler2D(dm, c.level)
    hb = gpu.ParticleHandler2D(dm, c.level, host_pipeline=chunks)
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


@pytest.mark.parametrize("name", ["cyl3_l2", "channel_fast"])
def test_repeated_projection_and_forced_correction_match_oracle(gpu, oracle, name):
    """projectVelocityOntoGrid twice in a row (the per-cell sums are still valid: same bits), then a correction that a reader of
    the particle velocities forces to be applied eagerly (in the permuted state of the lazy re-sort this materialises the sorted
    order first), then a third projection: particle set, counters and projected field must equal the oracle's."""
    c = cases.build_case(name)
    h, o = run_both(gpu, oracle, c.mesh, c.fx, c.fy, c.level, c.substeps, c.dt, 8, check_every=2)
    f, w = dev_field(c)
    w2 = (torch.zeros_like(f[0]), torch.zeros_like(f[0]))
    h.advect_particles(f, c.dt, c.substeps)
    o.advect_particles(c.fx, c.fy, c.dt, c.substeps)
    h.project_velocity_onto_grid(w)
    h.project_velocity_onto_grid(w2)  # partial sums are still valid: same bits
    assert np.array_equal(w[0].cpu().numpy(), w2[0].cpu().numpy())
    h.correct_particle_velocity(f, w)
    h.get_particles()                  # forces the deferred correction: the sums must be recomputed
    h.project_velocity_onto_grid(w2)
    wx, wy = np.zeros_like(c.fx), np.zeros_like(c.fx)
    o.project_velocity_onto_grid(wx, wy)
    o.correct_particle_velocity(c.fx, c.fy, wx, wy)
    o.project_velocity_onto_grid(wx, wy)
    assert rel_inf(w2[0].cpu().numpy(), wx) <= REL_TOL and rel_inf(w2[1].cpu().numpy(), wy) <= REL_TOL


def test_insitu_drop_in_of_the_unmodified_poiseuille_case():
    """The UNMODIFIED cases/PoiseuilleFlow2D/main.cu, linked once against the reference library and once against the library
    whose ParticleHandler2D is the B200 drop-in (gpupfem2_b200/shim, built by __graft_entry__.build() where /root/reference
    exists; the binaries travel to the box under oracle/_ref/): the particle counts printed by advectParticles must be identical
    in every one of the 500 coupled steps and the exported nodal fields must agree to the Krylov tolerance of the FEM stage."""
    import insitu_compare

    if not all(os.path.exists(b) for b in insitu_compare.binaries("poiseuille")):
        pytest.skip("oracle/_ref/Poiseuille_{ref,shim} not built (needs /root/reference at build time)")
    s = insitu_compare.compare("poiseuille")
    assert s["ref"]["rc"] == 0 and s["shim"]["rc"] == 0, s
    assert s["count_steps_compared"] >= 100 and s["count_identical_steps"] == s["count_steps_compared"], s
    assert s["ref"]["created"] == s["shim"]["created"]
    worst = max((v for d in s["field_rel_inf_diff"].values() if isinstance(d, dict) for v in d.values()), default=None)
    assert worst is not None and worst <= 1e-4, s["field_rel_inf_diff"]
