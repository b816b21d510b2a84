"""world_size-2 and -3 CPU tests (gloo) of the multi-GPU host logic in gpupfem2_b200/multi_gpu.py: strip partition,
interface node lists, particle migration (all_to_all of 64-byte records) and the projection halo sum.  The compute on
each rank is the CPU oracle; the result must equal the single-rank oracle on the same global problem:
owner cells / positions / local coordinates bit-exact, velocities and nodal field within 1e-12."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from gpupfem2_b200 import multi_gpu  # noqa: E402
from gpupfem2_b200.mesh import structured_channel  # noqa: E402

F64 = ("x", "y", "l0", "l1", "l2", "vx", "vy")


def to_records(state, idx):
    rec = np.empty((idx.size, 8), dtype=np.float64)
    rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3], rec[:, 4] = (state[k][idx] for k in ("x", "y", "l0", "l1", "l2"))
    tag = np.empty((idx.size, 2), dtype=np.uint32)
    tag[:, 0], tag[:, 1] = state["cell"][idx], state["id"][idx]
    rec[:, 5] = tag.view(np.float64)[:, 0]
    rec[:, 6], rec[:, 7] = state["vx"][idx], state["vy"][idx]
    return rec


def from_records(rec):
    tag = np.ascontiguousarray(rec[:, 5]).view(np.uint32).reshape(-1, 2)
    return {"x": rec[:, 0].copy(), "y": rec[:, 1].copy(), "l0": rec[:, 2].copy(), "l1": rec[:, 3].copy(), "l2": rec[:, 4].copy(),
            "vx": rec[:, 6].copy(), "vy": rec[:, 7].copy(), "cell": tag[:, 0].copy(), "id": tag[:, 1].copy()}


def concat(a, b):
    return {k: np.concatenate([a[k], b[k]]) for k in a}


def problem():
    import cases

    m = structured_channel(12, 6, 2.0, 1.0, colmajor=True)
    fx, fy = cases._mix(m, 0.5, 1.0, 0.3, 1.0)  # flow in +x: particles cross the strip interface, leave at x = lx
    return m, fx, fy, 3, 3, 0.2, 12


def pack_neighbour_buffer(buf, rec):
    """[64-byte header | cap records]: count in the first 32-bit word of the header (the layout of csrc MigrationHeader)."""
    buf.zero_()
    buf[0, :1].view(torch.int32)[0] = rec.shape[0]
    assert rec.shape[0] <= buf.shape[0] - 1, "migration buffer too small for the test"
    if rec.shape[0]:
        buf[1:1 + rec.shape[0]] = torch.from_numpy(rec)


def worker(rank, world, port, out, protocol="exact"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc

    m, fx, fy, level, S, dt, nsteps = problem()
    orc.complete_mesh(m)
    bounds = multi_gpu.strip_bounds(m.n_cells, world, align=2 * 6)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    iface = multi_gpu.interface_nodes(m.cells.astype(np.int64), bounds, rank)
    o = orc.OracleHandler(m, level)
    o.seed_particles()
    s = o.download()
    keep = (s["cell"] >= lo) & (s["cell"] < hi)
    o.upload({k: v[keep] for k, v in s.items()})
    o.init_particle_velocity(fx, fy)
    wx, wy = np.zeros_like(fx), np.zeros_like(fx)
    migrated = 0
    for _ in range(nsteps):
        o.move(fx, fy, dt, S)
        s = o.download()
        owner = multi_gpu.owner_of_cells(s["cell"].astype(np.int64), bounds)
        stay = np.nonzero(owner == rank)[0]
        send_counts, parts = [], []
        for r in range(world):
            idx = np.nonzero(owner == r)[0] if r != rank else np.empty(0, dtype=np.int64)
            send_counts.append(idx.size)
            parts.append(to_records(s, idx))
        migrated += sum(send_counts)
        if protocol == "exact":
            send = torch.from_numpy(np.concatenate(parts) if parts else np.empty((0, 8)))
            recv, recv_counts = multi_gpu.exchange_records(send, send_counts)
            arrived = recv.numpy()
        else:  # neighbour protocol: fixed-size buffers to / from the adjacent strips, the count travels in the header
            assert all(n == 0 for r, n in enumerate(send_counts) if abs(r - rank) != 1), "emigrant bound for a non-adjacent strip"
            cap = multi_gpu.migration_capacity(max(int(v.numel()) for v in iface.values()), level * level, floor=64)
            mk = lambda: torch.zeros((cap + 1, multi_gpu.RECORD_DOUBLES), dtype=torch.float64)  # noqa: E731
            sl, rl = (mk(), mk()) if rank > 0 else (None, None)
            sr, rr = (mk(), mk()) if rank + 1 < world else (None, None)
            if sl is not None:
                pack_neighbour_buffer(sl, parts[rank - 1])
            if sr is not None:
                pack_neighbour_buffer(sr, parts[rank + 1])
            multi_gpu.exchange_neighbours(sl, sr, rl, rr, rank, world)
            got = [b[1:1 + int(multi_gpu.header_count(b))].numpy() for b in (rl, rr) if b is not None]
            arrived = np.concatenate(got) if got else np.empty((0, 8))
        local = {k: v[stay] for k, v in s.items()}
        o.upload(concat(local, from_records(arrived)))
        o.check_distribution(fx, fy, lo, hi)
        acc = torch.from_numpy(o.project_accumulate())
        multi_gpu.exchange_interface(acc, iface)
        a = acc.numpy()
        with np.errstate(invalid="ignore", divide="ignore"):
            wx, wy = a[:, 0] / a[:, 2], a[:, 1] / a[:, 2]
        mine = np.zeros(m.n_nodes, dtype=bool)
        mine[np.unique(m.cells[lo:hi])] = True
        gx, gy = np.where(mine, wx, 0.0), np.where(mine, wy, 0.0)  # nodes this rank owns or shares are valid
        o.correct_particle_velocity(fx, fy, np.ascontiguousarray(gx), np.ascontiguousarray(gy))
    gathered = [None] * world
    dist.all_gather_object(gathered, (o.download(), wx, wy, mine, migrated))
    if rank == 0:
        torch.save(gathered, out)
    dist.barrier()
    dist.destroy_process_group()


def test_strip_bounds_and_owner():
    b = multi_gpu.strip_bounds(144, 2, align=12)
    assert list(b) == [0, 72, 144]
    b = multi_gpu.strip_bounds(16_000_000, 8, align=4000)
    assert b[0] == 0 and b[-1] == 16_000_000 and np.all(np.diff(b) % 4000 == 0) and np.all(np.diff(b) == 2_000_000)
    assert list(multi_gpu.owner_of_cells(np.array([0, 71, 72, 143]), np.array([0, 72, 144]))) == [0, 0, 1, 1]
    with pytest.raises(ValueError):
        multi_gpu.strip_bounds(24, 3, align=12)


def test_interface_nodes_are_the_shared_column():
    m = structured_channel(12, 6, 2.0, 1.0, colmajor=True)
    bounds = multi_gpu.strip_bounds(m.n_cells, 3, align=12)  # strips of 4 quad columns
    for rank in range(3):
        it = multi_gpu.interface_nodes(m.cells.astype(np.int64), bounds, rank)
        assert sorted(it) == [r for r in (rank - 1, rank + 1) if 0 <= r < 3]
        for r, nodes in it.items():
            col = max(rank, r) * 4  # shared node column i = 4 or 8; node id = i*(ny+1)+j
            assert nodes.tolist() == [col * 7 + j for j in range(7)]


def test_migration_capacity_default():
    assert multi_gpu.migration_capacity(2001, 16) == 12 * 2001 * 16  # channel16m: 24.6 MB per direction
    assert multi_gpu.migration_capacity(7, 4) == 16384  # floor for small meshes


@pytest.mark.parametrize("world,protocol", [(2, "exact"), (2, "neighbour"), (3, "neighbour")])
def test_multi_rank_step_equals_single_rank(tmp_path, oracle, world, protocol):
    import socket

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "gathered.pt")
    mp.spawn(worker, args=(world, port, out, protocol), nprocs=world, join=True)
    gathered = torch.load(out, weights_only=False)

    from helpers import REL_TOL, assert_states_equal, rel_inf

    m, fx, fy, level, S, dt, nsteps = problem()
    ref = oracle.OracleHandler(m, level)
    ref.seed_particles()
    ref.init_particle_velocity(fx, fy)
    rwx, rwy = np.zeros_like(fx), np.zeros_like(fx)
    for _ in range(nsteps):
        ref.step(fx, fy, rwx, rwy, dt, S)
    merged = gathered[0][0]
    for g in gathered[1:]:
        merged = concat(merged, g[0])
    assert_states_equal(merged, ref.download(), f"{world} ranks vs 1 rank ({protocol})")
    assert sum(g[4] for g in gathered) > 0, "no particle crossed the interface: the test would prove nothing"
    for (_, wx, wy, mine, _) in gathered:
        assert rel_inf(wx[mine], rwx[mine]) <= REL_TOL and rel_inf(wy[mine], rwy[mine]) <= REL_TOL


def test_interface_lists_are_symmetric_and_ordered():
    """Both strips of an interface must hold the SAME ascending node list: the P2P halo (k_halo_send / k_halo_add) and the NCCL
    halo exchange pair entries by position.  Structured strips, and an unstructured mesh cut by cell index (where a strip
    can touch non-adjacent strips: the P2P transport then declines and the NCCL path is used)."""
    cyl = np.load(os.path.join(ROOT, "tests", "golden", "mesh_cylinder3.npz"))
    cases_ = [(structured_channel(24, 8, 3.0, 1.0, colmajor=True).cells.astype(np.int64), 16, 4),
              (cyl["cells"].astype(np.int64), 1, 3)]
    for cells, align, world in cases_:
        n = (cells.shape[0] // align) * align
        bounds = multi_gpu.strip_bounds(n, world, align=align)
        lists = [multi_gpu.interface_nodes(cells[:n], bounds, r) for r in range(world)]
        for r in range(world):
            for s, nodes in lists[r].items():
                assert r in lists[s], (r, s)
                assert torch.equal(nodes, lists[s][r])
                assert torch.all(nodes[1:] > nodes[:-1])  # strictly ascending, no duplicates
    # structured strips only ever touch their neighbours
    cells = cases_[0][0]
    bounds = multi_gpu.strip_bounds(cells.shape[0], 4, align=16)
    for r in range(4):
        assert all(abs(s - r) == 1 for s in multi_gpu.interface_nodes(cells, bounds, r))
