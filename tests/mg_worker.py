"""Multi-GPU parity worker, launched by torchrun (one rank per GPU) from tests/test_gpu_multi.py:
the strip-partitioned CUDA path on N GPUs must equal the single-GPU CUDA path on the same global problem
(owner cells / positions / local coordinates bit-exact, velocities and nodal field within 1e-12)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def spill_case(handler, multi_gpu, rank, world, dev, protocol):
    """SURVEY N4 across a strip boundary.  A particle in the tolerance band of the LAST cell of rank 0 (second barycentric
    in [-2e-6, 0)) gets a flat sub-cell index that lands in the occupancy word of the next cell, i.e. the FIRST cell of
    rank 1; there it suppresses the re-seeding of that sub-cell.  The neighbour protocol carries those bits in the
    migration header, so the strip-partitioned run re-seeds exactly like a single GPU.  Returns (single, multi) counts."""
    nx, ny, level = 8, 4, 2
    ppc = level * level
    dm = handler.device_structured_channel(nx, ny, 2.0, 1.0, colmajor=True, device=dev)
    zero = torch.zeros(dm.n_nodes, dtype=torch.float64, device=dev)
    F, W = (zero, zero.clone()), (zero.clone(), zero.clone())
    bounds = multi_gpu.strip_bounds(dm.n_cells, world, align=2 * ny)
    ref = handler.ParticleHandler2D(dm, level)
    ref.seed_particles()
    ref.init_particle_velocity(F)
    s = ref.download()  # seeded order: particle of (cell, sub-cell) at cell * ppc + sub-cell
    c = int(bounds[1]) - 1  # last cell of rank 0
    tri = dm.cells[c].cpu().numpy().view(np.uint32)
    v = dm.vertices.cpu().numpy()[tri.astype(np.int64)]
    L = np.array([0.3, -1.0e-6, 0.7 + 1.0e-6])  # inside the +-2e-6 band, just across the edge opposite vertex 1
    pos = L[0] * v[0] + L[1] * v[1] + L[2] * v[2]
    keep = np.ones(s["x"].shape[0], dtype=bool)
    keep[(c + 1) * ppc + 0] = keep[(c + 1) * ppc + 1] = False  # empty the two sub-cells of cell c + 1 the spill can reach (4 + 2j, +1)
    st = {k: a[keep] for k, a in s.items()}
    add = {"x": pos[0], "y": pos[1], "l0": L[0], "l1": L[1], "l2": L[2], "vx": 0.0, "vy": 0.0, "cell": c, "id": 0}
    st = {k: np.concatenate([a, np.asarray([add[k]], dtype=a.dtype)]) for k, a in st.items()}
    ref.upload(st)
    ref.step(F, W, 0.01, 3)
    single = (ref.get_particle_count(), ref.stats()["added"])
    ref.close()
    h = multi_gpu.DistributedParticleHandler2D(dm, level, bounds, rank, world, migration=protocol)
    h.seed_particles()
    h.init_particle_velocity(F)
    own = (st["cell"] >= int(bounds[rank])) & (st["cell"] < int(bounds[rank + 1]))
    h.h.upload({k: a[own] for k, a in st.items()})
    h.step(F, W, 0.01, 3)
    multi = h.global_particle_count()
    h.close()
    return single, multi, h.protocol


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    from gpupfem2_b200 import handler, multi_gpu
    from helpers import REL_TOL, assert_states_equal, rel_inf

    nx, ny, level, S, nsteps = 48, 16, 4, 3, 10
    dm = handler.device_structured_channel(nx, ny, 6.0, 2.0, colmajor=True, device=dev)
    y = dm.vertices[:, 1].contiguous()
    x = dm.vertices[:, 0].contiguous()
    k = 2.0 * np.pi / 1.0
    fx = (4.0 * 1.0 * y * (2.0 - y) / 4.0 + 0.3 * torch.sin(k * x) * torch.cos(k * y)).contiguous()
    fy = (-0.3 * torch.cos(k * x) * torch.sin(k * y)).contiguous()
    dt = 0.25 * (6.0 / nx) * S
    F = (fx, fy)
    # fast order over NVLink peer memory (P2P inboxes, no NCCL in the loop), fast order with the neighbour protocol over NCCL
    # (fixed-size buffers, no host round trip), fast order with the exact-size protocol, deterministic order (exact-size protocol)
    for stable, protocol in ((False, "p2p"), (False, "neighbour"), (False, "exact"), (True, "exact")):
        W = (torch.zeros_like(fx), torch.zeros_like(fx))
        bounds = multi_gpu.strip_bounds(dm.n_cells, world, align=2 * ny)
        h = multi_gpu.DistributedParticleHandler2D(dm, level, bounds, rank, world, migration=protocol, stable_order=stable)
        # "p2p" falls back to "neighbour" (on every rank alike) where CUDA IPC between the GPUs is not available
        assert h.protocol == protocol or (protocol == "p2p" and h.protocol == "neighbour"), (h.protocol, protocol)
        protocol = h.protocol if protocol != "p2p" else f"p2p->{h.protocol}"
        h.seed_particles()
        h.init_particle_velocity(F)
        moved = 0
        for _ in range(nsteps):
            h.step(F, W, dt, S)
            moved += h.last_sent
        host_steps = 0
        if h.protocol == "p2p" and not stable:
            # the C-side driver of the strip step with HOST nodal buffers (pfem2_step_host_p2p), plain and pipelined in 3 chunks
            # (a second handle: host_pipeline is a create-time option); results land in pinned host arrays
            hF = [t.cpu().pin_memory() for t in F]
            hW = [torch.zeros_like(t).pin_memory() for t in hF]
            host_steps = 3
            for _ in range(host_steps):
                h.step_host(hF, hW, F, W, dt, S)
                moved += h.last_sent
            W = (hW[0].to(dev), hW[1].to(dev))
            h3 = multi_gpu.DistributedParticleHandler2D(dm, level, bounds, rank, world, migration="p2p", host_pipeline=3)
            h3.seed_particles()
            h3.init_particle_velocity(F)
            hW3 = [torch.zeros_like(t).pin_memory() for t in hF]
            for _ in range(nsteps + host_steps):
                h3.step_host(hF, hW3, F, None, dt, S)
            assert torch.equal(h3.state_checksum(), h.state_checksum()), "pipelined pfem2_step_host_p2p: state differs from the plain form"
            ilo, ihi, olo, ohi = h3.h.node_ranges(S)
            for a, b in zip(hW3, hW):
                e = (a[olo:ohi] - b[olo:ohi]).abs().max() / b.abs().max()
                assert float(e) <= REL_TOL, f"pipelined pfem2_step_host_p2p: projected field differs by {float(e):.3e}"
            h3.close()
        total = h.global_particle_count()
        state = h.download()
        mine = torch.zeros(dm.n_nodes, dtype=torch.bool, device=dev)
        mine[torch.unique(dm.cells[int(bounds[rank]):int(bounds[rank + 1])].to(torch.int64))] = True
        gathered = [None] * world
        dist.all_gather_object(gathered, (state, W[0].cpu().numpy(), W[1].cpu().numpy(), mine.cpu().numpy(), moved))
        if rank == 0:
            ref = handler.ParticleHandler2D(dm, level, stable_order=stable)
            ref.seed_particles()
            ref.init_particle_velocity(F)
            RW = (torch.zeros_like(fx), torch.zeros_like(fx))
            for _ in range(nsteps + host_steps):
                ref.step(F, RW, dt, S)
            merged = gathered[0][0]
            for g in gathered[1:]:
                merged = {kk: np.concatenate([merged[kk], g[0][kk]]) for kk in merged}
            assert merged["x"].shape[0] == total == ref.get_particle_count(), (merged["x"].shape[0], total, ref.get_particle_count())
            assert_states_equal(merged, ref.download(), f"{world} GPUs vs 1 GPU (stable={stable}, {protocol})")
            assert sum(g[4] for g in gathered) > 0
            rwx, rwy = RW[0].cpu().numpy(), RW[1].cpu().numpy()
            for (_, wx, wy, m, _) in gathered:
                assert rel_inf(wx[m], rwx[m]) <= REL_TOL and rel_inf(wy[m], rwy[m]) <= REL_TOL
            print(f"MG_OK world={world} stable={stable} protocol={protocol} particles={total} migrated={sum(g[4] for g in gathered)}")
            ref.close()
        h.close()
        dist.barrier()
    # tolerance-band spill of the occupancy bits across the strip boundary: carried by the migration header
    for want in ("p2p", "neighbour"):
        single, multi, protocol = spill_case(handler, multi_gpu, rank, world, dev, want)
        if rank == 0:
            assert single[1] == 1, f"the crafted state does not exercise the spill: single GPU re-seeded {single[1]} sub-cells, expected 1"
            assert multi == single[0], f"{world} GPUs hold {multi} particles, one GPU {single[0]}: spill bits lost at the strip boundary"
            print(f"MG_SPILL_OK world={world} protocol={want}->{protocol} single_gpu(count, added)={single} multi_gpu_count={multi}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
