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
    # fast order with the neighbour protocol (default: fixed-size buffers, no host round trip), fast order with the exact-size
    # protocol, deterministic order (exact-size protocol)
    for stable, protocol in ((False, "neighbour"), (False, "exact"), (True, "exact")):
        W = (torch.zeros_like(fx), torch.zeros_like(fx))
        bounds = multi_gpu.strip_bounds(dm.n_cells, world, align=2 * ny)
        h = multi_gpu.DistributedParticleHandler2D(dm, level, bounds, rank, world, migration=protocol, stable_order=stable)
        assert h.protocol == protocol, (h.protocol, protocol)
        h.seed_particles()
        h.init_particle_velocity(F)
        moved = 0
        for _ in range(nsteps):
            h.step(F, W, dt, S)
            moved += h.last_sent
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
            for _ in range(nsteps):
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
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
