"""Per-call breakdown of the multi-GPU step (rank 0 prints).  torchrun --nproc-per-node N tools/diag_mg.py"""
import os, sys, time, argparse, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from gpupfem2_b200 import handler, multi_gpu as mg

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
args = argparse.Namespace(workload="channel16m", level=0, substeps=3, cfl=0.25, capacity_factor=1.3)
nx, ny, lx, ly, level, umax, dt = bench.channel_params(args, world)
dm = handler.device_structured_channel(nx, ny, lx, ly, colmajor=True, device=dev)
fx, fy = bench.nodal_field(args, dm.vertices[:, 0].contiguous(), dm.vertices[:, 1].contiguous(), lx, ly, umax)
F = (fx.contiguous(), fy.contiguous()); W = (torch.zeros_like(F[0]), torch.zeros_like(F[0]))
h = mg.DistributedParticleHandler2D(dm, level, mg.strip_bounds(dm.n_cells, world, align=2 * ny), rank, world, max_division_level=8, capacity_factor=1.3)
h.seed_particles(); h.init_particle_velocity(F)
for _ in range(3): h.step(F, W, dt, 3)
torch.cuda.synchronize(); dist.barrier()
acc = {}
def timed(name, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t) * 1e3
    return r
L, hh = h.L, h.h
n = 10
for _ in range(n):
    timed("move", lambda: hh._check(L.pfem2_advect_move(hh._h, F[0].data_ptr(), F[1].data_ptr(), dt, 3), "m"))
    if h.protocol == "p2p":
        timed("emig_send", lambda: hh._check(L.pfem2_emigrants_send_p2p(hh._h, rank), "s"))
        timed("immig_recv", lambda: hh._check(L.pfem2_immigrants_recv_p2p(hh._h), "r"))
    elif h.protocol == "neighbour":
        b = h._nbr
        ptr = lambda t: t.data_ptr() if t is not None else None
        timed("emig_pack", lambda: hh._check(L.pfem2_emigrants_pack_neighbours(hh._h, rank, ptr(b["sl"]), ptr(b["sr"]), h.migration_cap), "p"))
        timed("exchange", lambda: mg.exchange_neighbours(b["sl"], b["sr"], b["rl"], b["rr"], rank, world, None))
        for key, left in (("rl", 1), ("rr", 0)):
            if b[key] is not None:
                timed("append", lambda: hh._check(L.pfem2_immigrants_append_device(hh._h, b[key].data_ptr(), h.migration_cap, left), "a"))
    else:
        counts = (C.c_int * world)()
        timed("emig_count", lambda: hh._check(L.pfem2_emigrants_count(hh._h, h.bounds.ctypes.data_as(C.POINTER(C.c_int)), world, counts), "c"))
        sc = [int(c) for c in counts]
        sb = torch.empty((sum(sc), 8), dtype=torch.float64, device=dev)
        if sb.numel(): timed("emig_pack", lambda: hh._check(L.pfem2_emigrants_pack(hh._h, sb.data_ptr(), sb.shape[0]), "p"))
        rb, rc = timed("exchange", lambda: mg.exchange_records(sb, sc, None))
        if rb.shape[0]: timed("append", lambda: hh._check(L.pfem2_immigrants_append(hh._h, rb.data_ptr(), rb.shape[0]), "a"))
    timed("finish", lambda: hh._check(L.pfem2_advect_finish(hh._h, F[0].data_ptr(), F[1].data_ptr()), "f"))
    timed("proj_acc", lambda: hh._check(L.pfem2_project_accumulate(hh._h, h.acc3.data_ptr()), "pa"))
    if h.protocol == "p2p":
        timed("halo", lambda: hh._check(L.pfem2_project_halo_p2p(hh._h, h.acc3.data_ptr()), "h"))
    else:
        timed("halo", lambda: mg.exchange_interface(h.acc3, h.iface, None))
    timed("proj_fin", lambda: hh._check(L.pfem2_project_finalize(hh._h, h.acc3.data_ptr(), W[0].data_ptr(), W[1].data_ptr()), "pf"))
    timed("correct", lambda: h.correct_particle_velocity(F, W))
    timed("count", lambda: h.get_particle_count())
if rank == 0:
    print("world", world, h.protocol, "cap", getattr(h, "migration_cap", None), {k: round(v / n, 3) for k, v in acc.items()}, "sum", round(sum(acc.values()) / n, 3))
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(n):
    h.step(F, W, dt, 3); h.get_particle_count()
torch.cuda.synchronize()
if rank == 0: print("untimed loop ms/step", round((time.perf_counter() - t) / n * 1e3, 3))
h.close(); dist.barrier(); dist.destroy_process_group()
