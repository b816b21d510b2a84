"""Strip-partitioned multi-GPU particle step: one process per GPU, torch.distributed for the plumbing.

The reference is single-GPU; this layer is new.  The cell index range [0, C) of ONE global mesh is cut into
contiguous strips (for the x-major channel numbering: slabs of quad columns); rank r owns the cells
[bounds[r], bounds[r+1]) and the particles inside them.  Per time step there are two real exchanges:

  1. particle migration after the move pass: particles whose new cell belongs to another strip are packed as
     64-byte records per destination rank and handed over (counts first, then payload: all_to_all_single);
  2. projection halo sum: the per-node accumulators {sum L v_x, sum L v_y, sum L} of the nodes shared by two
     strips are exchanged pairwise and added BEFORE the division (a + b is commutative in IEEE arithmetic, so
     both sides get the same bits).

Every rank holds the whole (read-only) mesh and nodal field: 16M triangles cost about 3 GB of HBM per GPU, and it
lets the locate step resolve a particle that crosses the interface exactly like the single-GPU rule (own cell,
then ascending one-ring) without halo bookkeeping.

The host logic here (partition, interface node lists, exchange protocol) is backend-agnostic: it moves torch
tensors, CUDA over NCCL in production and CPU over gloo in tests/test_multi_rank_gloo.py.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

RECORD_DOUBLES = 8  # 64-byte particle record = 8 x 8 bytes: {x, y | L0, L1 | L2, (cell, id) | vx, vy}


# ------------------------------------------------------------------------------------------------
# host logic (no CUDA)
# ------------------------------------------------------------------------------------------------
def strip_bounds(n_cells: int, n_ranks: int, align: int = 1) -> np.ndarray:
    """Cell bounds of n_ranks contiguous strips, each a multiple of `align` cells (align = 2*ny keeps whole quad
    columns of the x-major channel together), as equal as possible."""
    if n_cells % align:
        raise ValueError("n_cells must be a multiple of align")
    units = n_cells // align
    if units < n_ranks:
        raise ValueError("fewer strip units than ranks")
    b = [(units * r) // n_ranks * align for r in range(n_ranks + 1)]
    return np.asarray(b, dtype=np.int32)


def owner_of_cells(cells: np.ndarray, bounds: np.ndarray) -> np.ndarray:
    return np.searchsorted(bounds, cells, side="right") - 1


def interface_nodes(cells, bounds, rank: int) -> dict:
    """{neighbour rank: sorted node ids shared by this rank's cells and the neighbour's cells}.

    `cells` is the (C, 3) connectivity as a torch tensor (any device) or numpy array.  Only non-empty
    intersections are returned; for strips these are the adjacent ranks."""
    t = torch.as_tensor(cells) if not isinstance(cells, torch.Tensor) else cells
    t = t.to(torch.int64)
    mine = torch.unique(t[int(bounds[rank]):int(bounds[rank + 1])])
    out = {}
    for r in range(len(bounds) - 1):
        if r == rank or bounds[r] == bounds[r + 1]:
            continue
        if abs(r - rank) > 1 and t.shape[0] > 4_000_000:
            continue  # large strip meshes: only adjacent strips can share nodes
        theirs = torch.unique(t[int(bounds[r]):int(bounds[r + 1])])
        shared = mine[torch.isin(mine, theirs)]
        if shared.numel():
            out[r] = torch.sort(shared).values
    return out


def exchange_records(send_buf: torch.Tensor, send_counts, group=None):
    """send_buf: (sum(send_counts), 8) float64 records grouped by destination rank.  -> (recv_buf, recv_counts)."""
    world = dist.get_world_size(group)
    dev = send_buf.device
    sc = torch.tensor(list(send_counts), dtype=torch.int64, device=dev)
    rc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(v) for v in rc.tolist()]
    recv_buf = torch.empty((sum(recv_counts), RECORD_DOUBLES), dtype=torch.float64, device=dev)
    dist.all_to_all_single(recv_buf, send_buf, output_split_sizes=recv_counts, input_split_sizes=[int(v) for v in send_counts],
                           group=group)
    return recv_buf, recv_counts


def migration_capacity(n_interface_nodes: int, ppc: int, columns: int = 12, floor: int = 16384) -> int:
    """Default record capacity of one migration buffer of the neighbour protocol: the nominal population of `columns`
    cell layers along the interface (a layer has about n_interface_nodes cells), i.e. room for every particle within
    several cells of the interface to leave in one step even after the density has grown."""
    return max(int(floor), int(columns) * int(n_interface_nodes) * int(ppc))


def exchange_neighbours(send_left, send_right, recv_left, recv_right, rank: int, world: int, group=None):
    """Neighbour protocol: hand the fixed-size migration buffers ([header | capacity records], (cap + 1, 8) float64) to the
    adjacent strips.  No sizes are negotiated and nothing is read on the host: the record count travels in the header.
    On return the current stream waits for the receives (CUDA) / the receives are complete (CPU)."""
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, send_left, rank - 1, group=group))
        ops.append(dist.P2POp(dist.irecv, recv_left, rank - 1, group=group))
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, send_right, rank + 1, group=group))
        ops.append(dist.P2POp(dist.irecv, recv_right, rank + 1, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def header_count(buf: torch.Tensor) -> torch.Tensor:
    """Record count in the header of a migration buffer (a 0-d int32 tensor on the buffer's device; no synchronisation)."""
    return buf[0, :1].view(torch.int32)[0]


def exchange_interface(acc3: torch.Tensor, iface: dict, group=None):
    """Add the neighbours' accumulators of the shared nodes into acc3 (N, 3), in place."""
    if not iface:
        return
    ops, recv = [], {}
    send = {r: torch.index_select(acc3, 0, idx) for r, idx in iface.items()}
    for r in sorted(iface):
        recv[r] = torch.empty_like(send[r])
        ops.append(dist.P2POp(dist.isend, send[r], r, group=group))
        ops.append(dist.P2POp(dist.irecv, recv[r], r, group=group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for r in sorted(iface):
        # the node ids of one interface are unique, so this is one plain addition per entry; two contributions per shared
        # node: a + b == b + a bit for bit
        acc3.index_add_(0, iface[r], recv[r])


# ------------------------------------------------------------------------------------------------
# CUDA handler
# ------------------------------------------------------------------------------------------------
class DistributedParticleHandler2D:
    """ParticleHandler2D interface over a strip partition; same method names as handler.ParticleHandler2D."""

    def __init__(self, mesh, cell_division_level, bounds, rank, world, group=None, migration="p2p", migration_cap=0, **opts):
        from . import _lib, handler

        self._lib = _lib
        self.L = _lib.load()
        self.mesh = mesh
        self.rank, self.world, self.group = rank, world, group
        self.bounds = np.ascontiguousarray(bounds, dtype=np.int32)
        self.h = handler.ParticleHandler2D(mesh, cell_division_level, **opts)
        self.h._check(self.L.pfem2_set_owned_cells(self.h._h, int(self.bounds[rank]), int(self.bounds[rank + 1])), "set_owned_cells")
        self.iface = interface_nodes(mesh.cells.view(torch.int32), self.bounds, rank)
        self.acc3 = torch.zeros((mesh.n_nodes, 3), dtype=torch.float64, device=mesh.device)
        self._sent = 0
        self.last_received = 0
        # Migration / halo transport:
        #   "p2p"        NVLink peer memory: the pack kernel stores the emigrants' records straight into the neighbour's inbox,
        #                sequence flags instead of collectives, nothing on the host in the loop (falls back to "neighbour" if the
        #                CUDA IPC set-up fails on any rank);
        #   "neighbour"  fixed-size [header | records] buffers to / from the adjacent strips over ncclSend / ncclRecv, counts stay
        #                on the device (no host round trip inside advect_particles);
        #   "exact"      counts first (host), then exactly sized payloads (all_to_all_single).
        # "p2p" and "neighbour" need the fast order (the move pass lists its emigrants).
        self.protocol = os.environ.get("PFEM2_MG_PROTOCOL", migration)  # env: A/B measurements
        if self.protocol not in ("p2p", "neighbour", "exact"):
            raise ValueError("migration must be 'p2p', 'neighbour' or 'exact'")
        if opts.get("stable_order") or os.environ.get("PFEM2_MG_FUSED") == "0":
            self.protocol = "exact"
        self._nbr = None
        if self.protocol != "exact":
            n_if = max([int(v.numel()) for v in self.iface.values()] or [1])
            self.migration_cap = int(migration_cap) if migration_cap else migration_capacity(n_if, self.h.particles_per_cell)
            self.h._check(self.L.pfem2_set_rank_bounds(self.h._h, self.bounds.ctypes.data_as(C.POINTER(C.c_int)), self.world),
                          "set_rank_bounds")
        if self.protocol == "p2p" and not self._setup_p2p():
            self.protocol = "neighbour"
        if self.protocol == "neighbour":
            shape = (self.migration_cap + 1, RECORD_DOUBLES)
            mk = lambda: torch.zeros(shape, dtype=torch.float64, device=mesh.device)  # noqa: E731
            self._nbr = {"sl": mk() if rank > 0 else None, "rl": mk() if rank > 0 else None,
                         "sr": mk() if rank + 1 < world else None, "rr": mk() if rank + 1 < world else None}

    def _setup_p2p(self) -> bool:
        """Create this strip's inboxes, swap CUDA IPC handles with the neighbours, map theirs.  Collective: every rank takes the
        same decision (all ranks succeed or all fall back)."""
        h, L, rank, world = self.h, self.L, self.rank, self.world
        ok, handles = True, {}
        if any(abs(r - rank) != 1 for r in self.iface):
            ok = False  # not a strip partition: nodes shared with a non-adjacent rank
        if ok:
            for side, nb in ((0, rank - 1), (1, rank + 1)):
                if not 0 <= nb < world:
                    continue
                idx = self.iface[nb].to(torch.int32).cpu().numpy() if nb in self.iface else np.empty(0, dtype=np.int32)
                idx = np.ascontiguousarray(idx)
                buf = C.create_string_buffer(64)
                rc = L.pfem2_p2p_inbox_create(h._h, side, self.migration_cap, int(idx.size), idx.ctypes.data_as(C.POINTER(C.c_int)), buf)
                if rc:
                    ok = False
                    break
                handles[side] = buf.raw
        gathered = [None] * world
        dist.all_gather_object(gathered, (ok, handles), group=self.group)
        ok = all(g[0] for g in gathered)
        if ok:
            if rank > 0:
                ok = ok and L.pfem2_p2p_connect(h._h, 0, gathered[rank - 1][1][1]) == 0
            if rank + 1 < world:
                ok = ok and L.pfem2_p2p_connect(h._h, 1, gathered[rank + 1][1][0]) == 0
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok), group=self.group)
        return all(flags)

    @property
    def last_sent(self):
        """Particles this rank handed over in the last advect (neighbour protocol: read from the device on demand)."""
        if self._sent is None and self._nbr is not None:
            self._sent = sum(int(header_count(b).item()) for b in (self._nbr["sl"], self._nbr["sr"]) if b is not None)
        elif self._sent is None:
            n = C.c_int(0)
            self.h._check(self.L.pfem2_p2p_last_sent(self.h._h, C.byref(n)), "p2p_last_sent")
            self._sent = n.value
        return self._sent

    def seed_particles(self):
        self.h.seed_particles()

    def init_particle_velocity(self, vel):
        self.h.init_particle_velocity(vel)

    def advect_particles(self, vel, time_step, particle_substeps):
        h, L = self.h, self.L
        h._check(L.pfem2_advect_move(h._h, vel[0].data_ptr(), vel[1].data_ptr(), time_step, particle_substeps), "advect_move")
        if self.protocol == "p2p":
            # three library calls, all enqueued on the handle's stream: records go straight into the neighbours' HBM over NVLink
            h._check(L.pfem2_emigrants_send_p2p(h._h, self.rank), "emigrants_send_p2p")
            h._check(L.pfem2_immigrants_recv_p2p(h._h), "immigrants_recv_p2p")
            h._check(L.pfem2_advect_finish(h._h, vel[0].data_ptr(), vel[1].data_ptr()), "advect_finish")
            self._sent = None  # read from the device on demand
            return
        if self._nbr is not None:
            # everything below is enqueued without waiting for the device: the library works on the legacy default stream, which
            # is torch's current stream here, and torch orders the NCCL transfers against it (w.wait() is a stream-side wait)
            b = self._nbr
            ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
            h._check(L.pfem2_emigrants_pack_neighbours(h._h, self.rank, ptr(b["sl"]), ptr(b["sr"]), self.migration_cap), "emigrants_pack_neighbours")
            exchange_neighbours(b["sl"], b["sr"], b["rl"], b["rr"], self.rank, self.world, self.group)
            if b["rl"] is not None:
                h._check(L.pfem2_immigrants_append_device(h._h, b["rl"].data_ptr(), self.migration_cap, 1), "immigrants_append_device")
            if b["rr"] is not None:
                h._check(L.pfem2_immigrants_append_device(h._h, b["rr"].data_ptr(), self.migration_cap, 0), "immigrants_append_device")
            h._check(L.pfem2_advect_finish(h._h, vel[0].data_ptr(), vel[1].data_ptr()), "advect_finish")
            self._sent = None  # read from the send headers on demand
            return
        counts = (C.c_int * self.world)()
        h._check(L.pfem2_emigrants_count(h._h, self.bounds.ctypes.data_as(C.POINTER(C.c_int)), self.world, counts), "emigrants_count")
        send_counts = [int(c) for c in counts]
        send_buf = torch.empty((sum(send_counts), RECORD_DOUBLES), dtype=torch.float64, device=self.mesh.device)
        if send_buf.numel():
            h._check(L.pfem2_emigrants_pack(h._h, send_buf.data_ptr(), send_buf.shape[0]), "emigrants_pack")
        recv_buf, recv_counts = exchange_records(send_buf, send_counts, self.group)
        torch.cuda.current_stream(self.mesh.device).synchronize()  # NCCL ran on torch's stream; the library uses the null stream
        if recv_buf.shape[0]:
            h._check(L.pfem2_immigrants_append(h._h, recv_buf.data_ptr(), recv_buf.shape[0]), "immigrants_append")
        h._check(L.pfem2_advect_finish(h._h, vel[0].data_ptr(), vel[1].data_ptr()), "advect_finish")
        self._keep = (send_buf, recv_buf)
        self._sent, self.last_received = sum(send_counts), sum(recv_counts)

    def project_velocity_onto_grid(self, vel):
        h, L = self.h, self.L
        # no host synchronisation: the library works on the legacy default stream, which is also torch's current stream here,
        # and torch orders the NCCL transfers against it (w.wait() makes the current stream wait for the receive)
        h._check(L.pfem2_project_accumulate(h._h, self.acc3.data_ptr()), "project_accumulate")
        if self.protocol == "p2p":
            h._check(L.pfem2_project_halo_p2p(h._h, self.acc3.data_ptr()), "project_halo_p2p")
        else:
            exchange_interface(self.acc3, self.iface, self.group)
        h._check(L.pfem2_project_finalize(h._h, self.acc3.data_ptr(), vel[0].data_ptr(), vel[1].data_ptr()), "project_finalize")

    def correct_particle_velocity(self, vel, vel_old):
        self.h.correct_particle_velocity(vel, vel_old)

    def step(self, frozen, work, dt, substeps):
        self.advect_particles(frozen, dt, substeps)
        self.project_velocity_onto_grid(work)
        self.correct_particle_velocity(frozen, work)

    def get_particle_count(self):
        return self.h.get_particle_count()

    def global_particle_count(self):
        t = torch.tensor([self.get_particle_count()], dtype=torch.int64, device=self.mesh.device)
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def download(self):
        return self.h.download()

    def close(self):
        if self.protocol == "p2p" and dist.is_initialized():
            torch.cuda.synchronize(self.mesh.device)
            dist.barrier(group=self.group)  # nobody unmaps / frees an inbox a neighbour may still be storing into
        self.h.close()


# ------------------------------------------------------------------------------------------------
# bench.py entry for N > 1 (launched by torchrun, one rank per GPU)
# ------------------------------------------------------------------------------------------------
def bench_main(args, rank, world, local):
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import bench
    from . import handler

    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device(device))
    if args.workload not in bench.WORKLOADS:
        raise SystemExit("multi-GPU bench runs the synthetic channel workloads")
    nx, ny, lx, ly, level, umax, dt = bench.channel_params(args, world)
    dm = handler.device_structured_channel(nx, ny, lx, ly, colmajor=True, device=device)
    fx, fy = bench.nodal_field(args, dm.vertices[:, 0].contiguous(), dm.vertices[:, 1].contiguous(), lx, ly, umax)
    F = (fx.contiguous(), fy.contiguous())
    W = (torch.zeros_like(F[0]), torch.zeros_like(F[0]))
    bounds = strip_bounds(dm.n_cells, world, align=2 * ny)
    h = DistributedParticleHandler2D(dm, level, bounds, rank, world, max_division_level=8, capacity_factor=args.capacity_factor)
    h.seed_particles()
    h.init_particle_velocity(F)
    sampler = bench.ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        h.step(F, W, dt, args.substeps)
    h.get_particle_count()
    torch.cuda.synchronize()
    if sampler:
        sampler.wait_first_sample()
        sampler.mark()
    dist.barrier()
    h.h.set_profiling(True)
    h.h.phase_times(reset=True)
    launches0 = handler.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    counts, sent = [], 0
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(args.steps):
        h.step(F, W, dt, args.substeps)
        counts.append(h.get_particle_count())
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = e0.elapsed_time(e1)
    sent = h.last_sent * args.steps  # the last step's hand-over (read outside the timed region), steady state
    clocks = sampler.stop() if sampler else None
    phases = h.h.phase_times(reset=True)
    launches = handler.kernel_launches() - launches0
    t = torch.tensor([ms, float(sum(counts)), float(sent)], dtype=torch.float64, device=device)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        total_ms = float(tmax[0])
        psteps = float(tsum[1])
        value = psteps / (total_ms * 1e-3)
        peak, peak_src = bench.peaks()
        pmean_rank = sum(counts) / args.steps
        dom = max(bench.ALG_BYTES, key=lambda n: phases[n][0])
        dom_ms = phases[dom][0] / args.steps
        achieved = bench.ALG_BYTES[dom] * pmean_rank / (dom_ms * 1e-3) / 1e9
        out = {
            "metric": "particle-steps/sec (advect+locate+sort+project+correct)", "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak" if args.workload in bench.WEAK else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": bench.workload_description(args, world) + f", strip-partitioned over {world} GPUs (quad columns)",
                       "particles_mean": psteps / args.steps, "cells": dm.n_cells, "nodes": dm.n_nodes, "substeps": args.substeps, "dt": dt,
                       "l2": "inputs larger than L2", "timing": "CUDA events on rank-local default stream, max over ranks, barrier on both sides",
                       "migrated_particles_per_step": float(tsum[2]) / args.steps,
                       "migration_protocol": h.protocol,
                       "collectives": {
                           "p2p": "NVLink peer memory (CUDA IPC): emigrant records and interface-node accumulators stored straight into the "
                                  "neighbour strip's HBM, device-side sequence flags; no NCCL and no host in the loop",
                           "neighbour": "fixed-size migration buffers [header | records] to / from the adjacent strips (ncclSend / ncclRecv, "
                                        "counts stay on the device) + pairwise isend/irecv of interface-node accumulators (NCCL)",
                           "exact": "all_to_all_single (counts, 64-byte particle records) + pairwise isend/irecv of interface-node "
                                    "accumulators (NCCL)"}[h.protocol]},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "alg_bytes_per_particle": bench.ALG_BYTES[dom], "note": "rank 0, per GPU",
                         "step": {"achieved": bench.ALG_BYTES_STEP * value / 1e9 / world, "frac": bench.ALG_BYTES_STEP * value / 1e9 / world / peak,
                                  "alg_bytes_per_particle_step": bench.ALG_BYTES_STEP, "note": "per GPU"},
                         "phases": {k: {"ms_per_step": v[0] / args.steps} for k, v in phases.items()}},
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8 * world,
                    "note": "multi-GPU steps run through the public DistributedParticleHandler2D API; per step each rank reads back its "
                            "particle count (one host sync), nodal fields stay device-resident"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(out))
    h.close()
    dist.barrier()
    dist.destroy_process_group()
