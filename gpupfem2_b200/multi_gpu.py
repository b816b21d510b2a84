"""Strip-partitioned multi-GPU particle step: one process per GPU, torch.distributed for the plumbing.

The reference is single-GPU; this layer is new.  The cell index range [0, C) of ONE global mesh is cut into
contiguous strips (for the x-major channel numbering: slabs of quad columns); rank r owns the cells
[bounds[r], bounds[r+1]) and the particles inside them.  Per time step there are two real exchanges:

  1. particle migration after the move pass: particles whose new cell belongs to another strip are packed as
     64-byte records per destination rank and handed over (counts first, then payload: all_to_all_single);
  2. projection halo sum: the per-node accumulators {sum L v_x, sum L v_y, sum L} of the nodes shared by two
     strips are exchanged pairwise and added BEFORE the division (a + b is commutative in IEEE arithmetic, so
     both sides get the same bits).

The mesh is PARTITIONED (round 2): a rank holds its own cells plus a halo of the cells a particle can reach in one advect
(band width of the one-ring lists x substeps on either side of the strip, `halo_cells`), as a mesh slice with its own
numbering (handler.device_structured_channel(col_lo=, col_hi=) / handler.device_mesh_slice; `cell_base`, `node_base` are the
offsets into the global numbering).  Inside the halo the locate step resolves a crossing particle exactly like the single-GPU
rule (own cell, then ascending one-ring: the slice's one-ring lists are the global ones for every cell a particle of the strip
can visit); records that cross strips carry global cell ids.  Passing the whole global mesh to every rank (cell_base = 0) still
works and is what the replicated-mesh tests do.

The host logic here (partition, interface node lists, exchange protocol) is backend-agnostic: it moves torch
tensors, CUDA over NCCL in production and CPU over gloo in tests/test_multi_rank_gloo.py.
"""
from __future__ import annotations

import ctypes as C
import os

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
    intersections are returned; for strips these are the adjacent ranks.  Two strips whose node id ranges are disjoint
    cannot share a node, so the (expensive) set intersection is only made where the ranges overlap: nothing is assumed
    about the numbering, a partition that is not a strip partition yields its non-adjacent pairs as well."""
    t = torch.as_tensor(cells) if not isinstance(cells, torch.Tensor) else cells
    t = t.to(torch.int64)
    n_ranks = len(bounds) - 1
    span = {}
    for r in range(n_ranks):
        if bounds[r] < bounds[r + 1]:
            lo, hi = torch.aminmax(t[int(bounds[r]):int(bounds[r + 1])])
            span[r] = (int(lo), int(hi))
    if rank not in span:
        return {}
    mine = None
    out = {}
    for r in range(n_ranks):
        if r == rank or r not in span:
            continue
        if span[r][0] > span[rank][1] or span[r][1] < span[rank][0]:
            continue  # disjoint node id ranges: no shared node
        if mine is None:
            mine = torch.unique(t[int(bounds[rank]):int(bounds[rank + 1])])
        theirs = torch.unique(t[int(bounds[r]):int(bounds[r + 1])])
        shared = mine[torch.isin(mine, theirs)]
        if shared.numel():
            out[r] = torch.sort(shared).values
    return out


def halo_cells(band: int, substeps: int) -> int:
    """Cells a strip's mesh slice needs on either side of its own range: a particle's cell index changes by at most `band`
    (pfem2_mesh_band) per substep, and the one-ring scan of the last substep looks one more ring ahead."""
    return int(band) * (int(substeps) + 1)


def channel_slice_columns(nx: int, ny: int, bounds, rank: int, band: int, substeps: int):
    """Quad columns [col_lo, col_hi) of the x-major channel that rank `rank` needs: its own columns + the halo."""
    per_col = 2 * ny
    halo = -(-halo_cells(band, substeps) // per_col)
    return max(0, int(bounds[rank]) // per_col - halo), min(nx, -(-int(bounds[rank + 1]) // per_col) + halo)


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
        if mesh.device.type == "cuda" and torch.cuda.current_stream(mesh.device).cuda_stream != 0:
            # the library works on the legacy default stream; the NCCL transports order their transfers against torch's CURRENT stream
            raise RuntimeError("DistributedParticleHandler2D must be created and driven on torch's default stream")
        self.mesh = mesh
        self.rank, self.world, self.group = rank, world, group
        self.bounds = np.ascontiguousarray(bounds, dtype=np.int32)  # GLOBAL cell ids
        self.cell_base = int(getattr(mesh, "cell_base", 0))         # the mesh may be a slice of the global one (partitioned run)
        lo, hi = int(self.bounds[rank]) - self.cell_base, int(self.bounds[rank + 1]) - self.cell_base
        if not 0 <= lo <= hi <= mesh.n_cells:
            raise ValueError("the mesh slice does not contain the strip's own cells")
        self.h = handler.ParticleHandler2D(mesh, cell_division_level, **opts)
        if self.cell_base:
            self.h._check(self.L.pfem2_set_global_cell_offset(self.h._h, self.cell_base), "set_global_cell_offset")
        self.h._check(self.L.pfem2_set_owned_cells(self.h._h, lo, hi), "set_owned_cells")
        # interface nodes in the slice's numbering; both strips see the cells on either side of their common boundary (halo >= one
        # ring), so they find the same nodes in the same (ascending) order
        local_bounds = np.clip(self.bounds.astype(np.int64) - self.cell_base, 0, mesh.n_cells)
        self.iface = interface_nodes(mesh.cells.view(torch.int32), local_bounds, rank)
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
        if len(self.iface) == 2 and bool(torch.isin(self.iface[rank - 1], self.iface[rank + 1]).any()):
            ok = False  # a strip so thin that one node touches both neighbours: the fused halo kernel adds each side independently
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
        if self.protocol == "p2p":
            # one library call, everything enqueued on the handle's stream: boundary layers moved first, emigrants stored straight into
            # the neighbours' HBM over NVLink, interior moved while the delivery travels, immigrants appended, rank pass
            h._check(L.pfem2_advect_p2p(h._h, self.rank, vel[0].data_ptr(), vel[1].data_ptr(), time_step, particle_substeps), "advect_p2p")
            self._sent = None  # read from the device on demand
            return
        h._check(L.pfem2_advect_move(h._h, vel[0].data_ptr(), vel[1].data_ptr(), time_step, particle_substeps), "advect_move")
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
        if self.protocol == "p2p":
            h._check(L.pfem2_project_p2p(h._h, self.acc3.data_ptr(), vel[0].data_ptr(), vel[1].data_ptr()), "project_p2p")
            return
        # no host synchronisation: the library works on the legacy default stream, which is also torch's current stream here,
        # and torch orders the NCCL transfers against it (w.wait() makes the current stream wait for the receive)
        h._check(L.pfem2_project_accumulate(h._h, self.acc3.data_ptr()), "project_accumulate")
        exchange_interface(self.acc3, self.iface, self.group)
        h._check(L.pfem2_project_finalize(h._h, self.acc3.data_ptr(), vel[0].data_ptr(), vel[1].data_ptr()), "project_finalize")

    def correct_particle_velocity(self, vel, vel_old):
        self.h.correct_particle_velocity(vel, vel_old)

    def step(self, frozen, work, dt, substeps):
        self.advect_particles(frozen, dt, substeps)
        self.project_velocity_onto_grid(work)
        self.correct_particle_velocity(frozen, work)

    def step_host(self, h_frozen, h_work, d_frozen, d_work, dt, substeps) -> int:
        """The step with HOST nodal buffers (pinned torch tensors over this rank's node numbering): uploads the slice of the nodal
        field this strip's advect can read, runs the step, downloads the slice of the projected field this strip owns (interface
        nodes: identical bits on both strips after the halo sum) and returns the strip's particle count.
        P2P transport: ONE C call (pfem2_step_host_p2p) that pipelines the copies with the move pass and the projection and
        drives the whole multi-GPU step from C; NCCL transports: the slices are staged through d_frozen / d_work around step()."""
        if self.protocol == "p2p":
            n = C.c_int(0)
            self.h._check(self.L.pfem2_step_host_p2p(self.h._h, self.rank, h_frozen[0].data_ptr(), h_frozen[1].data_ptr(),
                                                     h_work[0].data_ptr(), h_work[1].data_ptr(), dt, substeps, C.byref(n)), "step_host_p2p")
            self._sent = None
            return n.value
        ilo, ihi, olo, ohi = self.h.node_ranges(substeps)
        for k in range(2):
            d_frozen[k][ilo:ihi].copy_(h_frozen[k][ilo:ihi], non_blocking=True)
        self.step(d_frozen, d_work, dt, substeps)
        for k in range(2):
            h_work[k][olo:ohi].copy_(d_work[k][olo:ohi], non_blocking=True)
        n = self.get_particle_count()
        torch.cuda.current_stream(self.mesh.device).synchronize()
        return n

    def host_bytes_per_step(self, substeps):
        ilo, ihi, olo, ohi = self.h.node_ranges(substeps)
        return 2 * 8 * (ihi - ilo), 2 * 8 * (ohi - olo) + 4

    def get_particle_count(self):
        return self.h.get_particle_count()

    def state_checksum(self) -> torch.Tensor:
        """handler.ParticleHandler2D.state_checksum summed over the strips: the checksum of the GLOBAL particle set, equal to a
        single GPU's on the same problem iff owner cells and positions agree bit for bit."""
        t = self.h.state_checksum()
        t[6] += t[0] * self.cell_base  # cell ids of a mesh slice -> global
        if dist.is_initialized():
            dist.all_reduce(t, group=self.group)  # wrapping int64 sums
        return t

    def global_particle_count(self):
        t = torch.tensor([self.get_particle_count()], dtype=torch.int64, device=self.mesh.device)
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def download(self):
        """This strip's particles; cell ids GLOBAL."""
        s = self.h.download()
        if self.cell_base:
            s["cell"] = (s["cell"].astype(np.int64) + self.cell_base).astype(np.uint32)
        return s

    def close(self):
        if self.protocol == "p2p" and dist.is_initialized():
            torch.cuda.synchronize(self.mesh.device)
            dist.barrier(group=self.group)  # nobody unmaps / frees an inbox a neighbour may still be storing into
        self.h.close()


# ------------------------------------------------------------------------------------------------
# run-time parity check of the strip-partitioned path (bench.py runs it inside the same torchrun before it times anything)
# ------------------------------------------------------------------------------------------------
def parity_selfcheck(rank, world, device, steps=6, migration="p2p") -> dict:
    """A small channel over `world` strips against ONE GPU on the same global problem (rank 0 runs both): global particle count
    and the order-independent state checksum (owner cells, positions, local coordinates: bit-exact) after every step, projected
    nodal field of every strip's nodes within 1e-12, and the tolerance-band spill of the occupancy bits across a strip
    boundary (crafted state, the strips must re-seed exactly like one GPU).  Raises on a mismatch.
    (tests/mg_worker.py is the full version: all transports, the stable order, canonicalised state comparison.)"""
    from . import handler

    nx, ny, level, S = 16 * world, 16, 4, 3
    lx, ly = 0.5 * world, 0.5
    k = 2.0 * np.pi / 0.25

    def field(m):
        x, y = m.vertices[:, 0].contiguous(), m.vertices[:, 1].contiguous()
        return ((4.0 * y * (0.5 - y) / 0.25 + 0.3 * torch.sin(k * x) * torch.cos(k * y)).contiguous(),
                (-0.3 * torch.cos(k * x) * torch.sin(k * y)).contiguous())

    dt = 0.3 * (lx / nx) * S
    bounds = strip_bounds(2 * nx * ny, world, align=2 * ny)
    # the strips run on PARTITIONED meshes (own columns + halo), the single-GPU reference on the global mesh
    band = handler.mesh_band(handler.device_structured_channel(4, ny, 4 * lx / nx, ly, colmajor=True, device=device))
    c0, c1 = channel_slice_columns(nx, ny, bounds, rank, band, S)
    dm = handler.device_structured_channel(nx, ny, lx, ly, colmajor=True, device=device, col_lo=c0, col_hi=c1)
    F = field(dm)
    x = F[0]
    h = DistributedParticleHandler2D(dm, level, bounds, rank, world, migration=migration)
    W = (torch.zeros_like(x), torch.zeros_like(x))
    ref = RW = RF = None
    n_glob = (nx + 1) * (ny + 1)
    if rank == 0:
        gm = handler.device_structured_channel(nx, ny, lx, ly, colmajor=True, device=device)
        RF = field(gm)
        ref = handler.ParticleHandler2D(gm, level)
        RW = (torch.zeros_like(RF[0]), torch.zeros_like(RF[0]))
        ref.seed_particles()
        ref.init_particle_velocity(RF)
    h.seed_particles()
    h.init_particle_velocity(F)
    own = slice(int(bounds[rank]) - dm.cell_base, int(bounds[rank + 1]) - dm.cell_base)
    mine = torch.unique(dm.cells[own].to(torch.int64))  # nodes of this strip's own cells (slice numbering)
    worst = torch.zeros(1, dtype=torch.float64, device=device)
    migrated = 0
    for s in range(steps):
        h.step(F, W, dt, S)
        migrated += h.last_sent
        cs = h.state_checksum()
        if rank == 0:
            ref.step(RF, RW, dt, S)
            rcs = ref.state_checksum()
            if not torch.equal(cs, rcs):
                raise RuntimeError(f"multi-GPU parity: state checksum after step {s + 1} differs: {cs.tolist()} on {world} GPUs, {rcs.tolist()} on one")
        ref_w = [RW[0] if rank == 0 else torch.empty(n_glob, dtype=torch.float64, device=device),
                 RW[1] if rank == 0 else torch.empty(n_glob, dtype=torch.float64, device=device)]
        for t in ref_w:
            dist.broadcast(t, 0)
        for a, b in zip(W, ref_w):
            worst = torch.maximum(worst, (a[mine] - b[mine + dm.node_base]).abs().max() / b.abs().max().clamp_min(1e-300))
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    moved = torch.tensor([migrated], dtype=torch.int64, device=device)
    dist.all_reduce(moved)
    if float(worst) > 1e-12:
        raise RuntimeError(f"multi-GPU parity: projected nodal field differs by {float(worst):.3e} relative (bar 1e-12)")
    if int(moved) == 0:
        raise RuntimeError("multi-GPU parity: the check case migrated no particle")
    out = {"gpus": world, "steps": steps, "particles": int(cs[0]), "migrated": int(moved), "nodal_rel_err": float(worst),
           "state_checksum": [int(v) for v in cs.tolist()], "transport": h.protocol}
    h.close()
    if ref is not None:
        ref.close()
    out["spill"] = _spill_selfcheck(rank, world, device, migration)
    return out


def _spill_selfcheck(rank, world, device, migration):
    """SURVEY N4 across a strip boundary (see tests/mg_worker.py::spill_case): a particle in the tolerance band of the LAST cell
    of rank 0 sets an occupancy bit in the word of the FIRST cell of rank 1 and suppresses one re-seed there."""
    from . import handler

    nx, ny, level = 4 * world, 4, 2
    ppc = level * level
    dm = handler.device_structured_channel(nx, ny, 0.5 * world, 0.5, colmajor=True, device=device)
    zero = torch.zeros(dm.n_nodes, dtype=torch.float64, device=device)
    F, W = (zero, zero.clone()), (zero.clone(), zero.clone())
    bounds = strip_bounds(dm.n_cells, world, align=2 * ny)
    ref = handler.ParticleHandler2D(dm, level)
    ref.seed_particles()
    ref.init_particle_velocity(F)
    s = ref.download()  # seeded order: particle of (cell, sub-cell) at cell * ppc + sub-cell
    c = int(bounds[1]) - 1
    tri = dm.cells[c].cpu().numpy().view(np.uint32)
    v = dm.vertices.cpu().numpy()[tri.astype(np.int64)]
    L = np.array([0.3, -1.0e-6, 0.7 + 1.0e-6])
    pos = L[0] * v[0] + L[1] * v[1] + L[2] * v[2]
    keep = np.ones(s["x"].shape[0], dtype=bool)
    keep[(c + 1) * ppc + 0] = keep[(c + 1) * ppc + 1] = False
    st = {k: a[keep] for k, a in s.items()}
    add = {"x": pos[0], "y": pos[1], "l0": L[0], "l1": L[1], "l2": L[2], "vx": 0.0, "vy": 0.0, "cell": c, "id": 0}
    st = {k: np.concatenate([a, np.asarray([add[k]], dtype=a.dtype)]) for k, a in st.items()}
    ref.upload(st)
    ref.step(F, W, 0.01, 3)
    single = (ref.get_particle_count(), ref.stats()["added"])
    ref.close()
    h = DistributedParticleHandler2D(dm, level, bounds, rank, world, migration=migration)
    h.seed_particles()
    h.init_particle_velocity(F)
    own = (st["cell"] >= int(bounds[rank])) & (st["cell"] < int(bounds[rank + 1]))
    h.h.upload({k: a[own] for k, a in st.items()})
    h.step(F, W, 0.01, 3)
    multi = h.global_particle_count()
    h.close()
    if single[1] != 1:
        raise RuntimeError(f"multi-GPU parity: the crafted spill state re-seeded {single[1]} sub-cells on one GPU, expected 1")
    if multi != single[0]:
        raise RuntimeError(f"multi-GPU parity: {world} GPUs hold {multi} particles, one GPU {single[0]}: spill bits lost at the strip boundary")
    return {"single_gpu_count": single[0], "multi_gpu_count": multi}
