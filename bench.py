#!/usr/bin/env python
"""bench.py -- particle-steps/sec of the PFEM-2 particle step on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = what the reference's cases do per time step to the particles:
    advectParticles(F, dt, S) [S x (advect + locate) + distribution check / re-seed]
    projectVelocityOntoGrid(W) ; correctParticleVelocity(F, W)
with a frozen synthetic nodal field F (the FEM stage is out of scope).  A particle-step is one live
particle carried through one such step; value = sum_k P(k) / device time.

Workloads (config.workload):
    channel16m  synthetic channel, 4000x2000 quads = 16M triangles, level 4 (16/cell, 256M particles):
                BASELINE.json configs[3], the configuration the metric is quoted on (default; fits one B200).
                --level 6 gives 36/cell (576M particles); the reference itself caps the level at 4.
    channel1m   synthetic channel, 1000x500 quads = 1M triangles x 16/cell   (configs[2])
    poiseuille  the shipped ChannelMesh (8756 triangles, level 2)              (configs[0], isolated mode)
    cylinder    the shipped CylinderMesh3 (34185 triangles, level 2)           (configs[1], isolated mode)

The JSON line carries `roofline` (dominant kernel, CUDA-event timed on the launching stream, against
MEASURED_PEAKS.json), `cpu_baseline` (the C++/OpenMP oracle on the host cores, bounded sample) and `e2e`
(the same metric through the C-ABI call pfem2_step_host with pinned HOST nodal buffers).

--impl reference runs the UNMODIFIED reference: its CUDA ParticleHandler2D (oracle/_ref, built from
/root/reference) on one B200 -- the reference has no CPU implementation of this path -- or, when that
build is absent, the oracle port on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES = {  # algorithmic bytes per particle per launch (64 B fp64 SoA state)
    # move pass = advect+locate (R(x,y,L,cell)=44 + W 44 taken alone) with the velocity correction of the previous step folded
    # in (pfem2_options.defer_correct; R(L,cell,v)=44 + W(v)=16 taken alone): the fused pass reads and writes every field ONCE,
    # R(x,y,L,cell,v)=60 + W 60
    "advect_locate": 120,
    "project_cells": 44,  # R(L,cell,v)
}
ALG_BYTES_STEP = 192  # the three calls as separate passes (88 + 44 + 60); the reference point of roofline.step
TRAFFIC_FILE = "r03_traffic.json"  # committed ncu --set full capture of the HEAD kernels (tools/traffic_from_ncu.py)

WORKLOADS = {
    # name: (nx, ny, lx, ly, default level)
    "channel16m": (4000, 2000, 20.0, 10.0, 4),
    "channel1m": (1000, 500, 10.0, 5.0, 4),
    # BASELINE.json configs[4]: 2M triangles PER GPU (nx grows with the GPU count: weak scaling), mean flow + a lattice of
    # Taylor-Green vortices at CFL ~0.9 per substep: multi-cell traversal (walks of 2-3 cells, jumps beyond the one-ring are
    # deleted) and strong depletion / accumulation that keeps the re-seeding busy
    "stress2m": (1000, 1000, 10.0, 10.0, 4),
}
WEAK = {"stress2m"}  # nx, lx scale with the number of GPUs


def nodal_field(args, x, y, lx, ly, umax):
    """Synthetic frozen nodal field of a workload from node coordinates (numpy arrays or torch tensors)."""
    if args.workload == "stress2m":
        import math
        mod = __import__("torch") if hasattr(x, "device") else __import__("numpy")
        k = 2.0 * math.pi / 0.5  # 50-cell vortices
        return (0.5 * umax + 0.5 * umax * mod.sin(k * x) * mod.cos(k * y), -0.5 * umax * mod.cos(k * x) * mod.sin(k * y))
    mod = __import__("torch") if hasattr(x, "device") else __import__("numpy")
    t = 4.0 * umax * y * (ly - y)
    # IEEE division like oracle/ref_harness.cu (on CUDA torch turns tensor / python-scalar into a multiplication by the rounded
    # reciprocal): both arms of the bench then advect through bit-identical nodal fields
    return (t / mod.full_like(t, ly * ly) if hasattr(x, "device") else t / (ly * ly), 0.0 * y)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

    Uses NVML in-process (nvidia_ml_py) from a background thread, one light query set every 50 ms.  A looping `nvidia-smi`
    process measurably slows the step down while it runs (small workloads: 0.2 -> 4 ms/step, channel16m e2e 18.8 -> 20.5 ms),
    so it is only the fallback when NVML cannot be imported."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period=float(os.environ.get("PFEM2_BENCH_NVML_PERIOD", "0.05"))):
        import threading

        self.rows, self.skip, self.p, self.f = [], 0, None, None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
            dev = pynvml.nvmlDeviceGetHandleByIndex(phys)
            smax = float(pynvml.nvmlDeviceGetMaxClockInfo(dev, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

            def loop():
                while not self._stop.is_set():
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(dev, pynvml.NVML_CLOCK_SM))
                        r = int(get_reasons(dev))
                        self.rows.append((sm, smax, [k for k, b in bits.items() if r & b]))
                    except Exception:
                        pass
                    self._stop.wait(period)

            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            try:
                self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                           "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
            except Exception:
                self.p = None

    def _smi_rows(self):
        out = []
        for r in open(self.f.name):
            t = r.strip().split(",")
            try:
                out.append((float(t[0]), float(t[1]), [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                                                            "sw_power_cap"), t[3:7]) if "Active" in v and "Not" not in v]))
            except Exception:
                continue
        return out

    def wait_first_sample(self, timeout=4.0):
        t0 = time.time()
        while time.time() - t0 < timeout:
            if (self._thread and self.rows) or (self.p is not None and os.path.getsize(self.f.name) > 0) or (not self._thread and self.p is None):
                return
            time.sleep(0.01)

    def mark(self):
        """Samples taken before this call (warm-up) are not reported."""
        self.skip = len(self.rows) if self._thread else (len(self._smi_rows()) if self.p is not None else 0)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self._thread:
            self._stop.set()
            self._thread.join(timeout=2)
            rows = list(self.rows)
            out["source"] = "NVML in-process, 50 ms period"
        elif self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
            self.f.flush()
            rows = self._smi_rows()
            os.unlink(self.f.name)
            out["source"] = "nvidia-smi -lms 200"
        else:
            return out
        rows = rows[self.skip:] if len(rows) > self.skip else rows[-1:]  # a run shorter than one period keeps the last sample
        if rows:
            out["sm_mhz"] = statistics.median(r[0] for r in rows)
            out["sm_max_mhz"] = rows[-1][1]
            out["reasons"] = sorted({n for r in rows for n in r[2]})
            out["samples"] = len(rows)
        return out


def bind_to_gpu_local_cpus(index):
    """Run this process on the CPUs NVML reports as local to GPU `index` (same NUMA node / PCIe root), so that the pinned host
    buffers of the e2e leg are allocated next to the GPU.  Returns the previous affinity mask (to restore) or None if nothing
    was changed.  PFEM2_BENCH_AFFINITY=0 disables it."""
    if os.environ.get("PFEM2_BENCH_AFFINITY", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
        dev = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(dev, (os.cpu_count() + 63) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cur = os.sched_getaffinity(0)
        want = local & cur  # never leave the cpuset the container grants
        if want and want != cur:
            os.sched_setaffinity(0, want)
            return cur
    except Exception:
        pass
    return None


def channel_params(args, world=1):
    nx, ny, lx, ly, lvl = WORKLOADS[args.workload]
    if args.workload in WEAK:
        nx, lx = nx * world, lx * world
    if args.workload == "stress2m":
        if args.cfl == 0.25:
            args.cfl = 0.9
        if args.capacity_factor == 1.3:
            args.capacity_factor = 3.0  # the reference only ever adds particles: this case doubles its count in ~20 steps
    level = args.level or lvl
    h = lx / nx
    umax = 1.0
    dt = args.cfl * h * args.substeps / umax  # CFL per substep at the channel centre
    return nx, ny, lx, ly, level, umax, dt


def workload_description(args, world=1):
    if args.workload in WORKLOADS:
        nx, ny, lx, ly, level, umax, dt = channel_params(args, world)
        field = "mean flow + Taylor-Green vortex lattice (50-cell vortices)" if args.workload == "stress2m" else "Poiseuille field"
        return (f"{args.workload}: structured channel {nx}x{ny} quads = {2 * nx * ny} triangles, level {level} "
                f"({level * level}/cell, {2 * nx * ny * level * level} particles seeded), {field}, "
                f"S={args.substeps}, CFL/substep={args.cfl}")
    return f"{args.workload}: shipped mesh (tests/golden fixture), level {args.level or 2}, S={args.substeps}, isolated mode"


# ------------------------------------------------------------------------------------------------
def build_problem(args, rank, world, device):
    """-> (DeviceMesh, level, F (fx, fy) device tensors, dt)."""
    import numpy as np
    import torch

    from gpupfem2_b200 import handler

    if args.workload in WORKLOADS:
        nx, ny, lx, ly, level, umax, dt = channel_params(args, world)
        if world > 1:
            # partitioned mesh: this rank generates only its own quad columns + the halo a particle can reach in one advect
            from gpupfem2_b200 import multi_gpu

            band = handler.mesh_band(handler.device_structured_channel(4, ny, 4 * lx / nx, ly, colmajor=True, device=device))
            bounds = multi_gpu.strip_bounds(2 * nx * ny, world, align=2 * ny)
            c0, c1 = multi_gpu.channel_slice_columns(nx, ny, bounds, rank, band, args.substeps)
            dm = handler.device_structured_channel(nx, ny, lx, ly, colmajor=True, device=device, col_lo=c0, col_hi=c1)
        else:
            dm = handler.device_structured_channel(nx, ny, lx, ly, colmajor=True, device=device)
        fx, fy = nodal_field(args, dm.vertices[:, 0].contiguous(), dm.vertices[:, 1].contiguous(), lx, ly, umax)
        return dm, level, (fx.contiguous(), fy.contiguous()), dt
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from gpupfem2_b200.mesh import HostMesh

    name = {"poiseuille": "channel", "cylinder": "cylinder3"}[args.workload]
    d = np.load(os.path.join(ROOT, "tests", "golden", f"mesh_{name}.npz"))
    hm = HostMesh(d["vertices"], d["cells"])
    dm = handler.DeviceMesh(hm, device=device)
    yv = hm.vertices[:, 1]
    if args.workload == "poiseuille":  # cases/PoiseuilleFlow2D: dt = 0.01, analytic parabola 0.5 y (1 - y)
        fx, dt = 0.5 * yv * (1.0 - yv), 0.01
    else:  # cases/Cylinder2D: dt = 0.001, inflow profile 4 U y (H - y) / H^2 with U = 1.5, H = 0.41
        fx, dt = 4.0 * 1.5 * yv * (0.41 - yv) / (0.41 * 0.41), 0.001
    fx = torch.as_tensor(np.ascontiguousarray(fx)).to(device)
    return dm, args.level or 2, (fx, torch.zeros_like(fx)), dt


def traffic_table():
    """DRAM bytes per particle of every kernel of the step, from the committed ncu --set full capture of the HEAD kernels
    (profiles/TRAFFIC_FILE: dram__bytes_read.sum + dram__bytes_write.sum per launch on channel16m)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
    except Exception:
        return None


def measure(args, rank, world, local, full):
    """One workload on `world` GPUs (this process = rank `rank`): warm-up, K device-timed steps, roofline, and -- full=True, the
    headline -- the end-to-end leg with HOST nodal buffers.  Returns the JSON dict on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist

    from gpupfem2_b200 import handler

    device = f"cuda:{local}"
    multi = world > 1
    t_setup = time.time()
    dm, level, F, dt = build_problem(args, rank, world, device)
    W = (torch.zeros_like(F[0]), torch.zeros_like(F[0]))
    lazy = bool(int(os.environ.get("PFEM2_LAZY_SORT", "1")))  # A/B switch: 0 = physical re-sort in every advect
    opts = dict(max_division_level=8, capacity_factor=args.capacity_factor, lazy_sort=lazy)
    def make_handler():
        if multi:
            from gpupfem2_b200 import multi_gpu

            ny = WORKLOADS[args.workload][1]
            bounds = multi_gpu.strip_bounds(dm.n_cells_global, world, align=2 * ny)
            hh = multi_gpu.DistributedParticleHandler2D(dm, level, bounds, rank, world, **opts)
            return hh, hh.h
        hh = handler.ParticleHandler2D(dm, level, host_pipeline=int(os.environ.get("PFEM2_HOST_PIPELINE", "0")),
                                       graph_advect=int(os.environ.get("PFEM2_GRAPH_ADVECT", "0")), **opts)
        return hh, hh

    h, inner = make_handler()
    h.seed_particles()
    h.init_particle_velocity(F)
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup
    small = inner.get_particle_count() * 64 < 256e6  # state could sit in the 126 MB L2 -> flush between timed iterations
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device=device) if small else None

    sampler = ClockSampler(local) if rank == 0 else None  # started before the warm-up so that its start-up (NVML init) is over
    for _ in range(args.warmup):
        h.step(F, W, dt, args.substeps)
    h.get_particle_count()
    torch.cuda.synchronize()
    if sampler:
        sampler.wait_first_sample()
    if multi:
        dist.barrier()
    # Per-phase CUDA events ride along in the timed region of the HBM-bound workloads.  The small shipped meshes are launch-bound and
    # run advectParticles as one CUDA-graph launch, which the phase events would switch off: they are timed plain, and the phase
    # breakdown comes from a second, profiled pass behind the timed region.
    inner.set_profiling(not small)
    inner.phase_times(reset=True)
    launches0 = handler.kernel_launches()
    if sampler:
        sampler.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    counts = []
    torch.cuda.synchronize()
    if multi:
        dist.barrier()
    for k in range(args.steps):
        if flush is not None:
            flush.fill_(float(k))
        ev[k][0].record()
        h.step(F, W, dt, args.substeps)
        ev[k][1].record()
        counts.append(h.get_particle_count())  # waits only for the advect's counter read-back
    torch.cuda.synchronize()
    if multi:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    launches = handler.kernel_launches() - launches0
    ms = [a.elapsed_time(b) for a, b in ev]
    # large states: one interval over the K steps (what lies between two steps is part of the job); flushed small states: the
    # flush sits between the per-step pairs
    total_ms = ev[0][0].elapsed_time(ev[-1][1]) if flush is None else sum(ms)
    if small:
        inner.set_profiling(True)
        inner.phase_times(reset=True)
        for k in range(args.steps):
            h.step(F, W, dt, args.substeps)
        h.get_particle_count()
    phases = inner.phase_times(reset=True)
    inner.set_profiling(False)
    checksum = h.state_checksum()  # order-independent, global (all-reduced over the strips): identical for every N
    t = torch.tensor([total_ms, float(sum(counts)), float(h.last_sent if multi else 0)], dtype=torch.float64, device=device)
    if multi:
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, psteps, sent_all = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        psteps = float(sum(counts))
    value = psteps / (total_ms * 1e-3)

    # roofline of the dominant kernel (by device time) among the algorithmic passes; rank 0's GPU
    peak, peak_src = peaks()
    pmean_rank = sum(counts) / args.steps
    per_phase = {}
    for name, (pms, calls) in phases.items():
        per_phase[name] = {"ms_per_step": pms / args.steps, "share": pms / total_ms if total_ms else 0.0}
        if name in ALG_BYTES and pms > 0:
            per_phase[name]["alg_GBps"] = ALG_BYTES[name] * pmean_rank / (pms / args.steps * 1e-3) / 1e9
    dom = max(ALG_BYTES, key=lambda n: phases[n][0])
    dom_ms = phases[dom][0] / args.steps
    achieved = ALG_BYTES[dom] * pmean_rank / (dom_ms * 1e-3) / 1e9
    tt = traffic_table() if args.workload == "channel16m" and (args.level or 4) == 4 and lazy else None
    traffic = step_traffic = None
    if tt:
        traffic = tt["phases"][dom]["bytes_per_particle"] * pmean_rank
        step_traffic = tt["step"]["bytes_per_particle"] * pmean_rank
    step_ach = ALG_BYTES_STEP * value / 1e9 / world
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "traffic_source": (f"profiles/{TRAFFIC_FILE} (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch at HEAD, scaled to "
                                   "this run's particle count)") if traffic else None,
                "peak_source": peak_src, "alg_bytes_per_particle": ALG_BYTES[dom],
                "alg_bytes_note": "move pass: R(x,y,L,cell,v) 60 + W 60 with the velocity correction of the previous step folded in "
                                  "(each field read and written once); projection: R(L,cell,v) 44",
                "step": {"achieved": step_ach, "frac": step_ach / peak, "alg_bytes_per_particle_step": ALG_BYTES_STEP,
                         "traffic": step_traffic, "note": "per GPU" if multi else None},
                "phases": per_phase}
    if multi:
        roofline["note"] = "rank 0, per GPU"

    e2e = None
    if full:
        # e2e: the same step with pinned HOST nodal buffers (H2D + D2H inside the timed region).  One GPU: the C-ABI call
        # pfem2_step_host; N GPUs: every rank stages its own slice of the nodal field (DistributedParticleHandler2D.step_host)
        prev_affinity = bind_to_gpu_local_cpus(local)  # pinned buffers next to the GPU (NUMA); restored before the CPU baseline runs
        hF = [t_.cpu().pin_memory() for t_ in F]
        hW = [torch.empty_like(t_).pin_memory() for t_ in hF]
        if multi:
            step_host = lambda: h.step_host(hF, hW, F, W, dt, args.substeps)  # noqa: E731
            h2d, d2h = h.host_bytes_per_step(args.substeps)
            hb = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=device)
            dist.all_reduce(hb)
            h2d, d2h = int(hb[0]), int(hb[1])
            api = ("pfem2_step_host_p2p (C ABI, one call per rank and step): the slice of the pinned host nodal field the strip's advect reads is "
                   "uploaded in chunks under the move pass, migration and halo sums over NVLink peer memory, the projected slice the strip "
                   "owns is downloaded under the projection (interface nodes after the halo sum)")
        else:
            step_host = lambda: h.step_host(hF[0], hF[1], hW[0], hW[1], dt, args.substeps)  # noqa: E731
            h2d, d2h = 2 * dm.n_nodes * 8, 2 * dm.n_nodes * 8 + 32
            api = ("pfem2_step_host (C ABI), pinned host nodal buffers in, projected nodal field + count out; the call pipelines the upload "
                   "with the move pass and the projection with the download in chunks of the cell range (pfem2_options.host_pipeline)")
        # the SAME steps of the same simulation as the device-timed leg (the flow develops: particle count and cost per step grow
        # with the step number), so that e2e and `value` differ by the host <-> device traffic only: a fresh handler, the same
        # warm-up steps (untimed, through the same call), then the K timed steps
        h.close()
        del h, inner
        torch.cuda.empty_cache()
        h, inner = make_handler()
        h.seed_particles()
        h.init_particle_velocity(F)
        for _ in range(args.warmup):
            step_host()
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        e2e_counts, e2e_ms = [], []
        t0 = time.perf_counter()
        for k in range(args.steps):
            t1 = time.perf_counter()
            e2e_counts.append(step_host())  # (returns when the projected field and the count are on the host)
            e2e_ms.append((time.perf_counter() - t1) * 1e3)
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        if prev_affinity:
            os.sched_setaffinity(0, prev_affinity)
        te = torch.tensor([t_e2e, float(sum(e2e_counts))], dtype=torch.float64, device=device)
        if multi:
            a, b = te.clone(), te.clone()
            dist.all_reduce(a, op=dist.ReduceOp.MAX)
            dist.all_reduce(b, op=dist.ReduceOp.SUM)
            t_e2e, e2e_ps = float(a[0]), float(b[1])
        else:
            e2e_ps = float(sum(e2e_counts))
        e2e = {"value": e2e_ps / t_e2e, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": t_e2e / args.steps * 1e3, "ms_per_step_min": min(e2e_ms), "ms_per_step_median": statistics.median(e2e_ms),
               "ms_per_step_max": max(e2e_ms), "timing": "host clock around K calls, synchronised on both sides, max over ranks",
               "host_cpus": ("bound to the GPU-local CPUs (NVML affinity)" if prev_affinity else "process default"), "api": api}
    mesh_bytes = sum(int(t_.numel()) * t_.element_size() for t_ in (dm.vertices, dm.cells, dm.inv_jacobi, dm.nbr_offsets, dm.nbr_indices))
    h.close()
    del h, inner
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    state_gb = pmean_rank * 64 / 1e9
    sm = sorted(ms)
    out = {
        "metric": "particle-steps/sec (advect+locate+sort+project+correct)", "value": value, "unit": "particle-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "ms_per_step_min": sm[0], "ms_per_step_median": statistics.median(sm), "ms_per_step_list": [round(v, 3) for v in ms],
        "higher_is_better": True, "scaling": "weak" if args.workload in WEAK else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_description(args, world) + (f", strip-partitioned over {world} GPUs (quad columns)" if multi else ""),
                   "particles_mean": psteps / args.steps, "cells": dm.n_cells_global or dm.n_cells, "nodes": dm.n_nodes_global or dm.n_nodes,
                   "cells_per_rank_mesh": dm.n_cells,
                   "substeps": args.substeps, "dt": dt, "lazy_sort": lazy,
                   "l2": (f"flushed between timed iterations (512 MiB write; state {state_gb:.3f} GB could fit the 126 MB L2)"
                          if small else f"inputs larger than L2 ({state_gb:.1f} GB of particle state per pass and GPU)"),
                   "timing": "CUDA events on the launching (legacy default) stream around the K steps" + (", max over ranks, barrier on both sides" if multi else ""),
                   "mesh_bytes_per_rank": mesh_bytes, "setup_s": t_setup},
        "roofline": roofline, "gpu_launches": int(launches), "clocks": clocks,
        "gpu_launches_note": ("kernel launches enqueued by the library in the timed region; a replayed CUDA graph of advectParticles "
                              "(small meshes) counts as the kernels it contains"),
        "state_checksum": {"value": [int(v) for v in checksum.tolist()],
                           "what": "order-independent wrapping int64 sums over the GLOBAL particle set after warm-up + timed steps: "
                                   "[count, bits(x), bits(y), bits(L0), bits(L1), bits(L2), cell]; identical for every N"},
    }
    if e2e:
        out["e2e"] = e2e
    if multi:
        out["config"]["migrated_particles_per_step"] = sent_all
        out["config"]["migration_protocol"] = h_protocol_description(os.environ.get("PFEM2_MG_PROTOCOL", "p2p"))
    return out


def h_protocol_description(protocol):
    return {"p2p": "p2p: NVLink peer memory (CUDA IPC): emigrant records and interface-node accumulators stored straight into the neighbour "
                   "strip's HBM, device-side sequence flags; no NCCL and no host in the loop",
            "neighbour": "neighbour: fixed-size migration buffers [header | records] to / from the adjacent strips (ncclSend / ncclRecv, "
                         "counts stay on the device) + pairwise isend/irecv of interface-node accumulators (NCCL)",
            "exact": "exact: all_to_all_single (counts, 64-byte particle records) + pairwise isend/irecv of interface-node accumulators "
                     "(NCCL)"}.get(protocol, protocol)


EXTRAS = (  # BASELINE.json's other headline sizes, nested under "extra" of the headline line at every N
    ("channel16m_level6", {"workload": "channel16m", "level": 6},
     "configs[3] bracket: 36/cell = 576M particles (BASELINE quotes 32/cell = 512M, not expressible in the reference: levels are squares)"),
    ("stress2m", {"workload": "stress2m", "level": 0}, "configs[4]: high-CFL vortex lattice, 2M triangles PER GPU (weak scaling)"),
    # north_star names NCCL send/recv for the migration and NCCL for the halo sums; the default transport is NVLink peer memory
    # without NCCL, so the NCCL form of the same headline workload is measured next to it (multi-GPU launches only)
    ("channel16m_nccl", {"workload": "channel16m", "level": 0, "protocol": "neighbour"},
     "the headline workload with the NCCL transport (ncclSend / ncclRecv migration buffers, NCCL halo sums) instead of peer-memory stores"),
)


def run_ours(args):
    import copy

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the particle step has no CPU fallback")
    torch.cuda.set_device(local)
    parity = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"  # stdout carries the ONE JSON line
        if args.workload not in WORKLOADS:
            raise SystemExit("multi-GPU bench runs the synthetic channel workloads")
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        from gpupfem2_b200 import multi_gpu

        # parity gate inside the same launch: N strips against one GPU on a small case (+ the N4 spill case); a mismatch raises
        parity = multi_gpu.parity_selfcheck(rank, world, f"cuda:{local}")
    out = measure(args, rank, world, local, full=True)
    extras = {}
    if args.extra and args.workload == "channel16m" and not args.level:
        for name, over, what in EXTRAS:
            if "protocol" in over and (world == 1 or os.environ.get("PFEM2_MG_PROTOCOL", "p2p") == over["protocol"]):
                continue
            a = copy.copy(args)
            a.workload, a.level = over["workload"], over["level"]
            a.cfl, a.capacity_factor = 0.25, 1.3  # (channel_params derives the stress case's own values from these defaults)
            a.steps, a.warmup = min(args.steps, 5), 3
            saved = os.environ.get("PFEM2_MG_PROTOCOL")
            if "protocol" in over:
                os.environ["PFEM2_MG_PROTOCOL"] = over["protocol"]
            try:
                line = measure(a, rank, world, local, full=False)
                if line:
                    line["what"] = what
            except Exception as e:  # an extra must never take the headline down
                line = {"error": f"{type(e).__name__}: {e}"[:300], "what": what}
            finally:
                if "protocol" in over:
                    if saved is None:
                        os.environ.pop("PFEM2_MG_PROTOCOL", None)
                    else:
                        os.environ["PFEM2_MG_PROTOCOL"] = saved
            if rank == 0:
                extras[name] = line
    if rank == 0:
        if parity:
            out["parity_check"] = "ok"
            out["parity"] = parity
        if extras:
            out["extra"] = extras
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def oracle_run(nx, ny, lx, ly, level, substeps, cfl, steps, warmup, args=None):
    """Time the C++/OpenMP oracle (test infrastructure, used here only as the reported CPU baseline)."""
    from gpupfem2_b200.mesh import poiseuille_field, structured_channel
    from oracle import oracle as orc
    import numpy as np

    m = orc.complete_mesh(structured_channel(nx, ny, lx, ly, colmajor=True))
    if args is not None and args.workload == "stress2m":
        fx, fy = nodal_field(args, m.vertices[:, 0], m.vertices[:, 1], lx, ly, 1.0)
        fx, fy = np.ascontiguousarray(fx), np.ascontiguousarray(fy)
    else:
        fx, fy = poiseuille_field(m, 1.0, ly)
    dt = cfl * (lx / nx) * substeps
    o = orc.OracleHandler(m, level, max_level=8)
    o.seed_particles()
    o.init_particle_velocity(fx, fy)
    wx, wy = np.zeros_like(fx), np.zeros_like(fx)
    for _ in range(warmup):
        o.step(fx, fy, wx, wy, dt, substeps)
    t0 = time.perf_counter()
    ps = 0
    for _ in range(steps):
        ps += o.step(fx, fy, wx, wy, dt, substeps)
    t = time.perf_counter() - t0
    return ps / t, t / steps * 1e3, orc.lib().orc_max_threads(), m.n_cells, ps / steps


def cpu_baseline(args):
    nx, ny, lx, ly = 1000, 500, 10.0, 5.0  # bounded sample: the 1M-triangle channel, same field / CFL / level
    level = args.level or 4
    if args.workload in ("poiseuille", "cylinder"):
        nx, ny, lx, ly, level = 200, 100, 10.0, 5.0, 2
    if args.workload == "stress2m":
        nx, ny, lx, ly = 700, 700, 7.0, 7.0  # same cell size and vortex lattice, ~1M triangles
        channel_params(args)  # sets the workload's CFL
    try:
        v, ms, cores, cells, p = oracle_run(nx, ny, lx, ly, level, args.substeps, args.cfl, 3, 1, args)
    except Exception as e:  # the oracle is optional infrastructure for this leg
        return {"value": None, "unit": "particle-steps/s", "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}
    return {"value": v, "unit": "particle-steps/s", "cores": cores, "kind": "port", "ms_per_step": ms,
            "sample": f"C++/OpenMP oracle, {cells} triangles x {level * level}/cell ({int(p)} particles), 3 steps after 1 warm-up"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    base = {"metric": "particle-steps/sec (advect+locate+sort+project+correct)", "unit": "particle-steps/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": workload_description(args)}}
    if args.workload in WORKLOADS and args.workload != "stress2m" and os.path.exists(exe):
        nx, ny, lx, ly, level, umax, dt = channel_params(args)
        level = min(level, 4)  # CONSTANTS::MAX_CELL_DIVISION_LEVEL
        cmd = [exe, "time", str(nx), str(ny), repr(lx), repr(ly), str(level), str(args.substeps), repr(dt), repr(umax),
               str(args.steps), str(args.warmup), "1"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=3000)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode == 0 and line:
            j = json.loads(line[-1])
            v = j["particle_steps_per_s"]
            base.update({"value": v, "ms_per_step": j["ms_per_step"],
                         "cpu_baseline": {"value": v, "unit": "particle-steps/s", "cores": 1, "kind": "reference",
                                          "sample": ("the reference's own CUDA ParticleHandler2D (it has no CPU implementation of this "
                                                     "path), unmodified sources rebuilt for sm_100a, full workload on ONE B200, "
                                                     f"{j['particles']} particles")},
                         "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
            base["config"]["reference_cells"] = j["cells"]
            print(json.dumps(base))
            return
        sys.stderr.write(f"ref_harness failed (rc={r.returncode}): {r.stderr[-400:]}\n")
    # fall back to the oracle port on the host cores (bounded sample)
    cb = cpu_baseline(args)
    cb_line = dict(cb)
    base.update({"value": cb["value"], "ms_per_step": cb.get("ms_per_step"), "cpu_baseline": cb_line,
                 "e2e": {"value": cb["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="channel16m", choices=list(WORKLOADS) + ["poiseuille", "cylinder"])
    ap.add_argument("--level", type=int, default=0)
    ap.add_argument("--substeps", type=int, default=3)
    ap.add_argument("--cfl", type=float, default=0.25, help="CFL per substep at the channel centre")
    ap.add_argument("--capacity-factor", type=float, default=1.3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", dest="extra", action="store_false", help="skip the nested level-6 / stress2m lines")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "ours" and args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself under torchrun, one rank per GPU (the driver's form)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", os.environ.get("MASTER_PORT", "29517"), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
