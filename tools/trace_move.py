"""Per-warp / per-tile timing of the gathered move pass at chosen steps of a bench workload (diagnosis build of the library):

    make -C gpupfem2_b200/csrc variant NAME=trace DEFS=-DPFEM2_MOVE_TRACE
    PFEM2_LIB_PATH=$PWD/gpupfem2_b200/_variants/libpfem2_trace.so python tools/trace_move.py --at 4,20,40

Writes gpurun_out/trace_move_<step>.npz = {warp: [warps, 4] (t_start, t_end in ns, SM id, tiles), tile: [tiles] ns per tile iteration,
cell_start: segment table of the sorted order the pass read} and prints the spread of the per-SM finish times."""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from gpupfem2_b200 import _lib, handler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="channel16m")
ap.add_argument("--level", type=int, default=0)
ap.add_argument("--substeps", type=int, default=3)
ap.add_argument("--cfl", type=float, default=0.25)
ap.add_argument("--capacity-factor", type=float, default=1.3)
ap.add_argument("--at", default="4,20,40")
ap.add_argument("--upto", type=int, default=0, help="trace every step up to this one, keep only those with a straggler (> 5 %% behind the median SM)")
args = ap.parse_args()
at = sorted(int(x) for x in args.at.split(","))
if args.upto:
    at = list(range(2, args.upto + 1))
lib = _lib.load()
lib.pfem2_debug_move_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
dm, level, F, dt = bench.build_problem(args, 0, 1, "cuda:0")
W = (torch.zeros_like(F[0]), torch.zeros_like(F[0]))
h = handler.ParticleHandler2D(dm, level, max_division_level=8, capacity_factor=args.capacity_factor)
h.seed_particles()
h.init_particle_velocity(F)
n_warps = 148 * 4 * 8
wbuf = torch.zeros(n_warps * 4, dtype=torch.int64, device="cuda:0")
tbuf = torch.zeros((1 << 24, 4), dtype=torch.int32, device="cuda:0")
os.makedirs("gpurun_out", exist_ok=True)
for step in range(at[-1] + 1):
    if step in at:
        wbuf.zero_()
        tbuf.zero_()
        assert lib.pfem2_debug_move_trace(wbuf.data_ptr(), tbuf.data_ptr()) == 0
    else:
        assert lib.pfem2_debug_move_trace(None, None) == 0
    n0 = h.get_particle_count()
    cs_before = h.cell_starts().clone() if step in at else None  # the segment table of the sorted order this pass reads
    h.step(F, W, dt, args.substeps)
    torch.cuda.synchronize()
    if step in at:
        w = wbuf.cpu().numpy().reshape(-1, 4)
        tiles = (n0 + 31) // 32
        t0 = w[:, 0].min()
        end = w[:, 1] - t0
        sm = w[:, 2]
        per_sm = np.array([end[sm == s].max() for s in range(148)])
        if args.upto and per_sm.max() < 1.05 * np.median(per_sm):
            print(f"step {step}: particles {n0}, pass {end.max() / 1e6:.3f} ms, balanced")
            continue
        parts = tbuf[:tiles].cpu().numpy().astype(np.int64)
        t = parts.sum(axis=1)
        print(f"step {step}: particles {n0}, pass {end.max() / 1e6:.3f} ms; per-SM finish min/median/max "
              f"{per_sm.min() / 1e6:.3f} / {np.median(per_sm) / 1e6:.3f} / {per_sm.max() / 1e6:.3f} ms; "
              f"per-warp finish p1/p50/p99 {np.percentile(end, 1) / 1e6:.3f} / {np.percentile(end, 50) / 1e6:.3f} / {np.percentile(end, 99) / 1e6:.3f}; "
              f"tile ns mean {t.mean():.0f} p50 {np.percentile(t, 50):.0f} p99 {np.percentile(t, 99):.0f} max {t.max()}")
        # the stragglers: where in their sequence they are slow, and in which part of the iteration the time goes
        K = tiles // n_warps
        P = parts[:K * n_warps].reshape(K, n_warps, 4)
        T = P.sum(axis=2)
        late = np.argsort(end)[::-1][:3]
        ref = np.median(P, axis=1)  # [K, 4]: the median warp at the same round
        keep = {}
        for lw in late:
            if end[lw] < 1.03 * np.median(end):
                continue
            slow = T[:, lw] > 1.25 * np.median(T, axis=1)
            ks = np.nonzero(np.convolve(slow, np.ones(25) / 25, mode="same") > 0.6)[0]
            rng = (int(ks.min()), int(ks.max())) if len(ks) else (0, 0)
            sel = slice(rng[0], rng[1] + 1)
            print(f"   warp {lw} (block {lw // 8}, warp {lw % 8} of it, SM {sm[lw]}) ends {end[lw] / 1e6:.3f} ms; slow rounds {rng} of {K}; "
                  f"parts [top+issue, wait tile, move, store+stats] slow range: {P[sel, lw].mean(axis=0).round(0)} vs median warp {ref[sel].mean(axis=0).round(0)}; "
                  f"outside: {np.delete(P[:, lw], np.r_[sel], axis=0).mean(axis=0).round(0)}")
            keep[f"warp{lw}"] = P[:, lw].astype(np.uint32)
            keep[f"sib{lw}"] = P[:, lw ^ 1].astype(np.uint32)
        np.savez_compressed(f"gpurun_out/trace_move_{step}.npz", warp=w, particles=n0, median_parts=ref.astype(np.uint32), **keep)
