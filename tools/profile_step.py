"""One particle step of a bench workload between cudaProfilerStart / Stop, for ncu --profile-from-start off:

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/step_kernels.csv python tools/profile_step.py [--workload channel16m] [--level 0]
    python tools/traffic_from_ncu.py gpurun_out/step_kernels.csv gpurun_out/step_particles.json profiles/r02_traffic.json

Writes gpurun_out/step_particles.json = {"particles": live count during the profiled step, ...}."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from gpupfem2_b200 import handler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="channel16m")
ap.add_argument("--level", type=int, default=0)
ap.add_argument("--substeps", type=int, default=3)
ap.add_argument("--cfl", type=float, default=0.25)
ap.add_argument("--capacity-factor", type=float, default=1.3)
ap.add_argument("--warmup", type=int, default=4)
ap.add_argument("--steps", type=int, default=1)
args = ap.parse_args()
dm, level, F, dt = bench.build_problem(args, 0, 1, "cuda:0")
W = (torch.zeros_like(F[0]), torch.zeros_like(F[0]))
h = handler.ParticleHandler2D(dm, level, max_division_level=8, capacity_factor=args.capacity_factor,
                              lazy_sort=bool(int(os.environ.get("PFEM2_LAZY_SORT", "1"))))
h.seed_particles()
h.init_particle_velocity(F)
for _ in range(args.warmup):
    h.step(F, W, dt, args.substeps)
n0 = h.get_particle_count()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.steps):
    h.step(F, W, dt, args.substeps)
n1 = h.get_particle_count()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"particles": n0, "particles_after": n1, "workload": bench.workload_description(args), "steps": args.steps},
          open("gpurun_out/step_particles.json", "w"))
print("profiled", args.steps, "step(s) with", n0, "particles")
