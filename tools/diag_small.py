import sys, os, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from gpupfem2_b200 import handler
args = argparse.Namespace(workload=sys.argv[1] if len(sys.argv) > 1 else "poiseuille", level=0, substeps=3, cfl=0.25, capacity_factor=1.3)
dm, level, F, dt = bench.build_problem(args, 0, 1, "cuda:0")
W = (torch.zeros_like(F[0]), torch.zeros_like(F[0]))
def run(tag, n=50, **opts):
    h = handler.ParticleHandler2D(dm, level, max_division_level=8, capacity_factor=1.3, **opts)
    h.seed_particles(); h.init_particle_velocity(F)
    for _ in range(3): h.step(F, W, dt, 3)
    h.get_particle_count(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    t = time.perf_counter(); e0.record()
    for k in range(n):
        t1 = time.perf_counter()
        h.step(F, W, dt, 3)
        t2 = time.perf_counter()
        h.get_particle_count()
        ts.append((t2 - t1, time.perf_counter() - t2))
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t) / n * 1e3
    ts.sort()
    print(f"{tag:28s} wall {wall:8.3f} ms/step  events {e0.elapsed_time(e1)/n:8.3f}  step-call median {ts[n//2][0]*1e3:.3f} max {ts[-1][0]*1e3:.3f} ms; count-call median {sorted(x[1] for x in ts)[n//2]*1e3:.3f}")
    h.close()
for rep in range(3):
    run("default (lazy re-sort)")
    run("physical re-sort", lazy_sort=False)
