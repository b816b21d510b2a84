"""How reproducible is the coupled Cylinder2D run itself?  Runs the UNMODIFIED reference binary twice and the drop-in in three
configurations, and prints the first step at which the per-step particle counts of each pair differ (the reference's projection uses
atomics and its Krylov solvers parallel reductions, so the coupled run is not bit-reproducible run to run)."""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpupfem2_b200.mesh import HostMesh, write_dat  # noqa: E402

work = tempfile.mkdtemp(prefix="insitu_var_")
d = np.load(os.path.join(ROOT, "tests", "golden", "mesh_cylinder3.npz"))
write_dat(os.path.join(work, "CylinderMesh3.dat"), HostMesh(d["vertices"], d["cells"]))
runs = [("ref_a", "Cylinder_ref", {}), ("ref_b", "Cylinder_ref", {}), ("shim_default", "Cylinder_shim", {}),
        ("shim_nograph", "Cylinder_shim", {"PFEM2_GRAPH_ADVECT": "-1"}), ("shim_physical", "Cylinder_shim", {"PFEM2_LAZY_SORT": "0", "PFEM2_GRAPH_ADVECT": "-1"})]
counts, ms = {}, {}
for tag, exe, env in runs:
    run = os.path.join(work, tag)
    os.makedirs(run)
    p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", exe)], cwd=run, capture_output=True, text=True, timeout=3000, env=dict(os.environ, **env))
    counts[tag] = [int(v) for v in re.findall(r"Particle handler contains (\d+) particles", p.stdout)]
    t = [float(v) for v in re.findall(r"Time of a simulation step:\s+([0-9.]+) ms", p.stdout)]
    ms[tag] = float(np.median(t)) if t else None
    print(tag, "rc", p.returncode, "steps", len(counts[tag]), "final", counts[tag][-1] if counts[tag] else None, "median step ms", ms[tag], flush=True)
tags = [r[0] for r in runs]
for i, a in enumerate(tags):
    for b in tags[i + 1:]:
        n = min(len(counts[a]), len(counts[b]))
        diff = [k + 1 for k in range(n) if counts[a][k] != counts[b][k]]
        print(f"{a} vs {b}: first differing step {diff[0] if diff else None}, differing steps {len(diff)} of {n}, max |diff| {max((abs(counts[a][k] - counts[b][k]) for k in range(n)), default=0)}")
