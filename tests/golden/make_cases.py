"""Write the case files the reference harness consumes (run here, before a gpurun call):

    python tests/golden/make_cases.py            -> build/golden_cases/<case>.bin
    gpurun -- 'for c in build/golden_cases/*.bin; do n=$(basename $c .bin); \
               oracle/_ref/ref_harness dump $c gpurun_out/golden/$n; done'
    python tests/golden/pack_golden.py           -> tests/golden/<case>.npz
"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from cases import ALL_CASES, build_case  # noqa: E402
from gpupfem2_b200.casefile import write_case  # noqa: E402
from oracle import oracle as orc  # noqa: E402

out = os.path.join(ROOT, "build", "golden_cases")
os.makedirs(out, exist_ok=True)
for name in ALL_CASES:
    c = build_case(name)
    orc.complete_mesh(c.mesh)  # one-ring CSR (validated against the O(C^2) definition in tests)
    write_case(os.path.join(out, name + ".bin"), c.mesh, c.fx, c.fy, c.level, c.substeps, c.dt, c.nsteps, c.dump_steps)
    print(name, c.mesh.n_nodes, c.mesh.n_cells)
