"""In-situ drop-in check (run on the GPU box, needs oracle/_ref/Poiseuille_{ref,shim} built by gpupfem2_b200/shim):
the UNMODIFIED reference case cases/PoiseuilleFlow2D/main.cu, once linked against the reference library and once
against the library whose ParticleHandler2D is the B200 drop-in.  Compares the per-step particle counts printed by
advectParticles and the nodal fields of the exported solution files.

    python tests/insitu_compare.py [poiseuille|cylinder] -> gpurun_out/insitu_summary[_cylinder].json
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpupfem2_b200.mesh import HostMesh, write_dat  # noqa: E402


def parse_vtu(path):
    """{name: array} of the ASCII DataArrays in a .vtu written by the reference's DataExport."""
    txt = open(path).read()
    out = {}
    for m in re.finditer(r'<DataArray[^>]*Name="([^"]+)"[^>]*>(.*?)</DataArray>', txt, flags=re.S):
        try:
            out[m.group(1)] = np.array(m.group(2).split(), dtype=np.float64)
        except ValueError:
            pass
    return out


def binaries(case="poiseuille"):
    exename = {"poiseuille": "Poiseuille", "cylinder": "Cylinder"}[case]
    return [os.path.join(ROOT, "oracle", "_ref", f"{exename}_{tag}") for tag in ("ref", "shim")]


def compare(case="poiseuille"):
    """Run the unmodified case against the reference library and against the drop-in; -> summary dict."""
    fixture, datname, exename, suffix = {"poiseuille": ("mesh_channel.npz", "ChannelMesh.dat", "Poiseuille", ""),
                                         "cylinder": ("mesh_cylinder3.npz", "CylinderMesh3.dat", "Cylinder", "_cylinder")}[case]
    work = tempfile.mkdtemp(prefix="insitu_")
    d = np.load(os.path.join(ROOT, "tests", "golden", fixture))
    write_dat(os.path.join(work, datname), HostMesh(d["vertices"], d["cells"]))
    res = {}
    for tag in ("ref", "shim"):
        run = os.path.join(work, tag)
        os.makedirs(run)
        exe = os.path.join(ROOT, "oracle", "_ref", f"{exename}_{tag}")
        t0 = time.time()
        p = subprocess.run([exe], cwd=run, capture_output=True, text=True, timeout=3000)
        res[tag] = {"rc": p.returncode, "wall_s": time.time() - t0, "stdout": p.stdout, "stderr": p.stderr[-2000:], "dir": run}
    summary = {}
    counts = {}
    for tag in res:
        counts[tag] = [int(v) for v in re.findall(r"Particle handler contains (\d+) particles", res[tag]["stdout"])]
        step_ms = [float(v) for v in re.findall(r"Time of a simulation step:\s+([0-9.]+) ms", res[tag]["stdout"])]
        summary[tag] = {"rc": res[tag]["rc"], "wall_s": res[tag]["wall_s"], "steps": len(counts[tag]),
                        "created": re.findall(r"Created (\d+) particles", res[tag]["stdout"]),
                        "median_step_ms": float(np.median(step_ms)) if step_ms else None, "stderr_tail": res[tag]["stderr"][-300:]}
    n = min(len(counts["ref"]), len(counts["shim"]))
    same = [a == b for a, b in zip(counts["ref"][:n], counts["shim"][:n])]
    summary["count_steps_compared"] = n
    summary["count_identical_steps"] = int(sum(same))
    summary["first_count_mismatch_step"] = (same.index(False) + 1) if False in same else None
    summary["max_count_rel_diff"] = float(max((abs(a - b) / a for a, b in zip(counts["ref"][:n], counts["shim"][:n])), default=0.0))
    summary["final_counts"] = [counts["ref"][n - 1] if n else None, counts["shim"][n - 1] if n else None]
    fields = {}
    for k in (10, 100, 250, 490, 500):
        fa, fb = (os.path.join(res[t]["dir"], f"solution{k:04d}.vtu") for t in ("ref", "shim"))
        if not (os.path.exists(fa) and os.path.exists(fb)):
            cand = [f for f in os.listdir(res["ref"]["dir"]) if f.startswith("solution")]
            fields["files_seen"] = sorted(cand)[:5]
            continue
        A, B = parse_vtu(fa), parse_vtu(fb)
        fields[k] = {name: float(np.max(np.abs(A[name] - B[name])) / max(np.max(np.abs(A[name])), 1e-300))
                     for name in A if name in B and A[name].shape == B[name].shape and A[name].size}
    summary["field_rel_inf_diff"] = fields
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    summary["case"] = case
    with open(os.path.join(ROOT, "gpurun_out", f"insitu_summary{suffix}.json"), "w") as f:
        json.dump(summary, f, indent=1)
    with open(os.path.join(ROOT, "gpurun_out", f"insitu_counts{suffix}.txt"), "w") as f:
        for i in range(n):
            f.write(f"{i + 1} {counts['ref'][i]} {counts['shim'][i]}\n")
    return summary


if __name__ == "__main__":
    print(json.dumps(compare(sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] in ("poiseuille", "cylinder") else "poiseuille"), indent=1))
