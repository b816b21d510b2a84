"""Pack the dumps the reference harness wrote on the GPU box (gpurun_out/golden/) into fixtures.

    python tests/golden/pack_golden.py      -> tests/golden/ref_<case>.npz

Each fixture holds, per kept dump step, the reference's particle state in canonical order
(sorted by (cell, x, y): SURVEY N6) and its projected nodal field, plus the per-step particle counts.
Only dumps the reference produced while following its own deletion rule are kept: where its
kDeleteParticles race (DESIGN.md "delete race") fired, the run is truncated before it.
KEEP lists what was verified race-free (see DESIGN.md for the evidence).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from cases import build_case  # noqa: E402
from gpupfem2_b200.casefile import canonical_order, read_dump  # noqa: E402

# case -> (dump steps kept, number of leading steps whose particle counts are kept)
KEEP = {
    "tiny_l1": ((0, 1, 5, 20), 20),
    "tiny_l2": ((0, 1, 2, 5, 10, 20), 22),
    "tiny_l3": ((0, 1, 5), 9),
    "tiny_l4": ((0, 1, 5), 7),
    "tiny_box": ((1, 10, 30), 30),
    "tiny_fast": ((0, 1), 1),  # plus every step in the reference's own ARRAY order, see RAW_ORDER below
    "channel_l2": ((1, 50), 50),
    "channel_fast_rev": ((30,), 30),
    "cyl3_box": ((60,), 60),
}

# cases additionally stored in the reference's array order at every step (S = 1), so that a test can restart
# from each reference state, decide from the array order whether the delete race could fire in that step
# (a doomed particle among the last n slots) and demand exact equality whenever it could not.
COMPACT = ("cyl3_box",)
RAW_ORDER = {"tiny_fast": tuple(range(0, 11))}

src = os.path.join(ROOT, "gpurun_out", "golden")
out = os.path.dirname(os.path.abspath(__file__))
for name, (steps, ncounts) in KEEP.items():
    case = build_case(name)
    data = {}
    counts = np.loadtxt(os.path.join(src, f"{name}_counts.txt"), dtype=np.int64).reshape(-1, 2)
    data["counts"] = counts[:ncounts, 1].astype(np.int32)
    invj = np.fromfile(os.path.join(src, f"{name}_invj.bin"), dtype=np.float64)
    data["invj_sha256"] = np.frombuffer(hashlib.sha256(invj.tobytes()).digest(), dtype=np.uint8)
    if invj.size <= 4 * 200:
        data["invj"] = invj.reshape(-1, 4)
    data["steps"] = np.asarray(steps, dtype=np.int32)
    for s in steps:
        d = read_dump(os.path.join(src, f"{name}_step{s:05d}.bin"))
        perm = canonical_order(d)
        compact = name in COMPACT
        keys = ("cell", "x", "y", "l0", "l1", "l2", "vx", "vy") if case.full_state else ("cell", "x", "y", "vx", "vy")
        if compact:  # big case: cells in full, positions as a digest, velocities every 16th particle
            keys = ("cell",)
            xy = b"".join(d[k][perm].tobytes() for k in ("x", "y"))
            data[f"s{s}_xy_sha256"] = np.frombuffer(hashlib.sha256(xy).digest(), dtype=np.uint8)
            data[f"s{s}_vx_16"] = d["vx"][perm][::16]
            data[f"s{s}_vy_16"] = d["vy"][perm][::16]
        for k in keys:
            data[f"s{s}_{k}"] = d[k][perm]
        if not case.full_state:  # local coordinates: keep a digest (they are bit-exact functions of the rest)
            lbytes = b"".join(d[k][perm].tobytes() for k in ("l0", "l1", "l2"))
            data[f"s{s}_l_sha256"] = np.frombuffer(hashlib.sha256(lbytes).digest(), dtype=np.uint8)
        if s > 0:
            data[f"s{s}_wx"] = d["wx"]
            data[f"s{s}_wy"] = d["wy"]
    for s in RAW_ORDER.get(name, ()):
        d = read_dump(os.path.join(src, f"{name}_step{s:05d}.bin"))
        for k in ("cell", "x", "y", "l0", "l1", "l2", "vx", "vy"):
            data[f"raw{s}_{k}"] = d[k]
    path = os.path.join(out, f"ref_{name}.npz")
    np.savez_compressed(path, **data)
    print(name, steps, os.path.getsize(path) // 1024, "KiB")
