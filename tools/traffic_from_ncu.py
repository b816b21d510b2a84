"""ncu launch list of ONE particle step (tools/profile_step.py) -> profiles/rNN_traffic.json: DRAM bytes and time per kernel, per
bench phase and for the whole step, per particle.  bench.py scales these to a run's particle count for roofline.traffic."""
import csv
import json
import sys

PHASE_OF = {  # kernel name prefix -> bench phase
    "k_move_gather": "advect_locate", "k_move_tiles": "advect_locate", "k_pack_nodal": "advect_locate",
    "k_project_cells": "project_cells", "k_project_nodes": "project_nodes", "k_snapshot_dv": "correct", "k_correct": "correct",
}


def main(csv_path, particles_json, out_path):
    rows = [r for r in csv.reader(l for l in open(csv_path) if l.startswith('"'))]
    hdr = rows[0]
    col = {n: hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID")}
    per_launch = {}
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        d = per_launch.setdefault(r[col["ID"]], {"name": r[col["Kernel Name"]]})
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3,
                 "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(unit, 1.0)
        d[r[col["Metric Name"]]] = v * scale
    meta = json.load(open(particles_json))
    P = float(meta["particles"]) * meta.get("steps", 1)
    kernels, phases = {}, {}
    tot_b = tot_ms = 0.0
    for d in per_launch.values():
        short = d["name"].split("(")[0].split("<")[0].replace("void ", "").replace("pfem2::", "").strip()
        b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        ms = d.get("gpu__time_duration.sum", 0.0)
        k = kernels.setdefault(short, {"launches": 0, "dram_bytes": 0.0, "gpu_time_ms": 0.0})
        k["launches"] += 1
        k["dram_bytes"] += b
        k["gpu_time_ms"] += ms
        ph = next((p for pre, p in PHASE_OF.items() if short.startswith(pre)), "reorder")
        q = phases.setdefault(ph, {"dram_bytes": 0.0, "gpu_time_ms": 0.0})
        q["dram_bytes"] += b
        q["gpu_time_ms"] += ms
        tot_b += b
        tot_ms += ms
    for t in list(kernels.values()) + list(phases.values()):
        t["bytes_per_particle"] = t["dram_bytes"] / P
        t["GBps"] = t["dram_bytes"] / (t["gpu_time_ms"] * 1e-3) / 1e9 if t["gpu_time_ms"] else None
    out = {"source": "ncu --profile-from-start off --clock-control none, every kernel of one particle step (tools/profile_step.py); "
                     "dram__bytes_read.sum + dram__bytes_write.sum and gpu__time_duration.sum per launch (cold-cache, serialised)",
           "workload": meta.get("workload"), "particles_per_step": meta["particles"], "steps_profiled": meta.get("steps", 1),
           "kernels": kernels, "phases": phases,
           "step": {"dram_bytes": tot_b, "gpu_time_ms": tot_ms, "bytes_per_particle": tot_b / P, "launches": len(per_launch)}}
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps({k: round(v["bytes_per_particle"], 1) for k, v in phases.items()}), "step", round(tot_b / P, 1), "B/particle",
          round(tot_ms, 3), "ms")


if __name__ == "__main__":
    main(*sys.argv[1:4])
