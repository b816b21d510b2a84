#!/bin/bash
# A/B of how the gathered move pass hands out its tiles (profiles/r03_summary.md §2): the default (groups of 4 tiles through one global
# cursor) against the fixed warp-strided share of round 2 and other group sizes, 40 steps so that the late-step stragglers show up.
#   make -C gpupfem2_b200/csrc variant NAME=static DEFS=-DPFEM2_MOVE_GDYN=0
#   make -C gpupfem2_b200/csrc variant NAME=gdyn8 DEFS=-DPFEM2_MOVE_GDYN=8
#   gpurun -- bash tools/ab_move_claims.sh
V=$PWD/gpupfem2_b200/_variants
export AB_STEPS=40
bash tools/ab_bench.sh default "PFEM2_X=0" static "PFEM2_LIB_PATH=$V/libpfem2_static.so" gdyn8 "PFEM2_LIB_PATH=$V/libpfem2_gdyn8.so"
python - <<'PY'
import json
for t in ("default", "static", "gdyn8"):
    try:
        j = json.loads([l for l in open(f"gpurun_out/ab_{t}.json") if l.startswith("{")][-1])
        print(t, j["ms_per_step_list"])
    except Exception as e:
        print(t, "failed", e)
PY
