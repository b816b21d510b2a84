#!/bin/bash
# round-3 A/B: groups of G consecutive tiles claimed through one global cursor (PFEM2_MOVE_GDYN = 4 / 8) against the fixed warp-strided share
V=$PWD/gpupfem2_b200/_variants
PFEM2_LIB_PATH=$V/libpfem2_gdyn4.so timeout 600 python -m pytest tests/test_gpu_lazy.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
export AB_STEPS=40
bash tools/ab_bench.sh st "PFEM2_X=0" gdyn4 "PFEM2_LIB_PATH=$V/libpfem2_gdyn4.so" gdyn8 "PFEM2_LIB_PATH=$V/libpfem2_gdyn8.so" gdyn4b "PFEM2_LIB_PATH=$V/libpfem2_gdyn4.so"
python - <<'PY'
import json
for t in ("st","gdyn4","gdyn8","gdyn4b"):
    try:
        j=json.loads([l for l in open(f"gpurun_out/ab_{t}.json") if l.startswith("{")][-1]); print(t, j["ms_per_step_list"])
    except Exception as e: print(t, "failed", e)
PY
