#!/bin/bash
# usage: runexp.sh tag [ENV=VAL ...] ; runs bench (5 steps) and prints phases
tag=$1; shift
env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/exp_$tag.json"))
    print("$tag", round(j["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in j["roofline"]["phases"].items()}, "e2e", round(j["e2e"]["ms_per_step"],2))
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/exp_$tag.err").read()[-600:])
PY
