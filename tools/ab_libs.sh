#!/bin/bash
# A/B of kernel variants built side by side: tools/ab_libs.sh tag1=path1 tag2=path2 ...   (bench.py, 10 steps, phases printed)
mkdir -p gpurun_out
for kv in "$@"; do
  tag=${kv%%=*}; lib=${kv#*=}
  PFEM2_LIB_PATH=$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/ab_$tag.json") if l.startswith("{")][-1])
    print("$tag", round(j["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in j["roofline"]["phases"].items()}, "e2e", round(j["e2e"]["ms_per_step"],2))
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/ab_$tag.err").read()[-800:])
PY
done
