#!/bin/bash
# isolated particle step on the shipped meshes (BASELINE configs[0], configs[1]): graphs on (default) / off
mkdir -p gpurun_out
for w in poiseuille cylinder; do
  for g in 0 -1; do
    PFEM2_GRAPH_ADVECT=$g python bench.py --workload $w --steps 200 --warmup 10 --no-cpu-baseline --no-extra > gpurun_out/small_${w}_g$g.json 2> gpurun_out/small_${w}_g$g.err || tail -3 gpurun_out/small_${w}_g$g.err
    python tools/show_bench.py gpurun_out/small_${w}_g$g.json | head -1
  done
done
