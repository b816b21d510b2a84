#!/bin/bash
# A/B of library builds / env switches on one box: tools/ab_bench.sh tag1 "ENV=.. ENV=.." tag2 "..." ...   (results: gpurun_out/ab_<tag>.json)
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  tag=$1; envs=$2; shift 2
  env $envs python bench.py --steps ${AB_STEPS:-10} --warmup 3 --no-cpu-baseline --no-extra ${AB_ARGS:-} > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err || tail -3 gpurun_out/ab_$tag.err
  python tools/show_bench.py gpurun_out/ab_$tag.json | head -2
done
