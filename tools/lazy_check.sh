#!/bin/bash
# Hardware bring-up of the lazy re-sort (pfem2_options.lazy_sort): the first gpurun call of round 2.
#   make -C gpupfem2_b200/csrc all variants && gpurun --timeout 1800 -- 'bash tools/lazy_check.sh'
# 1. parity tests of the path in both tile layouts (each in its own process under `timeout`: a faulting kernel poisons the
#    context and must not hang the box);  2. A/B bench default vs lazy (64-byte swizzled tiles, then linear tiles) on channel16m.
# Everything lands in gpurun_out/lazy_*.
mkdir -p gpurun_out
# 0. the open hardware question in isolation: gather4 through 64-byte-swizzled maps -> tile store ("wrong pieces 0" = the rows land
#    where the 32-row tile expects them); linear maps as the control
for m in 2 3; do for sw in 0 1; do timeout 120 tools/micro/_bin/tma_g4s4 $m 1 4 24 $sw | tee -a gpurun_out/lazy_gather4_swizzle.log; done; done
export PFEM2_TEST_LAZY=1
for k in "refuses" "reference_dumps and swizzle64" "reference_dumps and linear" "oracle and swizzle64" "oracle and linear" \
         "cylinder or clamped or growth" "eager or step_host_and_upload" "pipelined"; do
  tag=$(echo "$k" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_gpu_lazy.py -x -q -k "$k" > gpurun_out/lazy_test_$tag.log 2>&1
  echo "== $k: rc=$? $(tail -1 gpurun_out/lazy_test_$tag.log)"
done
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "lazy_sort or checksum" > gpurun_out/lazy_test_fullsize.log 2>&1
echo "== full-size properties (16M and 256M particles) with lazy_sort: rc=$? $(tail -1 gpurun_out/lazy_test_fullsize.log)"
run() { # tag ENV=VAL ...
  tag=$1; shift
  timeout 900 env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/lazy_bench_$tag.json 2> gpurun_out/lazy_bench_$tag.err
  python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/lazy_bench_$tag.json") if l.startswith("{")][-1])
    print("$tag", round(j["ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in j["roofline"]["phases"].items()},
          "e2e", round(j["e2e"]["ms_per_step"], 2))
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/lazy_bench_$tag.err").read()[-800:])
PY
}
run default PFEM2_LAZY_SORT=0
# A/B of the spill-free FAST move pass against the form of rounds 1a-1d (build the variant first, here: make -C gpupfem2_b200/csrc variants)
if [ -f gpupfem2_b200/_variants/libpfem2_nofast.so ]; then run default_nofast PFEM2_LAZY_SORT=0 PFEM2_LIB_PATH=$PWD/gpupfem2_b200/_variants/libpfem2_nofast.so; fi
run lazy_swizzle64 PFEM2_LAZY_SORT=1 PFEM2_LAZY_SWIZZLE=1
run lazy_linear PFEM2_LAZY_SORT=1 PFEM2_LAZY_SWIZZLE=0
# 3. only if the lazy bench ran: launch list of one lazy step (shares) and one --set full capture of its three new kernels
if grep -q '"value"' gpurun_out/lazy_bench_lazy_swizzle64.json 2>/dev/null; then
  PFEM2_LAZY_SORT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 60 --csv \
    --log-file gpurun_out/lazy_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  PFEM2_LAZY_SORT=1 timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'k_advect_locate_lazy|k_rank|k_project_cells_lazy' -s 9 -c 3 -o gpurun_out/lazy_prof \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/lazy_prof.log 2>&1
  ls -la gpurun_out/lazy_prof.ncu-rep gpurun_out/lazy_launches.csv
fi
