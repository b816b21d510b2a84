#!/bin/bash
# 2-GPU check of the migration protocols: parity (tests/test_gpu_multi.py) + A/B bench lines.  gpurun --gpus ${NG:-2} -- bash tools/run_mg2.sh
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/mg2_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/mg2_test.log 2>&1; echo "mg test rc=$?"; tail -3 gpurun_out/mg2_test.log
run() { # tag proto workload steps
  PFEM2_MG_PROTOCOL=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus ${NG:-2} --steps $4 --warmup 3 --workload $3 --no-cpu-baseline > gpurun_out/mg2_$1.json 2> gpurun_out/mg2_$1.err
  echo "$1 rc=$?"
  python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/mg2_$1.json") if l.startswith("{")][-1])
    print("$1", round(j["ms_per_step"],3), "ms/step", round(j["value"]/1e9,2), "G", j["config"].get("migration_protocol"), j["config"].get("migrated_particles_per_step"),
          {k:round(v["ms_per_step"],3) for k,v in j["roofline"]["phases"].items()})
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/mg2_$1.err").read()[-1500:])
PY
}
run c16_exact exact channel16m 20
run c16_nbr neighbour channel16m 20
run c16_nbr_b neighbour channel16m 20
run stress_nbr neighbour stress2m 20
run stress_exact exact stress2m 20
