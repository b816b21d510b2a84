#!/bin/bash
# N-GPU check of the three migration transports: parity + spill (tests/test_gpu_multi.py), per-call breakdown, bench lines.
#   gpurun --gpus N -- 'NG=N bash tools/run_mg_p2p.sh'
NG=${NG:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/mgp_test.log 2>&1; echo "mg test rc=$?"; grep -E "MG_|passed|failed|Error|assert" gpurun_out/mgp_test.log | tail -12
for p in p2p neighbour; do
  PFEM2_MG_PROTOCOL=$p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29581 tools/diag_mg.py 2>&1 | grep -E "^world|untimed|Error|error|Traceback" | tail -5
done
run() { # tag proto workload steps
  PFEM2_MG_PROTOCOL=$2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus $NG --steps $4 --warmup 3 --workload $3 --no-cpu-baseline > gpurun_out/mgp_$1.json 2> gpurun_out/mgp_$1.err
  echo "$1 rc=$?"
  python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/mgp_$1.json") if l.startswith("{")][-1])
    print("$1", round(j["ms_per_step"],3), "ms/step", round(j["value"]/1e9,2), "G", j["config"].get("migration_protocol"), j["config"].get("migrated_particles_per_step"),
          {k:round(v["ms_per_step"],3) for k,v in j["roofline"]["phases"].items()})
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/mgp_$1.err").read()[-1500:])
PY
}
run c16_p2p p2p channel16m 20
run c16_nbr neighbour channel16m 20
run c16_p2p_b p2p channel16m 20
run stress_p2p p2p stress2m 20
