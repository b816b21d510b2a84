#!/bin/bash
# evidence at HEAD: the -m gpu suite, the small shipped cases, the ncu launch list of one step and a --set full capture of the three main kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
bash tools/small_cases.sh
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r03_step_kernels.csv python tools/profile_step.py > gpurun_out/r03_profile_step.log 2>&1
tail -1 gpurun_out/r03_profile_step.log
cp gpurun_out/step_particles.json gpurun_out/r03_step_particles.json
ncu --profile-from-start off --set full --import-source on --clock-control none -k 'regex:k_move_gather|k_rank|k_project_cells_lazy' \
    -o gpurun_out/r03_head python tools/profile_step.py > gpurun_out/r03_ncu_full.log 2>&1
tail -2 gpurun_out/r03_ncu_full.log
ls -la gpurun_out/*.ncu-rep
