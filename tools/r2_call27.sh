#!/bin/bash
# round 2, last profile: ncu launch list (time + DRAM bytes per launch) of one eager step with the shipped defaults
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
COMMIT=${1:-unknown}
timeout 60 python tools/profile_step.py > gpurun_out/profile_step.log 2>&1 &&
timeout 100 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/r2j_ncu_launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r2j_ncu_launches.csv gpurun_out/r2j_gemm_traffic.json $COMMIT > gpurun_out/r2j_launch_summary.txt 2>&1; head -20 gpurun_out/r2j_launch_summary.txt
