#!/bin/bash
# Profiling session on the GPU box (one ncu family per call): launch list of ONE eager training step (time + DRAM
# bytes per launch) and `ncu --set full` captures of the fused token-mixing kernels and the three dominant GEMM kinds.
# Every command runs plain first; ncu only follows a clean exit (B200_PROFILING.md).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/profile_step.py --dump gpurun_out/gemm_table.txt > gpurun_out/profile_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
python tools/tokenmix_bench.py --tower image --iters 1 > gpurun_out/tm_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:token_mix -c 12 -o gpurun_out/tm_full \
    python tools/tokenmix_bench.py --tower image --iters 1 > gpurun_out/ncu_tm.log 2>&1
python tools/gemm_bench.py lin3 dw3 dz2 --iters 1 > gpurun_out/gemm_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 12 -o gpurun_out/gemm_full \
    python tools/gemm_bench.py lin3 dw3 dz2 --iters 1 > gpurun_out/ncu_gemm.log 2>&1
tail -n 2 gpurun_out/profile_step.log gpurun_out/ncu_tm.log gpurun_out/ncu_gemm.log
