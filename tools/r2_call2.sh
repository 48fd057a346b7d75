#!/bin/bash
# round 2, GPU call 2: full GPU suite (new parity tests), bench with the eager baseline, token-mix timeline, ncu launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2_gpu_tests1.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
for tw in image text; do
  MC_LIB=$PWD/clip-mixer_b200/libmixerclip_trace.so timeout 120 python tools/tokenmix_bench.py --iters 1 --tower $tw > gpurun_out/tm_trace_$tw.out 2> gpurun_out/tm_trace_$tw.log
done
timeout 120 python tools/tokenmix_bench.py > gpurun_out/r2_tokenmix_bench0.txt 2>&1
MC_TM_NO_AUG=1 timeout 120 python tools/tokenmix_bench.py > gpurun_out/r2_tokenmix_bench0_noaug.txt 2>&1
timeout 200 python tools/profile_step.py --dump gpurun_out/r2_gemm_table0.txt > gpurun_out/profile_step.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/r2_launches0.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
for cfg in "MC_GEMM_EPIBUF=1" "MC_GEMM_EPIBUF=2"; do
  echo "== $cfg" >> gpurun_out/r2_epibuf.txt
  env $cfg timeout 120 python tools/gemm_bench.py lin3 lin4 dz2 dv dw3 txt_lin3 txt_lin4 >> gpurun_out/r2_epibuf.txt 2>&1
done
for bn in 0 128 192 224; do
  echo "== text-tower shapes, MC_GEMM_BN=$bn (0 = cost model)" >> gpurun_out/r2_epibuf.txt
  env MC_GEMM_BN=$bn MC_GEMM_DEBUG=0 timeout 120 python tools/gemm_bench.py txt_lin3 txt_lin4 txt_dz2 txt_dv txt_dw3 txt_dw4 lin3 lin4 >> gpurun_out/r2_epibuf.txt 2>&1
done
env MC_GEMM_EPIBUF=2 timeout 200 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -2 >> gpurun_out/r2_epibuf.txt
STEPS=15 bash tools/env_sweep.sh "MC_GEMM_EPIBUF=1" "MC_GEMM_EPIBUF=2" "MC_PDL=1" "MC_PDL=1 MC_GEMM_EPIBUF=2" "MC_PDL=0" >> gpurun_out/r2_epibuf.txt 2>&1
MC_PDL=1 timeout 300 python -m pytest tests/test_train_step_gpu.py tests/test_model_parity_gpu.py -x -q 2>&1 | tail -2 >> gpurun_out/r2_epibuf.txt
tail -5 gpurun_out/r2_gpu_tests1.log; tail -c 600 gpurun_out/r2_bench1.json
