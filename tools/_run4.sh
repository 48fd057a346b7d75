timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -3 > gpurun_out/sweep4.log
for p in 0 1; do
  echo "== MC_GEMM_PROD2=$p" >> gpurun_out/sweep4.log
  MC_GEMM_PROD2=$p python tools/gemm_bench.py lin3 lin4 dz2 dv dw3 dw4 txt_lin3 txt_lin4 >> gpurun_out/sweep4.log 2>&1
done
bash tools/env_sweep.sh "MC_GEMM_PROD2=0 MC_SM_SPLIT=off" "MC_GEMM_PROD2=1 MC_SM_SPLIT=off" "MC_GEMM_PROD2=1" >> gpurun_out/sweep4.log 2>&1
cat gpurun_out/sweep4.log
