timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -3 > gpurun_out/run8.log
for z in 1 2; do
  echo "== MC_GEMM_ZDEPTH=$z" >> gpurun_out/run8.log
  MC_GEMM_ZDEPTH=$z python tools/gemm_bench.py dz2 lin4 txt_lin4 >> gpurun_out/run8.log 2>&1
  MC_GEMM_ZDEPTH=$z MC_GEMM_DEBUG_SKIP=3 python tools/gemm_bench.py dz2 lin4 >> gpurun_out/run8.log 2>&1
done
bash tools/env_sweep.sh "MC_GEMM_ZDEPTH=1 MC_SM_SPLIT=off" "MC_GEMM_ZDEPTH=2 MC_SM_SPLIT=off" >> gpurun_out/run8.log 2>&1
cat gpurun_out/run8.log
