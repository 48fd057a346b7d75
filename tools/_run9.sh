rm -f gpurun_out/run9.log
for cfg in "MC_GEMM_ZPROMO=0 MC_GEMM_ZDEPTH=1" "MC_GEMM_ZPROMO=128 MC_GEMM_ZDEPTH=1" "MC_GEMM_ZPROMO=256 MC_GEMM_ZDEPTH=1" "MC_GEMM_ZPROMO=256 MC_GEMM_ZDEPTH=2" "MC_GEMM_RPROMO=256"; do
  echo "== $cfg" >> gpurun_out/run9.log
  env $cfg python tools/gemm_bench.py dz2 lin4 >> gpurun_out/run9.log 2>&1
  env $cfg MC_GEMM_DEBUG_SKIP=3 MC_GEMM_PROD2=0 python tools/gemm_bench.py dz2 lin4 2>&1 | grep -v timeout >> gpurun_out/run9.log
done
cat gpurun_out/run9.log
