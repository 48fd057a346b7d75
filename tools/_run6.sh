for k in 0 1 2 4 5 6 3; do
  echo "== MC_GEMM_DEBUG_SKIP=$k (1 no loads, 2 no MMAs, 4 no epilogue)" >> gpurun_out/skip_matrix.log
  MC_GEMM_DEBUG_SKIP=$k timeout 120 python tools/gemm_bench.py lin3 txt_lin3 dw3 dz2 lin4 >> gpurun_out/skip_matrix.log 2>&1
done
cat gpurun_out/skip_matrix.log
