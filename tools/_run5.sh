for s in lin3 dz2 txt_lin3 dw3; do
  python tools/gemm_bench.py $s --iters 1 > gpurun_out/gemm_plain_$s.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 3 --launch-count 1 -f -o gpurun_out/r1s3_gemm_$s \
      python tools/gemm_bench.py $s --iters 1 > gpurun_out/ncu_gemm_$s.log 2>&1
  tail -1 gpurun_out/ncu_gemm_$s.log
done
