#!/bin/bash
# sweep tile width / split-K / cluster / epilogue kind for the main GEMM shapes (tools/gemm_bench.py)
for shape in lin3 lin4 dz2 dv dw3 dw4 txt_lin3 txt_lin4; do
  for cl in 1 2; do for bn in 128 192 256; do for sp in 1 2 4; do for te in 0 1; do
    case $shape in dw3|dw4) ;; *) [ $sp -gt 1 ] && continue;; esac
    r=$(MC_GEMM_CL=$cl MC_GEMM_BN=$bn MC_GEMM_SPLIT=$sp MC_GEMM_TMA_EPI=$te timeout -s KILL 60 python tools/gemm_bench.py $shape --iters 10 2>&1 | tail -1 | awk '{print $(NF-3), $(NF-1)}')
    echo "$shape cl=$cl bn=$bn split=$sp tma=$te : $r"
  done; done; done; done
done
