#!/bin/bash
# round 2, GPU call 17: block-local row sums in ln_bwd_rows; kernel-level timing of the LayerNorm prologue; ncu of the fused forward
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_call17.txt; : > $O
timeout 300 python -m pytest tests/test_rowwise_gpu.py -q 2>&1 | tail -3 >> $O
timeout 120 python tools/rowwise_bench.py 2>&1 | grep -v "^{" >> $O
timeout 120 python tools/tokenmix_bench.py --only fwd,fwd_ln,ln+fwd 2>&1 | grep -v "^{" >> $O
timeout 120 python tools/tokenmix_bench.py --tower image --only fwd_ln --iters 4 > gpurun_out/plain_tm_ln.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:token_mix -s 4 -c 1 -f -o /tmp/r2_tm_ln python tools/tokenmix_bench.py --tower image --only fwd_ln --iters 4 > gpurun_out/ncu_tm_ln.log 2>&1 &&
{ python tools/ncu_multi_summary.py /tmp/r2_tm_ln.ncu-rep 1 > gpurun_out/r2_ncu_full_tokenmix_fwd_ln.txt 2>&1; rm -f /tmp/r2_tm_ln.ncu-rep; }
STEPS=20 bash tools/env_sweep.sh "MC_TM_FUSE_LN=1" "MC_TM_FUSE_LN=1" >> $O 2>&1
cat $O; head -30 gpurun_out/r2_ncu_full_tokenmix_fwd_ln.txt
