#!/bin/bash
# round 2, GPU call 19: -DTM_TRACE timeline of the token-mixing forward kernel, bf16-operand and LayerNorm-prologue variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tw in image text; do
MC_LIB=$PWD/clip-mixer_b200/libmixerclip_trace.so timeout 200 python tools/tokenmix_bench.py --iters 1 --tower $tw --only fwd,fwd_ln > gpurun_out/tm_trace.out 2> /tmp/tm_trace.log
{ echo "== $tw tower: fwd (bf16 operand in)"; python tools/tm_trace_summary.py /tmp/tm_trace.log 3; echo "== $tw tower: fwd_ln (LayerNorm in the prologue)"; python tools/tm_trace_summary.py /tmp/tm_trace.log 7; } > gpurun_out/r2_tm_trace_$tw.txt 2>&1
done
cat gpurun_out/r2_tm_trace_image.txt
