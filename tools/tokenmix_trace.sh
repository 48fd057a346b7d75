#!/bin/bash
# Debug timeline of the fused token-mixing kernels.  Build the traced library first (here, no GPU needed):
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -DTM_TRACE \
#        -o clip-mixer_b200/libmixerclip_trace.so clip-mixer_b200/csrc/*.cu
# then run this on the GPU box: CTA 0 records a clock per pipeline event and role; events land in gpurun_out/tm_trace.log.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
MC_LIB=$PWD/clip-mixer_b200/libmixerclip_trace.so python tools/tokenmix_bench.py --iters 1 --tower ${1:-image} > gpurun_out/tm_trace.out 2> gpurun_out/tm_trace.log
tail -3 gpurun_out/tm_trace.out
