#!/bin/bash
# First GPU call of round 2: the experiments prepared at the end of round 1 (DESIGN.md "Prepared for round 2"), one box,
# ~3 minutes.  Build the variant library HERE first (no GPU needed), it travels with the snapshot:
#   tools/round2_first.sh build        # nvcc -DMC_UNIFORM_WARP_IDX -> clip-mixer_b200/libmixerclip_uniform.so
#   gpurun --timeout 400 -- 'bash tools/round2_first.sh'
cd "$(dirname "$0")/.."
U=clip-mixer_b200/libmixerclip_uniform.so
if [ "$1" = "build" ]; then
  (cd clip-mixer_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
      -DMC_UNIFORM_WARP_IDX -o ../libmixerclip_uniform.so *.cu) && ls -la $U
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ring_handover tools/ubench/ring_handover.cu -lcuda
  exit $?
fi
mkdir -p gpurun_out
L=gpurun_out/round2_first.log
: > $L
echo "== ring hand-over probe incl. engine look-alikes" >> $L
RING_V2=1 timeout 60 ./tools/ubench/ring_handover >> $L 2>&1
SHAPES="lin3 lin4 dz2 dv dw3 txt_lin3 txt_lin4"
for cfg in "MC_GEMM_EXP=0" "MC_GEMM_EXP=1" "MC_GEMM_EXP=2" "MC_GEMM_EXP=3"; do
  echo "== product build, $cfg" >> $L
  env $cfg timeout 120 python tools/gemm_bench.py $SHAPES >> $L 2>&1
done
if [ -f $U ]; then
  echo "== -DMC_UNIFORM_WARP_IDX build: GEMM tests" >> $L
  MC_LIB=$PWD/$U timeout 200 python -m pytest tests/test_gemm_gpu.py tests/test_tokenmix_gpu.py -x -q 2>&1 | tail -3 >> $L
  for cfg in "MC_GEMM_EXP=0" "MC_GEMM_EXP=3"; do
    echo "== uniform build, $cfg" >> $L
    env MC_LIB=$PWD/$U $cfg timeout 120 python tools/gemm_bench.py $SHAPES >> $L 2>&1
    echo "== uniform build, $cfg, loads off / MMAs off (clocks per k-block: compare with the probe's 515 / 289)" >> $L
    env MC_LIB=$PWD/$U $cfg MC_GEMM_PROD2=0 MC_GEMM_DEBUG_SKIP=5 timeout 60 python tools/gemm_bench.py lin3 dw3 2>&1 | grep -v timeout >> $L
    env MC_LIB=$PWD/$U $cfg MC_GEMM_PROD2=0 MC_GEMM_DEBUG_SKIP=6 timeout 60 python tools/gemm_bench.py lin3 dw3 2>&1 | grep -v timeout >> $L
  done
  echo "== uniform build: model parity + bench" >> $L
  MC_LIB=$PWD/$U timeout 300 python -m pytest tests/test_model_parity_gpu.py tests/test_train_step_gpu.py -x -q 2>&1 | tail -3 >> $L
  MC_LIB=$PWD/$U MC_GEMM_EXP=3 bash tools/env_sweep.sh "MC_SM_SPLIT=auto" >> $L 2>&1
fi
bash tools/env_sweep.sh "MC_SM_SPLIT=auto" >> $L 2>&1
cat $L
