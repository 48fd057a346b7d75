#!/bin/bash
# round 2, GPU call 8: after the epilogue-slot fix - GEMM probe under SM limits, whole GPU suite, parity probes, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_gemm_limit_probe2.txt; : > $O
for lim in 48 80 84 64 148; do
  MC_SM_LIMIT=$lim timeout 300 python tools/gemm_limit_probe.py 2>/dev/null >> $O
done
echo "fails: $(grep -c FAIL $O) ok: $(grep -c ': ok' $O)"; grep FAIL $O | head
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | grep -E "^\[|passed|failed|FAILED|Error|EXEMPT|skipped|losses" | tail -60 > gpurun_out/r2_gpu_tests3.log
tail -25 gpurun_out/r2_gpu_tests3.log | cut -c1-400
P=gpurun_out/r2_parity_probe3.txt; : > $P
run() { echo "== $*" >> $P; env "$@" 2>&1 | grep -E "^\{|Error|error" | cut -c1-330 >> $P; }
run X=1 timeout 300 python tools/parity_probe.py --batch 256 --tag graph-two-streams-auto --repeat 3
run MC_SM_SPLIT=84,64 timeout 300 python tools/parity_probe.py --batch 64 --no-graph --tag eager-two-streams-84-64 --repeat 2
grep -o '"sm_split": [^,]*, [0-9]*\]\|"median": [0-9.e-]*\|"whole_model": [0-9.e-]*' $P | paste - - - | head
STEPS=20 bash tools/env_sweep.sh "MC_SM_SPLIT=auto" "MC_SM_SPLIT=off" "MC_SM_SPLIT=auto" > gpurun_out/r2_bench_after_fix.txt 2>&1
cat gpurun_out/r2_bench_after_fix.txt
