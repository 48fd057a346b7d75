#!/bin/bash
# round 2, GPU call 4: where does the bf16 gradient error at 256 samples come from (tools/parity_probe.py matrix);
# loss traces of the no-sync graph test; token-mix wgrad CTA balance A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2_parity_probe.txt; : > $P
run() { echo "== $*" >> $P; env "$@" 2>&1 | grep -E "^\{|Error|error" | tail -2 >> $P; }
run X=1 timeout 300 python tools/parity_probe.py --batch 256 --tag default-graph
run X=1 timeout 300 python tools/parity_probe.py --batch 256 --eager --tag eager-single-stream
run X=1 timeout 300 python tools/parity_probe.py --batch 256 --reference --tag reference-bf16-autocast
run MC_TOKENMIX=gemm timeout 300 python tools/parity_probe.py --batch 256 --tag tokenmix-as-gemms
run MC_LIB=$PWD/clip-mixer_b200/libmixerclip_gg.so timeout 300 python tools/parity_probe.py --batch 256 --tag gelu-grad-ex2-rcp
run MC_TM_NO_AUG=1 timeout 300 python tools/parity_probe.py --batch 256 --tag no-aug
for b in 64 16; do
  run X=1 timeout 300 python tools/parity_probe.py --batch $b --tag default-graph
  run X=1 timeout 300 python tools/parity_probe.py --batch $b --reference --tag reference-bf16-autocast
done
timeout 300 python -m pytest tests/test_bench_path_gpu.py -q -s -k "no_host_sync" 2>&1 | grep -E "^\[|losses|passed|failed|assert" | tail -12 > gpurun_out/r2_nosync.txt
T=gpurun_out/r2_tokenmix_wgrad_balance.txt; : > $T
for cfg in "MC_TM_WGRAD_EQUAL=1" "MC_TM_WGRAD_EQUAL=0"; do
  echo "== $cfg" >> $T
  env $cfg timeout 120 python tools/tokenmix_bench.py --only wgrad 2>&1 | tail -2 >> $T
done
timeout 200 python -m pytest tests/test_tokenmix_gpu.py -q 2>&1 | tail -2 >> $T
cat $P | cut -c1-400; cat gpurun_out/r2_nosync.txt; cat $T
