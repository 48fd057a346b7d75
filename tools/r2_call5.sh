#!/bin/bash
# round 2, GPU call 5: bisect the graph / two-stream discrepancy
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_bisect.txt; : > $O
for m in eager1 eager2 graph1 graph2; do
  for s in "--sync" ""; do
    timeout 120 python tools/graph_vs_eager.py $m $s 2>&1 | grep "^{" >> $O
  done
done
MC_SM_SPLIT=off timeout 120 python tools/graph_vs_eager.py graph2 2>&1 | grep "^{" >> $O
MC_SM_SPLIT=off timeout 120 python tools/graph_vs_eager.py eager2 2>&1 | grep "^{" >> $O
timeout 120 python tools/graph_vs_eager.py graph2 --precision bf16 2>&1 | grep "^{" >> $O
timeout 120 python tools/graph_vs_eager.py eager1 --precision bf16 --sync 2>&1 | grep "^{" >> $O
P=gpurun_out/r2_parity_probe2.txt; : > $P
run() { echo "== $*" >> $P; env "$@" 2>&1 | grep -E "^\{|Error|error" | cut -c1-330 >> $P; }
run X=1 timeout 300 python tools/parity_probe.py --batch 64 --eager --tag eager-single-stream
run X=1 timeout 300 python tools/parity_probe.py --batch 64 --no-graph --tag eager-two-streams --repeat 3
run MC_SM_SPLIT=off timeout 300 python tools/parity_probe.py --batch 64 --no-graph --tag eager-two-streams-nosplit --repeat 3
run MC_SM_SPLIT=84,64 timeout 300 python tools/parity_probe.py --batch 64 --no-graph --tag eager-two-streams-84-64 --repeat 3
run X=1 timeout 300 python tools/parity_probe.py --batch 64 --single-stream --tag graph-single-stream --repeat 3
run MC_SM_SPLIT=off timeout 300 python tools/parity_probe.py --batch 64 --tag graph-two-streams-nosplit --repeat 3
run X=1 timeout 300 python tools/parity_probe.py --batch 64 --tag graph-two-streams-auto --repeat 3
cat $O | cut -c1-420; cat $P | cut -c1-250
