#!/bin/bash
# round 2, GPU call 23: token-mixing forward store pass without spills (two residual chunks in flight) - tests, kernel timing, full suite, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_call23.txt; : > $O
timeout 200 python -m pytest tests/test_tokenmix_gpu.py -q 2>&1 | tail -3 >> $O
timeout 120 python tools/tokenmix_bench.py --only fwd,fwd_ln,ln+fwd 2>&1 | grep -v "^{" >> $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 >> $O
timeout 300 python bench.py --no-eager-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python - >> $O <<'PY'
import json
d = None
for line in open("gpurun_out/r2i_bench.json"):
    if line.startswith("{"):
        d = json.loads(line)
tm = d["roofline"].get("second") or {}
print("bench:", round(d["ms_per_step"], 3), "ms", round(d["value"]), "samples/s; e2e", round(d["e2e"]["value"]), "; gemm frac", round(d["roofline"]["frac"], 3), "; split", d["sm_split"]["image_text_sms"])
PY
grep -o '"token_mix_fwd": {[^}]*}' gpurun_out/r2i_bench.json | head -1 >> $O
cat $O
