#!/bin/bash
# round 2 profiling call: `ncu --set full` of every GEMM kind of the step, the three fused token-mixing kernels and the
# contrastive head (FFMA and tensor-core slab path).  Every command runs plain first; ncu follows only a clean exit.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# the .ncu-rep files are summarised ON the box and deleted: gpurun only merges back <= 64 MiB
OUT=gpurun_out/r2_ncu_full_gemm.txt; : > $OUT
for shp in lin3 lin4 dz2 dw3 txt_lin3; do
  timeout 60 python tools/gemm_bench.py $shp --iters 4 > gpurun_out/plain_$shp.log 2>&1 &&
  timeout 300 $NCU -k regex:gemm_tc -s 4 -c 1 -f -o /tmp/r2_gemm_$shp python tools/gemm_bench.py $shp --iters 4 > gpurun_out/ncu_$shp.log 2>&1 &&
  { echo "=== tools/gemm_bench.py $shp (ncu --set full, 5th launch)" >> $OUT; python tools/ncu_summary.py /tmp/r2_gemm_$shp.ncu-rep 12 >> $OUT 2>&1; rm -f /tmp/r2_gemm_$shp.ncu-rep; }
done
timeout 60 python tools/tokenmix_bench.py --tower image --iters 4 > gpurun_out/plain_tm.log 2>&1 &&
timeout 400 $NCU -k regex:token_mix -c 12 -f -o /tmp/r2_tm_full python tools/tokenmix_bench.py --tower image --iters 1 > gpurun_out/ncu_tm.log 2>&1 &&
{ python tools/ncu_multi_summary.py /tmp/r2_tm_full.ncu-rep 4 > gpurun_out/r2_ncu_full_tokenmix.txt 2>&1; rm -f /tmp/r2_tm_full.ncu-rep; }
cat > /tmp/head_once.py <<'PY'
import os, sys, math, torch
sys.path.insert(0, os.getcwd())
from clip_mixer_b200 import ops
dev, E, n, N = "cuda:0", 512, 4096, 32768
g = torch.Generator().manual_seed(1)
ua = torch.nn.functional.normalize(torch.randn(N, E, generator=g), dim=1).to(dev)
ta = torch.nn.functional.normalize(torch.randn(N, E, generator=g), dim=1).to(dev)
t = torch.tensor([math.log(1 / 0.07)], device=dev)
ws = torch.empty(ops.head_workspace_bytes(n, N, E) // 4, device=dev)
loss, dls = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
dui, dut = torch.empty(n, E, device=dev), torch.empty(n, E, device=dev)
for _ in range(2):
    ops.head_fwd_bwd(ua[:n], ta[:n], ua, ta, t, n, N, E, 0, 1.0, loss, dui, dut, dls, ws)
torch.cuda.synchronize()
print("ok")
PY
for tc in 0 1; do
  MC_HEAD_TC=$tc timeout 120 python /tmp/head_once.py > gpurun_out/plain_head$tc.log 2>&1 &&
  MC_HEAD_TC=$tc timeout 400 ncu --set full --clock-control none -k regex:head_ -s 4 -c 4 -f -o /tmp/r2_head_tc$tc python /tmp/head_once.py > gpurun_out/ncu_head$tc.log 2>&1 &&
  { python tools/ncu_multi_summary.py /tmp/r2_head_tc$tc.ncu-rep > gpurun_out/r2_ncu_full_head_tc$tc.txt 2>&1; rm -f /tmp/r2_head_tc$tc.ncu-rep; }
done
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/r2_ncu_full_*
