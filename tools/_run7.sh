timeout 600 python -m pytest tests/test_model_parity_gpu.py tests/test_gemm_gpu.py -x -q -s 2>&1 | grep -E "grads_worst|grads worst|passed|failed|Error|error" | tail -30 > gpurun_out/parity7.log
python tools/gemm_bench.py dz2 lin3 >> gpurun_out/parity7.log 2>&1
MC_GEMM_DEBUG_SKIP=3 python tools/gemm_bench.py dz2 >> gpurun_out/parity7.log 2>&1
cat gpurun_out/parity7.log
