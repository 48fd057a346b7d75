#!/usr/bin/env python
"""Which GEMM-engine case breaks when the grid is sized for a share of the SMs?  Runs the B/32 channel-mix cases of
tests/test_gemm_gpu.py (and the shapes of a 64-sample step) one by one under the current MC_SM_LIMIT and prints pass/fail."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gemm_gpu as T  # noqa: E402

CASES = {
    "lin3_1600": lambda: T.run_case("tc", 50 * 32, 3072, 768, bias_mode=1, act=1, zout=True, c_bf16=True),
    "lin4_1600": lambda: T.run_case("tc", 50 * 32, 768, 3072, bias_mode=1, residual=True),
    "txt_lin3_2464": lambda: T.run_case("tc", 77 * 32, 2048, 512, bias_mode=1, act=1, zout=True, c_bf16=True),
    "wgrad_1600": lambda: T.run_case("tc", 3072, 768, 50 * 32, a_major=1, b_major=1, accumulate=True, split_k=0),
    "dz2_1600": lambda: T.run_case("tc", 50 * 32, 3072, 768, b_major=1, act=2, c_bf16=True),
    "dv_1600": lambda: T.run_case("tc", 50 * 32, 768, 3072, b_major=1),
    "lin3_3200": lambda: T.run_case("tc", 50 * 64, 3072, 768, bias_mode=1, act=1, zout=True, c_bf16=True),
    "lin4_3200": lambda: T.run_case("tc", 50 * 64, 768, 3072, bias_mode=1, residual=True),
    "dz2_3200": lambda: T.run_case("tc", 50 * 64, 3072, 768, b_major=1, act=2, c_bf16=True),
    "dv_3200": lambda: T.run_case("tc", 50 * 64, 768, 3072, b_major=1),
    "wgrad3_3200": lambda: T.run_case("tc", 3072, 768, 50 * 64, a_major=1, b_major=1, accumulate=True, split_k=0),
    "wgrad4_3200": lambda: T.run_case("tc", 768, 3072, 50 * 64, a_major=1, b_major=1, accumulate=True, split_k=0),
    "txt_lin3_4928": lambda: T.run_case("tc", 77 * 64, 2048, 512, bias_mode=1, act=1, zout=True, c_bf16=True),
    "txt_lin4_4928": lambda: T.run_case("tc", 77 * 64, 512, 2048, bias_mode=1, residual=True),
    "txt_dz2_4928": lambda: T.run_case("tc", 77 * 64, 2048, 512, b_major=1, act=2, c_bf16=True),
    "txt_dv_4928": lambda: T.run_case("tc", 77 * 64, 512, 2048, b_major=1),
    "txt_wgrad3_4928": lambda: T.run_case("tc", 2048, 512, 77 * 64, a_major=1, b_major=1, accumulate=True, split_k=0),
    "txt_wgrad4_4928": lambda: T.run_case("tc", 512, 2048, 77 * 64, a_major=1, b_major=1, accumulate=True, split_k=0),
}
only = sys.argv[1:]
for name, fn in CASES.items():
    if only and name not in only:
        continue
    sys.stderr.write(f"-- {name}\n")
    sys.stderr.flush()
    try:
        fn()
        print(f"limit={os.environ.get('MC_SM_LIMIT')} {name}: ok", flush=True)
    except AssertionError as e:
        print(f"limit={os.environ.get('MC_SM_LIMIT')} {name}: FAIL {str(e)[:160]}", flush=True)
