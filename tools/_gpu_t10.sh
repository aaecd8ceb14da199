timeout 300 python -m pytest tests/test_gpu_blocks.py -q -rf -k "ln_gemm" 2>&1 | tail -2
timeout 120 python tools/ab_kernels.py ln qkv ln_qkv
