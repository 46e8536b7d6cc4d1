#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_resize.py -m gpu -x -q 2>&1 | tail -2
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 400 compute-sanitizer --tool memcheck --print-limit 10 python tools/resize_sanitize.py 2>&1 | tail -2
KB_ONLY=resize timeout 300 python tools/kernel_bench.py 2> gpurun_out/kernels_r1k.err | grep -i "resize\|prepare" > gpurun_out/kernels_r1k_resize.jsonl; cut -c1-200 gpurun_out/kernels_r1k_resize.jsonl
