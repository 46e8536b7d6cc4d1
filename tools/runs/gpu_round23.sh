#!/bin/bash
# resize kernel v2: parity + numbers
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_resize.py -m gpu -x -q > gpurun_out/pytest_resize_v2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_resize_v2.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r1h.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r1h.log
KB_ONLY=resize timeout 300 python tools/kernel_bench.py 2> gpurun_out/kernels_v2.err | grep -i "resize\|prepare" > gpurun_out/kernels_r1h_resize_v2.jsonl; echo "kb rc=$?"; cut -c1-330 gpurun_out/kernels_r1h_resize_v2.jsonl; tail -3 gpurun_out/kernels_v2.err
