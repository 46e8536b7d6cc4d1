#!/bin/bash
mkdir -p gpurun_out
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 800 compute-sanitizer --tool memcheck --print-limit 20 python tools/resize_sanitize.py > gpurun_out/sanitize_resize.log 2>&1; echo "sanitizer rc=$?"; grep -c "Invalid\|misaligned" gpurun_out/sanitize_resize.log; tail -8 gpurun_out/sanitize_resize.log
