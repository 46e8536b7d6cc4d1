#!/bin/bash
mkdir -p gpurun_out
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 420 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not large_gallery and not train_step" > gpurun_out/sanitize_parity.log 2>&1; echo "sanitizer rc=$?"
grep -c "Invalid\|misaligned" gpurun_out/sanitize_parity.log; grep "Invalid\|misaligned\|     at \|passed\|failed\|ERROR SUMMARY" gpurun_out/sanitize_parity.log | sort | uniq -c | sort -rn | head -20
