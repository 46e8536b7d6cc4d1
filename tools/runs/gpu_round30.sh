#!/bin/bash
mkdir -p gpurun_out
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 420 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_spec.py tests/test_gpu_tc.py -m gpu -q -x -k "match_vs_oracle or nan_and_zero or gallery_builder or heatmap or evaluate_ranks_vs_oracle or exact_finish" > gpurun_out/sanitize_tc.log 2>&1; echo "sanitizer rc=$?"
grep -c "Invalid\|misaligned" gpurun_out/sanitize_tc.log; grep "Invalid\|misaligned\|     at \|passed\|failed\|ERROR SUMMARY" gpurun_out/sanitize_tc.log | sort | uniq -c | sort -rn | head -20; tail -5 gpurun_out/sanitize_tc.log
