#!/bin/bash
mkdir -p gpurun_out
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 500 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests -m gpu -q -x -k "large_gallery or train_step or properties or sharded or semantic or spectral_pair or spectral_rows or spectral_and_direct or topk_merge or pipeline" > gpurun_out/sanitize_rest.log 2>&1; echo "sanitizer rc=$?"
grep -c "Invalid\|misaligned" gpurun_out/sanitize_rest.log; grep "Invalid\|misaligned\|     at \|passed\|failed\|ERROR SUMMARY" gpurun_out/sanitize_rest.log | sort | uniq -c | sort -rn | head -20; tail -4 gpurun_out/sanitize_rest.log
