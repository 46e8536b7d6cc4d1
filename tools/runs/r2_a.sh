#!/bin/bash
# round 2, first GPU pass of the fp16 / deferral / finish rewrite: spectral + dense sweep tests
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2_a.log 2>&1
timeout 900 python -m pytest tests/test_gpu_spec.py -x -q -m gpu >> gpurun_out/r2_a.log 2>&1
echo "spec rc=$?" >> gpurun_out/r2_a.log
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -m gpu >> gpurun_out/r2_a.log 2>&1
echo "tc rc=$?" >> gpurun_out/r2_a.log
tail -60 gpurun_out/r2_a.log
