#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/tests.log | cut -c1-400
