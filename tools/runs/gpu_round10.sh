#!/bin/bash
# Re-entry verification: full GPU test suite, smoke, default bench, reference arm.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/tests.log | cut -c1-400
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench.log
tail -3 gpurun_out/bench.err
timeout 600 python tools/kernel_bench.py > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; echo "kernel_bench rc=$?"; cut -c1-300 gpurun_out/kernels.jsonl
