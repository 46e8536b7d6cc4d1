#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/tests.log | cut -c1-300
timeout 900 python tools/kernel_bench.py > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; echo "kernel_bench rc=$?"; cut -c1-230 gpurun_out/kernels.jsonl; tail -3 gpurun_out/kernels.err
