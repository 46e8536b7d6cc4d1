#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider -k "polar or bilinear" > gpurun_out/tests_polar.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/tests_polar.log | cut -c1-300
KB_ONLY=polar timeout 600 python tools/kernel_bench.py > gpurun_out/kernels_polar.jsonl 2> gpurun_out/kernels_polar.err; echo "kernel_bench rc=$?"; cut -c1-230 gpurun_out/kernels_polar.jsonl; tail -3 gpurun_out/kernels_polar.err
for pw in 8 32; do WITW_POLAR_PW=$pw KB_ONLY=polar timeout 600 python tools/kernel_bench.py 2>/dev/null | grep "polar_quadrant" | cut -c1-200 | sed "s/^/pw=$pw /"; done
