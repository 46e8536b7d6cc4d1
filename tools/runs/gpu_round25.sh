#!/bin/bash
# round 1i: full GPU suite at HEAD, smoke, resize numbers, bench lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_r1i.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_r1i.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r1i.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r1i.log
KB_ONLY=resize timeout 300 python tools/kernel_bench.py 2> gpurun_out/kernels_r1i.err | grep -i "resize\|prepare" > gpurun_out/kernels_r1i_resize.jsonl; echo "kb rc=$?"; cut -c1-230 gpurun_out/kernels_r1i_resize.jsonl; tail -3 gpurun_out/kernels_r1i.err
timeout 600 python bench.py > gpurun_out/bench_r1i_n1.json 2> gpurun_out/bench_r1i_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r1i_n1.json; tail -3 gpurun_out/bench_r1i_n1.err
