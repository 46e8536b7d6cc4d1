#!/bin/bash
# round 1h: full GPU suite at HEAD, smoke, bench line (with dense sweep + gallery sweep), resize kernel numbers
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_r1h.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_r1h.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r1h.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r1h.log
KB_ONLY=resize timeout 300 python tools/kernel_bench.py > gpurun_out/kernels_r1h.jsonl 2> gpurun_out/kernels_r1h.err; echo "kb rc=$?"; cut -c1-300 gpurun_out/kernels_r1h.jsonl; tail -3 gpurun_out/kernels_r1h.err
timeout 600 python bench.py > gpurun_out/bench_r1h_n1.json 2> gpurun_out/bench_r1h_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_r1h_n1.json; tail -3 gpurun_out/bench_r1h_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1h_reference.json 2>/dev/null; echo "ref rc=$?"; cat gpurun_out/bench_r1h_reference.json
