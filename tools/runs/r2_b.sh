#!/bin/bash
# round 2: whole GPU suite, first bench line of the new path, hard-noise calibration
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_b_tests.log
tail -30 gpurun_out/r2_b_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_r2b.err
timeout 300 python tools/calibrate_hard.py > gpurun_out/hard_r2b.jsonl 2>&1
cat gpurun_out/hard_r2b.jsonl
