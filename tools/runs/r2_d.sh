#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python tools/var_bench2.py libwitw_b200.so libwitw_b200_hooks.so > gpurun_out/var2_r2d.jsonl 2>&1
cat gpurun_out/var2_r2d.jsonl
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_d_tests.log
tail -15 gpurun_out/r2_d_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
echo "bench rc=$?"; tail -c 2000 gpurun_out/bench_r2d.err
python tools/ring_roof.py > gpurun_out/ring_r2d.log 2>&1; tail -12 gpurun_out/ring_r2d.log
