#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_g_tests.log
tail -15 gpurun_out/r2_g_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err
echo "bench rc=$?"; tail -c 1000 gpurun_out/bench_r2g.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2g.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2g.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/b_r2g.log 2>&1
python tools/launch_table.py gpurun_out/launches_r2g.csv 2>&1 | head -12
