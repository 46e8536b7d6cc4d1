#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/r2_h_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_h_tests.log
tail -8 gpurun_out/r2_h_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2h.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/b_r2h.log 2>&1
python tools/launch_table.py gpurun_out/launches_r2h.csv 2>&1 | head -10
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err
echo "bench rc=$?"; tail -c 1000 gpurun_out/bench_r2h.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2h.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['hard']['ms_per_step'], d['e2e_resident']['ms_per_step'])"
