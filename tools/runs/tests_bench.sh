#!/bin/bash
# gpurun -- 'bash tools/runs/tests_bench.sh [tag]': the whole GPU suite, one bench line with its side measurements, and the ncu
# launch list of the step (per-launch times are cold-cache and serialised: shares, not absolutes).  Everything lands in gpurun_out/.
cd "$GRAFT_REPO_ROOT"
TAG=${1:-run}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests_$TAG.log 2>&1
echo "tests rc=$?" >> gpurun_out/tests_$TAG.log
tail -5 gpurun_out/tests_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; tail -c 1000 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-extras > /dev/null 2>&1
python tools/launch_table.py gpurun_out/launches_$TAG.csv 24 > gpurun_out/launches_$TAG.txt 2>&1
head -12 gpurun_out/launches_$TAG.txt
