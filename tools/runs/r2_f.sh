#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2f.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/b_r2f.log 2>&1
python tools/launch_table.py gpurun_out/launches_r2f.csv > gpurun_out/launches_r2f.txt 2>&1
tail -40 gpurun_out/launches_r2f.txt
