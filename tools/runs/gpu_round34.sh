#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches90.csv python bench.py --fov 90 --steps 1 --warmup 3 --no-extras > gpurun_out/bench_under_ncu90.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_table.py gpurun_out/launches90.csv 16 > gpurun_out/launches_r1k_fov90.txt; head -12 gpurun_out/launches_r1k_fov90.txt | cut -c1-170; tail -17 gpurun_out/launches_r1k_fov90.txt | cut -c1-120
