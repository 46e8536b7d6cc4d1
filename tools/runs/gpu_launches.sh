#!/bin/bash
# ncu launch list of the bench step (per-launch durations; cold-cache, serialised: compare shares).
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_table.py gpurun_out/launches.csv | head -60
