#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for k in 1 2; do
WITW_BENCH_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/bench_r2e_$k.json 2> gpurun_out/bench_r2e_$k.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2e_$k.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
tail -25 gpurun_out/bench_r2e_$k.err
done
