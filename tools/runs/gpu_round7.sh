#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/tests.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log | cut -c1-300; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench.log')); print('value',b['value'],'ms',b['ms_per_step'],'frac',b['roofline']['frac'],'kernel_ms',b['roofline']['kernel_ms'],'e2e',b['e2e']['value'],b['clocks'])
PY
tail -3 gpurun_out/bench.err
timeout 900 python tools/kernel_bench.py > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; echo "kernels rc=$?"; grep -E "topk|polar_quadrant" gpurun_out/kernels.jsonl; tail -3 gpurun_out/kernels.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
