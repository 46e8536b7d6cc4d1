#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench.log')); print('value',b['value'],'ms',b['ms_per_step'],'frac',b['roofline']['frac'],'kernel_ms',b['roofline']['kernel_ms'],'e2e',b['e2e']['value'],b['e2e']['ms_per_step'],b['clocks'],b['cpu_baseline'])
PY
tail -3 gpurun_out/bench.err
timeout 900 python tools/kernel_bench.py > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; echo "kernel_bench rc=$?"; cut -c1-200 gpurun_out/kernels.jsonl | grep -v "exact\|full dist"
