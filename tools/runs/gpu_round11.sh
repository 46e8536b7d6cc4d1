#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/tests.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench.log')); print('value',b['value'],'ms',b['ms_per_step'],'frac',b['roofline']['frac'],'kernel_ms',b['roofline']['kernel_ms'],'e2e',b['e2e']['value'],b['clocks'])
PY
tail -3 gpurun_out/bench.err
bash tools/gpu_launches.sh
python tools/launch_table.py gpurun_out/launches.csv 20 | tail -21
