#!/bin/bash
# full GPU suite + bench (spectral default, hankel for comparison) + launch list + one full ncu capture of match_spec_kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/tests.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench.log')); print('value',b['value'],'ms',b['ms_per_step'],'frac',b['roofline']['frac'],'kernel_ms',b['roofline']['kernel_ms'],'e2e',b['e2e']['value'],b['e2e']['ms_per_step'],b['clocks'],b['recall'])
PY
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --sweep hankel --steps 10 > gpurun_out/bench_hankel.log 2> gpurun_out/bench_hankel.err; echo "bench hankel rc=$?"; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_hankel.log')); print('hankel value',b['value'],'ms',b['ms_per_step'],'kernel_ms',b['roofline']['kernel_ms'],'e2e',b['e2e']['value'])
PY
bash tools/gpu_launches.sh > /dev/null
python tools/launch_table.py gpurun_out/launches.csv 20 | tail -45
timeout 900 ncu --set full --clock-control none --import-source on -k regex:match_spec_kernel -s 3 -c 1 -o gpurun_out/match_spec_r1g -f python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/summarize_ncu.py gpurun_out/match_spec_r1g.ncu-rep > gpurun_out/match_spec_r1g.txt 2>&1; head -40 gpurun_out/match_spec_r1g.txt
