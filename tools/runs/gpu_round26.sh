#!/bin/bash
# round 1j: bench line with the pipelined D2H, launch list of the step
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r1j_n1.json 2> gpurun_out/bench_r1j_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r1j_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["dense_sweep"]["frac"])
PY
tail -3 gpurun_out/bench_r1j_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_table.py gpurun_out/launches.csv 20 > gpurun_out/launches_r1j.txt; head -24 gpurun_out/launches_r1j.txt | cut -c1-170
