#!/bin/bash
# bench variations: configs[2] (90 deg, 10k x 10k), configs[4] sweep half (90 deg, 100k gallery), default line
mkdir -p gpurun_out
timeout 600 python bench.py --fov 90 --steps 10 > gpurun_out/bench_r1k_n1_fov90.json 2> gpurun_out/bench_fov90.err; echo "fov90 rc=$?"; tail -3 gpurun_out/bench_fov90.err
timeout 600 python bench.py --fov 90 --gallery-per-gpu 100000 --steps 5 --no-extras > gpurun_out/bench_r1k_n1_fov90_100k.json 2> gpurun_out/bench_100k.err; echo "100k rc=$?"; tail -3 gpurun_out/bench_100k.err
timeout 600 python bench.py > gpurun_out/bench_r1k_n1.json 2> gpurun_out/bench_r1k.err; echo "default rc=$?"; tail -3 gpurun_out/bench_r1k.err
python - <<'PY'
import json
for f in ("bench_r1k_n1_fov90", "bench_r1k_n1_fov90_100k", "bench_r1k_n1"):
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, d["config"]["workload"][:80], "| value %.0f ms %.3f | e2e %.0f ms %.3f h2d_alone %.2f | kernel_ms %.3f frac %.2f | dense %s | cpu %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_alone_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"],
        d.get("dense_sweep", {}).get("frac"), d["cpu_baseline"]["value"]))
PY
