#!/bin/bash
# bench as the first GPU process on a fresh box (what the driver does), then the configs[2] / configs[4] variations
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r1k_n1.json 2> gpurun_out/bench_r1k.err; echo "default rc=$?"; tail -3 gpurun_out/bench_r1k.err
timeout 600 python bench.py --fov 90 > gpurun_out/bench_r1k_n1_fov90.json 2> gpurun_out/bench_fov90.err; echo "fov90 rc=$?"; tail -3 gpurun_out/bench_fov90.err
timeout 600 python bench.py --fov 90 --gallery-per-gpu 100000 --steps 5 --no-extras > gpurun_out/bench_r1k_n1_fov90_100k.json 2> gpurun_out/bench_100k.err; echo "100k rc=$?"; tail -3 gpurun_out/bench_100k.err
python - <<'PY'
import json
for f in ("bench_r1k_n1", "bench_r1k_n1_fov90", "bench_r1k_n1_fov90_100k"):
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, "| value %.0f ms %.3f | e2e %.0f ms %.3f | kernel_ms %.3f | dense %s | cpu %.2f | clocks %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"],
        d.get("dense_sweep", {}).get("frac"), d["cpu_baseline"]["value"], d["clocks"]))
PY
