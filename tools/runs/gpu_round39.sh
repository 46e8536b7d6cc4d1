#!/bin/bash
# 8 GPUs: the bench line at N=8 (10k items per GPU) with the final code
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 2> gpurun_out/bench_r1k_n8.err | grep "^{" > gpurun_out/bench_r1k_n8.json; echo "bench n8 rc=$?"; tail -3 gpurun_out/bench_r1k_n8.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r1k_n8.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, {k: d["e2e"][k] for k in ("value", "ms_per_step", "h2d_alone_ms")}, d["roofline"]["kernel_ms"], d["clocks"])
PY
