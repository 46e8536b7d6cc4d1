#!/bin/bash
# 2 GPUs: bench line at N=2 with the pipelined result download; gloo-free check that both ranks agree
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r1j_n2.json 2> gpurun_out/bench_r1j_n2.err; echo "bench n2 rc=$?"; cut -c1-300 gpurun_out/bench_r1j_n2.json; tail -5 gpurun_out/bench_r1j_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r1j_n2.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"], d["recall"])
PY
