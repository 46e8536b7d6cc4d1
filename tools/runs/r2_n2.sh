#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r2_n$N.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r2_n$N.json") if l.startswith("{")][-1])
print('value',d['value'],'ms',d['ms_per_step'],'kernel',d['roofline']['kernel_ms'],'e2e',d['e2e']['ms_per_step'])
for k in ('configs[3] 1M gallery sharded','exchange'): print(k, d.get(k))
PY
