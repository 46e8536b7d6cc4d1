#!/bin/bash
# gpurun --gpus N -- 'bash tools/runs/multi_gpu.sh N [tag]': the bench line at N GPUs (weak scaling, one 10k-item shard per GPU) with
# its side measurements (BASELINE configs[3]: the 1M-tile gallery sharded over the N GPUs; the exchange alone).
cd "$GRAFT_REPO_ROOT"
N=${1:-2}
TAG=${2:-run}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
echo "bench rc=$?"; tail -c 800 gpurun_out/bench_${TAG}_n$N.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_${TAG}_n$N.json") if l.startswith("{")][-1])
print('value',d['value'],'ms',d['ms_per_step'],'kernel',d['roofline']['kernel_ms'],'e2e',d['e2e']['ms_per_step'])
for k in ('configs[3] 1M gallery sharded','exchange'): print(k, d.get(k))
PY
