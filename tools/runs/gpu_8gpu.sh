#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"; cat gpurun_out/bench_n$N.log | cut -c1-400; tail -3 gpurun_out/bench_n$N.err
G=$((1000000 / N))
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 4 --warmup 3 --gallery-per-gpu $G > gpurun_out/bench_1m_n$N.log 2> gpurun_out/bench_1m_n$N.err; echo "bench 1M n$N rc=$?"; cat gpurun_out/bench_1m_n$N.log | cut -c1-400; tail -3 gpurun_out/bench_1m_n$N.err
