#!/bin/bash
mkdir -p gpurun_out
P=tools/probe/tma_probe
for args in "3 132 129 0 128" "3 132 129 124 1" "3 132 129 128 127" "3 132 129 4 3" "3 132 129 2 0"; do timeout 60 $P $args >> gpurun_out/tma_probe2.log 2>&1; done
cat gpurun_out/tma_probe2.log
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/tests.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 900 python tools/kernel_bench.py > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; echo "kernels rc=$?"; cat gpurun_out/kernels.jsonl; tail -3 gpurun_out/kernels.err
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:match_tc -s 3 -c 1 -o gpurun_out/match_tc_r1b python bench.py --steps 1 --warmup 3 > gpurun_out/ncu2.log 2>&1; echo "ncu tc rc=$?"
cat > /tmp/polar_prof.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, witw_b200 as W
x = torch.randn(512, 3, 256, 256, device='cuda')
for _ in range(3): y = W.polar_transform(x)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:polar_quadrant -s 2 -c 1 -o gpurun_out/polar_r1 python /tmp/polar_prof.py > gpurun_out/ncu3.log 2>&1; echo "ncu polar rc=$?"
