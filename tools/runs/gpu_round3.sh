#!/bin/bash
mkdir -p gpurun_out
P=tools/probe/tma_probe
for args in "3 132 129 1 128" "3 132 64 1 128" "3 64 129 1 128" "3 64 64 0 0" "3 128 129 0 0" "3 132 129 127 1" "3 128 128 127 1" "3 132 120 1 1" "3 132 124 1 1" "2 132 129 1 128" "2 64 64 0 0" "3 256 64 0 0" "3 64 256 0 0" "3 128 127 1 1"; do
  timeout 60 $P $args >> gpurun_out/tma_probe.log 2>&1
done
cat gpurun_out/tma_probe.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -m gpu -q --no-header -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_polar_fast_kernel --deselect tests/test_gpu_parity.py::test_polar_fast_kernel_many_planes > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:match_tc -s 3 -c 1 -o gpurun_out/match_tc_r1 python bench.py --steps 1 --warmup 3 > gpurun_out/bench_under_ncu2.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
