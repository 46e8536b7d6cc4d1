#!/bin/bash
mkdir -p gpurun_out
for pw in 32 16 8; do WITW_POLAR_PW=$pw timeout 120 python tools/polar_pw.py >> gpurun_out/polar_pw.log 2>&1; done; cat gpurun_out/polar_pw.log
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/tests.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 900 python tools/kernel_bench.py > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; echo "kernels rc=$?"; cat gpurun_out/kernels.jsonl; tail -3 gpurun_out/kernels.err
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:match_tc -s 3 -c 1 -o gpurun_out/match_tc_r1c python bench.py --steps 1 --warmup 3 > gpurun_out/ncu2.log 2>&1; echo "ncu tc rc=$?"
