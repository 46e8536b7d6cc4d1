#!/bin/bash
# bring-up of the spectral sweep: parity tests first, then timing per cluster size
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
WITW_SPEC_CS=1 timeout 900 python -m pytest tests/test_gpu_spec.py -m gpu -q --no-header -p no:cacheprovider -k "not cluster_sizes" > gpurun_out/spec_tests_cs1.log 2>&1; echo "spec tests cs1 rc=$?"; tail -40 gpurun_out/spec_tests_cs1.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_spec.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/spec_tests.log 2>&1; echo "spec tests rc=$?"; tail -40 gpurun_out/spec_tests.log | cut -c1-300
: > gpurun_out/spec_bench.jsonl
for cs in 1 2 4 8; do
  WITW_SPEC_CS=$cs SPEC_BENCH_IMPLS=spectral timeout 300 python tools/spec_bench.py 360 >> gpurun_out/spec_bench.jsonl 2>> gpurun_out/spec_bench.err; echo "cs=$cs rc=$?"
done
WITW_SPEC_CS=1 WITW_SPEC_SKIP_IFFT=1 SPEC_BENCH_IMPLS=spectral timeout 300 python tools/spec_bench.py 360 >> gpurun_out/spec_bench.jsonl 2>> gpurun_out/spec_bench.err
WITW_SPEC_CS=4 WITW_SPEC_SKIP_IFFT=1 SPEC_BENCH_IMPLS=spectral timeout 300 python tools/spec_bench.py 360 >> gpurun_out/spec_bench.jsonl 2>> gpurun_out/spec_bench.err
SPEC_BENCH_IMPLS=hankel timeout 300 python tools/spec_bench.py 360 >> gpurun_out/spec_bench.jsonl 2>> gpurun_out/spec_bench.err
cut -c1-420 gpurun_out/spec_bench.jsonl; tail -5 gpurun_out/spec_bench.err
