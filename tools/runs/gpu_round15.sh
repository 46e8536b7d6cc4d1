#!/bin/bash
# spectral sweep after the control restructure: parity tests, then timing breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/spec_tests.log 2>&1; echo "spec tests rc=$?"; tail -30 gpurun_out/spec_tests.log | cut -c1-300
: > gpurun_out/spec_breakdown.jsonl
for dbg in 0 16 3 7 11 2 1; do
  WITW_SPEC_DEBUG=$dbg SPEC_BENCH_IMPLS=spectral timeout 300 python tools/spec_bench.py 360 >> gpurun_out/spec_breakdown.jsonl 2>> gpurun_out/spec_breakdown.err; echo "dbg=$dbg rc=$?"
done
python - <<'PY'
import json
for l in open('gpurun_out/spec_breakdown.jsonl'):
    d=json.loads(l); print('dbg',d['debug'],'kernel_ms %.2f count_only_ms %.2f topk_merge %.2f gprep %.3f qprep %.3f'%(d['sweep_kernel_ms'],d['sweep_count_only_ms'],d['sweep_topk16_merge_ms'],d['gallery_prep_ms'],d['query_prep_ms']))
PY
tail -3 gpurun_out/spec_breakdown.err
