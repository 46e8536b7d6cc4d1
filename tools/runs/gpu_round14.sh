#!/bin/bash
# timing breakdown of the spectral sweep (WITW_SPEC_DEBUG bits, see csrc/match_spec.cu)
mkdir -p gpurun_out
: > gpurun_out/spec_breakdown.jsonl
for dbg in 0 16 3 19 7 11 15 2; do
  WITW_SPEC_CS=1 WITW_SPEC_DEBUG=$dbg SPEC_BENCH_IMPLS=spectral timeout 300 python tools/spec_bench.py 360 >> gpurun_out/spec_breakdown.jsonl 2>> gpurun_out/spec_breakdown.err; echo "dbg=$dbg rc=$?"
done
WITW_SPEC_CS=2 WITW_SPEC_DEBUG=16 SPEC_BENCH_IMPLS=spectral timeout 300 python tools/spec_bench.py 360 >> gpurun_out/spec_breakdown.jsonl 2>> gpurun_out/spec_breakdown.err
python - <<'PY'
import json
for l in open('gpurun_out/spec_breakdown.jsonl'):
    d=json.loads(l); print('cs',d['cs'],'dbg',d['debug'],'kernel_ms %.2f count_only_ms %.2f'%(d['sweep_kernel_ms'],d['sweep_count_only_ms']))
PY
tail -3 gpurun_out/spec_breakdown.err
