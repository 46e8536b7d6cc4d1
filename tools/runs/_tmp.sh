#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for st in 10 20; do
WITW_BENCH_TRACE=1 python bench.py --steps $st --warmup 5 --no-extras 2> gpurun_out/trace_$st.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('steps $st', d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"
grep "^step" gpurun_out/trace_$st.err | tail -$((st+0)) | tr '\n' ' '; echo
done
