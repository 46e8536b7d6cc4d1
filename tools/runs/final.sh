#!/bin/bash
# gpurun -- 'bash tools/runs/final.sh [tag]': everything the round's N = 1 evidence consists of, in one call.
cd "$GRAFT_REPO_ROOT"
TAG=${1:-r2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
bash tools/runs/tests_bench.sh $TAG
bash tools/runs/profile_step.sh $TAG
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>/dev/null; echo "reference rc=$?"
PYTORCH_NO_CUDA_MEMORY_CACHING=1 bash tools/runs/sanitize.sh $TAG > /dev/null 2>&1; tail -25 gpurun_out/sanitizer_$TAG.txt
