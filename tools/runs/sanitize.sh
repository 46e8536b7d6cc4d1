#!/bin/bash
# gpurun -- 'bash tools/runs/sanitize.sh [tag]': compute-sanitizer over the small-shape GPU tests of the sweeps, the fp32 finish and
# the pipeline (memcheck on all of them; racecheck and synccheck on the sweep / finish tests).
cd "$GRAFT_REPO_ROOT"
TAG=${1:-run}
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_$TAG.txt
SMALL='match_vs_oracle or exact_finish_matches or list_overflow or duplicate or nan_and_zero or heatmap_sweep_one or gallery_builder or random_shapes or prepared_gallery'
echo "# compute-sanitizer, $(nvidia-smi --query-gpu=name --format=csv,noheader), $(date -u +%F)" > $OUT
for tool in memcheck racecheck synccheck; do
  sel="$SMALL"; files="tests/test_gpu_spec.py tests/test_gpu_tc.py"
  if [ $tool != memcheck ]; then sel='match_vs_oracle or exact_finish_matches or list_overflow'; fi
  echo "## $tool: pytest $files -k \"$sel\"" >> $OUT
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest $files -x -q -m gpu -k "$sel" > gpurun_out/san_$tool.log 2>&1
  echo "exit code $?" >> $OUT
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|error" gpurun_out/san_$tool.log | tail -12 >> $OUT
done
cat $OUT
