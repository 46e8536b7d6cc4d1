#!/bin/bash
# gpurun -- 'bash tools/runs/profile_step.sh [tag]': one ncu --set full capture of every kernel of the bench step (one GPU, never
# a multi-rank command), summarised by tools/summarize_ncu.py; then the ring roof of the spectral sweep (tools/ring_roof.py).
cd "$GRAFT_REPO_ROOT"
TAG=${1:-run}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:match_spec_kernel|finish_pairs_kernel|spec_gallery_prep_kernel|spec_query_prep_kernel|finish_topk_kernel' -s 15 -c 5 -f -o gpurun_out/step_$TAG python bench.py --steps 2 --warmup 3 --no-extras > /dev/null 2>&1
python tools/summarize_ncu.py gpurun_out/step_$TAG.ncu-rep > gpurun_out/step_$TAG.txt 2>&1
python tools/summarize_ncu.py gpurun_out/step_$TAG.ncu-rep match_spec_kernel > gpurun_out/match_spec_$TAG.txt 2>&1
python tools/ring_roof.py > gpurun_out/ring_$TAG.log 2>&1
tail -20 gpurun_out/ring_$TAG.log
