#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python tools/spec_pair_ab.py > gpurun_out/pair_ab.log 2>&1; echo "ab rc=$?"; tail -8 gpurun_out/pair_ab.log
timeout 900 python -m pytest tests/test_gpu_spec.py -x -q -m gpu 2>&1 | tail -5
