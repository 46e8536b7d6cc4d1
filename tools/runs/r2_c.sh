#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python tools/var_bench2.py > gpurun_out/var2_r2c.jsonl 2>&1
cat gpurun_out/var2_r2c.jsonl
