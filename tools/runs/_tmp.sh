#!/bin/bash
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L
timeout 400 python -m pytest tests/test_gpu_peer.py -x -q -m gpu 2>&1 | tail -15
CUDA_VISIBLE_DEVICES=0 timeout 400 python -m pytest tests/test_gpu_peer.py -x -q -m gpu 2>&1 | tail -15
bash tools/runs/multi_gpu.sh 2 r2n
