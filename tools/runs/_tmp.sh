#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python tools/spec_pair_ab.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_spec.py -x -q -m gpu 2>&1 | tail -3
WITW_RING_MODES=ring_alone,epilogue_alone,ring timeout 400 python tools/ring_roof.py > gpurun_out/ring_x.log 2>&1; python -c "
import json; d=json.load(open('gpurun_out/sweep_roof_v2.json')); print({k: round(v,3) for k,v in d['kernel_ms'].items()})"
