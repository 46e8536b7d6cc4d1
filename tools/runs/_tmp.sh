#!/bin/bash
cd "$GRAFT_REPO_ROOT"
WITW_RING_MODES=ring_alone,full_half_b,ring_alone_half_b timeout 400 python tools/ring_roof.py > gpurun_out/ring_x.log 2>&1; python -c "
import json; d=json.load(open('gpurun_out/sweep_roof_v2.json')); print({k: round(v,3) for k,v in d['kernel_ms'].items()})"
