#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 400 python tools/ring_roof.py > gpurun_out/ring_r2p.log 2>&1; echo "rc=$?"; tail -22 gpurun_out/ring_r2p.log
