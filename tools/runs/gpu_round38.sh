#!/bin/bash
mkdir -p gpurun_out
for kb in 72 56 36; do
echo "== preferred smem $kb KB"
WITW_RESIZE_SMEM_KB=$kb KB_ONLY=resize timeout 300 python tools/kernel_bench.py 2> gpurun_out/kernels_r1k.err | grep -i "resize_norm_kernel\|prepare" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  %-75s %.4f ms frac %.3f' % (d['kernel'][:75], d['ms'], d['frac']))"
done
