#!/bin/bash
mkdir -p gpurun_out
for args in "--fov 90 --steps 10" "--fov 90 --steps 20" "--fov 360 --steps 10" "--fov 90 --steps 10 --warmup 10" "--fov 180 --steps 20"; do
  timeout 300 python bench.py $args --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$args', '| ms_per_step %.3f kernel_ms %.3f e2e %.3f warmup %d clocks %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['warmup'], d['clocks']))"
done
