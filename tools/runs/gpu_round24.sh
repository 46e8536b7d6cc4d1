#!/bin/bash
mkdir -p gpurun_out
KB_ONLY=resize timeout 300 python tools/kernel_bench.py 2> gpurun_out/kernels_v2.err | grep -i "resize\|prepare" > gpurun_out/kernels_r1h_resize_v2.jsonl; cut -c1-200 gpurun_out/kernels_r1h_resize_v2.jsonl
for which in aerial pano; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resize_norm -s 2 -c 1 -o gpurun_out/resize_${which}_r1h -f python tools/resize_probe.py $which > gpurun_out/ncu_resize.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py gpurun_out/resize_${which}_r1h.ncu-rep > gpurun_out/resize_${which}_r1h.txt 2>&1; head -64 gpurun_out/resize_${which}_r1h.txt | grep -v "^LTS\|^TPC\|tensor"
ncu -i gpurun_out/resize_${which}_r1h.ncu-rep --page details --csv 2>/dev/null | grep -i "Stall\|Issue Slots Busy\|No Eligible\|Theoretical Occ\|Achieved Occ\|highest-utilized" | cut -d, -f12-16 | cut -c1-200 | head -30
done
