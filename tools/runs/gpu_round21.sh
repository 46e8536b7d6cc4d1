#!/bin/bash
mkdir -p gpurun_out
for k in rank_count_vec4 topk_filter polar_quadrant; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/${k}_r1g -f python tools/k14_probe.py > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
  python tools/summarize_ncu.py gpurun_out/${k}_r1g.ncu-rep > gpurun_out/${k}_r1g.txt 2>&1; head -32 gpurun_out/${k}_r1g.txt | grep -i "kernel:\|dram__bytes\|time_duration\|dram_throughput\|wavefronts_mem_shared.sum.pct\|inst_executed.avg\|lts__t_sector_hit"
done
