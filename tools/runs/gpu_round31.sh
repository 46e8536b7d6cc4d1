#!/bin/bash
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python tools/resize_sanitize.py > gpurun_out/racecheck_resize.log 2>&1; echo "racecheck rc=$?"
grep -c "hazard" gpurun_out/racecheck_resize.log; grep "hazard\|RACECHECK SUMMARY\|done" gpurun_out/racecheck_resize.log | sort | uniq -c | sort -rn | head; 
timeout 300 compute-sanitizer --tool synccheck python tools/resize_sanitize.py > gpurun_out/synccheck_resize.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/synccheck_resize.log
