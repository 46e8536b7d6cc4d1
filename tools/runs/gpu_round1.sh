#!/bin/bash
# First GPU pass: parity tests of the simple kernels, then the tcgen05 kernel in its four debug modes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/parity.log 2>&1; echo "parity rc=$?" | tee -a gpurun_out/parity.log
for cg in 1 2; do for fb in 1 0; do
  WITW_TC_CG=$cg WITW_TC_FULL_B=$fb timeout 120 python tools/tc_debug.py 360 40 24 > gpurun_out/tc_cg${cg}_fb${fb}.log 2>&1; echo "tc cg=$cg fb=$fb rc=$?" | tee -a gpurun_out/tc_summary.log
done; done
WITW_TC_CG=2 timeout 120 python tools/tc_debug.py 90 40 24 > gpurun_out/tc_90.log 2>&1; echo "tc90 rc=$?" | tee -a gpurun_out/tc_summary.log
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/tc_tests.log 2>&1; echo "tc tests rc=$?" | tee -a gpurun_out/tc_summary.log
tail -5 gpurun_out/parity.log; cat gpurun_out/tc_summary.log; tail -15 gpurun_out/tc_cg2_fb0.log
