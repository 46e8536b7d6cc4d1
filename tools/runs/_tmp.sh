#!/bin/bash
cd "$GRAFT_REPO_ROOT"
nvidia-smi topo -m 2>&1 | head -30
lscpu | grep -i "numa\|socket\|^CPU(s)\|model name"
python - <<'PY'
import os, pynvml
pynvml.nvmlInit()
n = pynvml.nvmlDeviceGetCount()
print("devices", n, "affinity of this process", len(os.sched_getaffinity(0)))
for i in range(n):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    words = (os.cpu_count() + 63) // 64
    try:
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        print(i, "cpus", len(cpus), cpus[:4], "...", cpus[-4:])
    except Exception as e:
        print(i, "err", e)
PY
cat /sys/devices/system/node/node*/cpulist 2>/dev/null | head
