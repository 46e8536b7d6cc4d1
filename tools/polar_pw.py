#!/usr/bin/env python
"""Times the staged polar kernel for the warp-patch width in WITW_POLAR_PW (8, 16 or 32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from witw_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libwitw_b200_hooks.so")   # the WITW_* switches exist only in the hooks build (make -C witw_b200/csrc HOOKS=1)
import witw_b200 as W
x = torch.randn(1024, 3, 256, 256, device="cuda")
ref = W.polar_transform(x[:4], exact=True)
out = W.polar_transform(x)
err = (out[:4] - ref).abs().max().item()
for _ in range(3): W.polar_transform(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): W.polar_transform(x)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print("PW=%s  %.4f ms  %.1f GB/s  max err vs exact %.2e" % (os.environ.get("WITW_POLAR_PW", "8"), ms, 4.0 * 1024 * 3 * (65536 * 2) / ms / 1e6, err))
