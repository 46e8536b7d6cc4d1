#!/usr/bin/env python
"""Debug harness for the tcgen05 kernel: runs one small sweep in the mode given by the
environment (WITW_TC_CG=1|2, WITW_TC_FULL_B=0|1) and prints error statistics against a
float64 model of the same arithmetic.  Usage: python tools/tc_debug.py [fov] [G] [Q]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import witw_oracle as O
from witw_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libwitw_b200_hooks.so")   # the WITW_* switches exist only in the hooks build (make -C witw_b200/csrc HOOKS=1)
import witw_b200 as W

fov = int(sys.argv[1]) if len(sys.argv) > 1 else 360
G = int(sys.argv[2]) if len(sys.argv) > 2 else 40
Q = int(sys.argv[3]) if len(sys.argv) > 3 else 24
print("mode: CG=%s FULL_B=%s fov=%d G=%d Q=%d" % (os.environ.get("WITW_TC_CG", "2"), os.environ.get("WITW_TC_FULL_B", "0"), fov, G, Q), flush=True)
ov, su, sh = O.synth_features(G, Q, fov=fov, noise=1.0, seed=1)
ori, dist = W.match(ov.cuda(), su.cuda(), path="tc")
torch.cuda.synchronize()
ori, dist = ori.cpu(), dist.cpu()
corr = O.fused_fp64(ov.bfloat16().float(), su.bfloat16().float())[0]
ref_ori, ref = O.match(ov, su)
print("orientation agreement with fp32 reference: %.4f" % (ori == ref_ori).float().mean().item())
same = ori == ref_ori
if same.any():
    print("max |dist - ref| where orientation agrees: %.3e" % (dist - ref).abs()[same].max().item())
print("dist sample", dist[:3, :4].tolist())
print("ref  sample", ref[:3, :4].tolist())
print("ori  sample", ori[:3, :6].tolist(), "ref", ref_ori[:3, :6].tolist())
print("nan count", int(torch.isnan(dist).sum()))
