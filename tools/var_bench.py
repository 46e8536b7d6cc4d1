#!/usr/bin/env python
"""Times the sweep kernel of library variants (witw_b200/libwitw_var_*.so, built by hand for an experiment) at 10k x 10k.
One subprocess per variant: python tools/var_bench.py [names...]"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(path):
    sys.path.insert(0, ROOT)
    import torch
    from witw_b200 import _lib
    _lib.LIB_PATH = path
    from witw_b200 import ops
    dev = torch.device("cuda")
    out = {"lib": os.path.basename(path)}
    for fov in (360, 90):
        sw = int(fov / 360 * 512) // 8
        gen = torch.Generator(device=dev).manual_seed(0)
        n = 10000
        ov = torch.randn(n, 16, 4, 64, device=dev, generator=gen) * 0.06
        su = torch.randn(n, 16, 4, sw, device=dev, generator=gen) * 0.06
        gal, qry = ops.GalleryIndex(ov, sw), ops.QueryBatch(su)
        times, full = [], []
        for i in range(8):
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.RankEvaluation(gal, qry, topk=10, events=ev).result()
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                times.append(ev[0].elapsed_time(ev[1]))
                full.append(a.elapsed_time(b))
        out["fov%d_sweep_ms" % fov] = round(sum(times) / len(times), 4)
        out["fov%d_eval_ms" % fov] = round(sum(full) / len(full), 4)
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        names = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "witw_b200", "libwitw_var_*.so")))
        for n in names:
            path = n if os.path.isabs(n) else os.path.join(ROOT, "witw_b200", n if n.endswith(".so") else "libwitw_var_%s.so" % n)
            subprocess.call([sys.executable, os.path.abspath(__file__), "--child", path])
