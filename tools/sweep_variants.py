#!/usr/bin/env python
"""Sweep-kernel time of library builds on the bench workload (planted, noise 0.5), raw vs deferral mode, with and without the fused
top-k: every witw_b200/libwitw_*.so found (the shipped library, the hooks build, hand-built experiment variants) in its own
subprocess.  This is how the code-generation trap of the spectral sweep's epilogue was found (profiles/sweep_codegen_r2c.jsonl)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def child(path):
    sys.path.insert(0, ROOT)
    import torch
    from witw_b200 import _lib
    _lib.LIB_PATH = path
    from witw_b200 import ops
    import bench
    dev = torch.device("cuda")
    out = {"lib": os.path.basename(path)}
    for fov in (360, 90):
        bench.SW = sw = int(fov / 360 * 512) // 8
        ov, su = bench.make_data(torch, dev, 10000, 10000, seed=100, noise=0.5)
        gal, qry = ops.GalleryIndex(ov, sw), ops.QueryBatch(su)
        pq = torch.arange(10000, device=dev)
        d_true, _ = ops.pair_distances_prepared(gal, qry, pq, pq)
        t32 = pq.to(torch.int32)
        for mode in ("raw", "defer", "raw_notopk", "defer_notopk"):
            times = []
            for i in range(8):
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                cnt = torch.zeros(10000, dtype=torch.int32, device=dev)
                defer = ops.Deferral(10000, 10000, dev) if mode.startswith("defer") else None
                ops.sweep_tc(gal, qry, d_true=d_true, true_idx=t32, rank_count=cnt, topk=0 if mode.endswith("notopk") else 16, events=ev, deferral=defer)
                torch.cuda.synchronize()
                if i >= 3:
                    times.append(ev[0].elapsed_time(ev[1]))
            out["fov%d_%s" % (fov, mode)] = round(sum(times) / len(times), 4)
    print(json.dumps(out))

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        names = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "witw_b200", "libwitw_*.so")))
        for n in names:
            path = n if os.path.isabs(n) else os.path.join(ROOT, "witw_b200", n)
            subprocess.call([sys.executable, os.path.abspath(__file__), "--child", path])
