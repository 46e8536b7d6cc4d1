#!/usr/bin/env python
"""Spectral sweep, one CTA per tile (variant 1) against CTA pairs (variant 2): the raw fp16 outputs of both must agree bit for
bit (same operands, same accumulation order), and the sweep-kernel time of both on the bench workload."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from witw_b200 import _lib, ops
import bench


def raw(gal, qry, variant, topk=16):
    _lib.call("witw_match_spec_variant", variant)
    q = qry.Q
    pq = torch.arange(q, device=gal.device) % gal.G
    d_true, _ = ops.pair_distances_prepared(gal, qry, pq, torch.arange(q, device=gal.device))
    cnt = torch.zeros(q, dtype=torch.int32, device=gal.device)
    r = ops.sweep_tc(gal, qry, want_dist=True, want_ori=True, d_true=d_true, true_idx=pq.to(torch.int32), rank_count=cnt, topk=topk)
    torch.cuda.synchronize()
    return r["dist"], r["ori"], cnt, r["topk_dist"], r["topk_idx"]


def main():
    dev = torch.device("cuda")
    ok = True
    for (g, q, fov) in ((1000, 300, 360), (1003, 129, 90), (8, 1, 360), (9, 130, 360), (4099, 515, 180), (17, 64, 360)):
        bench.SW = sw = int(fov / 360 * 512) // 8
        ov, su = bench.make_data(torch, dev, g, q, seed=7, noise=3.0)
        gal, qry = ops.GalleryIndex(ov, sw), ops.QueryBatch(su)
        a = raw(gal, qry, 1)
        b = raw(gal, qry, 2)
        same = [bool(torch.equal(x, y)) if x.dtype != torch.float32 else bool(torch.equal(x.view(torch.int32), y.view(torch.int32))) for x, y in zip(a, b)]
        print(json.dumps({"G": g, "Q": q, "fov": fov, "dist_ori_count_topkd_topki_equal": same,
                          "max_abs_dist_diff": float((a[0] - b[0]).abs().nan_to_num(0).max())}), flush=True)
        ok = ok and all(same)
    out = {}
    for fov in (360, 90):
        bench.SW = sw = int(fov / 360 * 512) // 8
        ov, su = bench.make_data(torch, dev, 10000, 10000, seed=100, noise=0.5)
        gal, qry = ops.GalleryIndex(ov, sw), ops.QueryBatch(su)
        pq = torch.arange(10000, device=dev)
        d_true, _ = ops.pair_distances_prepared(gal, qry, pq, pq)
        t32 = pq.to(torch.int32)
        for variant in (1, 2):
            _lib.call("witw_match_spec_variant", variant)
            for mode in ("raw", "defer"):
                times = []
                for i in range(8):
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    cnt = torch.zeros(10000, dtype=torch.int32, device=dev)
                    defer = ops.Deferral(10000, 10000, dev) if mode == "defer" else None
                    ops.sweep_tc(gal, qry, d_true=d_true, true_idx=t32, rank_count=cnt, topk=16, events=ev, deferral=defer)
                    torch.cuda.synchronize()
                    if i >= 3:
                        times.append(ev[0].elapsed_time(ev[1]))
                out["fov%d_v%d_%s" % (fov, variant, mode)] = round(sum(times) / len(times), 4)
    print(json.dumps(out), flush=True)
    _lib.call("witw_match_spec_variant", 2)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
