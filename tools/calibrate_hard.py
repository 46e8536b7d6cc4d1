#!/usr/bin/env python
"""Recall of the bench workload as a function of the planted noise (10k x 10k): picks bench.py's HARD_NOISE."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import witw_b200 as W
dev = torch.device("cuda")
for fov, noises in ((360, (8, 10, 12, 14, 16, 20, 25)), (90, (3, 4, 5, 6, 8, 10))):
    bench.SW = int(fov / 360 * 512) // 8
    for noise in noises:
        ov, su = bench.make_data(torch, dev, 10000, 10000, seed=300, noise=float(noise))
        ranks = W.evaluate_ranks(ov, su, path="tc")
        st = W.ops.evaluate_ranks_prepared.last_stats
        rec = W.recall_from_ranks(ranks)
        print(json.dumps({"fov": fov, "noise": noise, "top_one": float(rec["top_one"]), "top_percent": float(rec["top_percent"]), "median": float(rec["median"]),
                          "deferred": int(st["deferred"].sum()), "max_per_query": int(st["deferred"].max()), "flagged": st["flagged"]}), flush=True)
