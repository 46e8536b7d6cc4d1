#!/usr/bin/env python
"""Timing of the spectral sweep (csrc/match_spec.cu) against the dense-contraction sweep at 10k x 10k.
Run once per cluster size: WITW_SPEC_CS=1|2|4|8 python tools/spec_bench.py [fov ...]  (one JSON line per case)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from witw_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libwitw_b200_hooks.so")   # the WITW_* switches exist only in the hooks build (make -C witw_b200/csrc HOOKS=1)
from witw_b200 import ops


def timeit(fn, iters=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
fovs = [int(a) for a in sys.argv[1:]] or [360]
impls = os.environ.get("SPEC_BENCH_IMPLS", "spectral,hankel").split(",")
G = Q = int(os.environ.get("SPEC_BENCH_N", "10000"))
for fov in fovs:
    sw = int(fov / 360 * 512) // 8
    ov = torch.randn(G, 16, 4, 64, device=dev, generator=gen) * 0.06
    su = torch.randn(Q, 16, 4, sw, device=dev, generator=gen) * 0.06
    d_true = torch.full((Q,), 1.0, device=dev)
    t32 = torch.arange(Q, dtype=torch.int32, device=dev)
    cnt = torch.zeros(Q, dtype=torch.int32, device=dev)
    for impl in impls:
        gal, qry = ops.GalleryIndex(ov, sw, impl=impl), ops.QueryBatch(su, impl=impl)
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ms_all = timeit(lambda: ops.sweep_tc(gal, qry, d_true=d_true, true_idx=t32, rank_count=cnt, topk=16, events=ev))
        torch.cuda.synchronize()
        ms_k = ev[0].elapsed_time(ev[1])
        ms_cnt = timeit(lambda: ops.sweep_tc(gal, qry, d_true=d_true, true_idx=t32, rank_count=cnt))
        ms_gp = timeit(lambda: ops.GalleryIndex(ov, sw, keep_fp32=False, impl=impl))
        ms_qp = timeit(lambda: ops.QueryBatch(su, keep_fp32=False, impl=impl))
        print(json.dumps({"impl": impl, "fov": fov, "G": G, "Q": Q, "cs": os.environ.get("WITW_SPEC_CS", "default"),
                          "debug": os.environ.get("WITW_SPEC_DEBUG", "0"),
                          "sweep_topk16_merge_ms": ms_all, "sweep_kernel_ms": ms_k, "sweep_count_only_ms": ms_cnt,
                          "gallery_prep_ms": ms_gp, "query_prep_ms": ms_qp,
                          "eff_tflops": 2.0 * 64 * 64 * sw * G * Q / ms_k / 1e9, "queries_per_s": Q / ms_k * 1e3}), flush=True)
        del gal, qry
