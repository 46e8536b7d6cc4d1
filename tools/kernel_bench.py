#!/usr/bin/env python
"""Per-kernel roofline measurements (CUDA events, inputs larger than L2, >=3 warm-ups).
Prints one JSON line per kernel; `python tools/kernel_bench.py > profiles/kernels_rNN.jsonl` on the GPU box."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import witw_b200 as W
from witw_b200 import ops

peaks = {"hbm_gbs": 6455.6, "bf16_tflops": 1663.2, "bf16_tflops_sustained": 1397.8}
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.isfile(pk):
    peaks.update(json.load(open(pk)))


def timeit(fn, iters=10, warm=3):
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


def emit(name, ms, unit, achieved, peak, extra=None):
    line = {"kernel": name, "ms": ms, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak}
    line.update(extra or {})
    print(json.dumps(line), flush=True)


dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)

# K1 polar: 1024 tiles x 3 planes (805 MB in, 805 MB out)
for c, n in ((3, 1024), (5, 600)):
    tiles = torch.randn(n, c, 256, 256, device=dev, generator=gen)
    ms = timeit(lambda: W.polar_transform(tiles))
    byts = 4.0 * n * c * (256 * 256 + 128 * 512)
    emit("polar_quadrant_kernel C=%d" % c, ms, "GB/s", byts / ms / 1e6, peaks["hbm_gbs"], {"tiles": n, "tiles_per_s": n / ms * 1e3})
    ms = timeit(lambda: W.polar_transform(tiles, exact=True), iters=5)
    emit("bilinear_gather_kernel (exact) C=%d" % c, ms, "GB/s", byts / ms / 1e6, peaks["hbm_gbs"], {"tiles": n, "tiles_per_s": n / ms * 1e3})
    del tiles

# K1 fused with ImageNormalization: uint8 tiles in, normalised fp32 polar out
tiles8 = torch.randint(0, 256, (2048, 3, 256, 256), device=dev, dtype=torch.uint8, generator=gen)
ms = timeit(lambda: W.normalized_polar(tiles8))
byts = 2048 * 3 * (256 * 256 * 1.0 + 128 * 512 * 4.0)
emit("polar_quadrant_kernel<uint8> C=3 (normalise + polar)", ms, "GB/s", byts / ms / 1e6, peaks["hbm_gbs"], {"tiles": 2048, "tiles_per_s": 2048 / ms * 1e3})
ms = timeit(lambda: W.normalized_polar(tiles8, exact=True), iters=5)
emit("bilinear_gather_kernel<uint8> (exact) C=3", ms, "GB/s", byts / ms / 1e6, peaks["hbm_gbs"], {"tiles": 2048, "tiles_per_s": 2048 / ms * 1e3})
del tiles8

# Resize + ImageNormalization front end (csrc/resize.cu): CVUSA-sized raw images in, model-sized normalised images out
# algorithmic bytes: every source byte once + every output byte once
for name, shape, out_hw, aa in (("aerial 750x750 u8 -> 256x256, antialias", (512, 3, 750, 750), (256, 256), True),
                                ("aerial 750x750 u8 -> 256x256, no antialias", (512, 3, 750, 750), (256, 256), False),
                                ("panorama 224x1232 u8 -> 128x512, antialias", (1024, 3, 224, 1232), (128, 512), True),
                                ("tile 256x256 u8 -> 256x256 (normalise only)", (2048, 3, 256, 256), (256, 256), False)):
    raw = torch.randint(0, 256, shape, device=dev, dtype=torch.uint8, generator=gen)
    ms = timeit(lambda: W.resize_normalize(raw, out_hw[0], out_hw[1], aa, mean=ops.IMG_MEAN, std=ops.IMG_STD))
    byts = shape[0] * shape[1] * (shape[2] * shape[3] * 1.0 + out_hw[0] * out_hw[1] * 4.0)
    emit("resize_norm_kernel " + name, ms, "GB/s", byts / ms / 1e6, peaks["hbm_gbs"], {"images": shape[0], "images_per_s": shape[0] / ms * 1e3})
    del raw
raw_ov = torch.randint(0, 256, (512, 3, 750, 750), device=dev, dtype=torch.uint8, generator=gen)
raw_su = torch.randint(0, 256, (512, 3, 224, 1232), device=dev, dtype=torch.uint8, generator=gen)
ms = timeit(lambda: W.prepare_pair(raw_su, raw_ov, fov=360, panorama=True, start=77))
emit("prepare_pair (resize+norm surface, resize+norm aerial, polar) 512 CVUSA-sized pairs", ms, "GB/s",
     512 * 3 * (750 * 750 + 224 * 1232 + 4.0 * (128 * 512 * 2 + 256 * 256 * 2 + 128 * 512)) / ms / 1e6, peaks["hbm_gbs"],
     {"pairs": 512, "pairs_per_s": 512 / ms * 1e3})
del raw_ov, raw_su

if os.environ.get("KB_ONLY") in ("polar", "resize"):
    sys.exit(0)

# K4 rank / top-k on a materialised 10k x 10k matrix (400 MB)
d = torch.rand(10000, 10000, device=dev, generator=gen)
ms = timeit(lambda: W.rank_from_distances(d))
emit("rank_count_vec4_kernel 10k x 10k", ms, "GB/s", 4e8 / ms / 1e6, peaks["hbm_gbs"])
ms = timeit(lambda: W.topk_from_distances(d, 10), iters=3, warm=1)
emit("topk_from_distances k=10 10k x 10k (threshold + filter + select)", ms, "GB/s", 4e8 / ms / 1e6, peaks["hbm_gbs"])
del d

# K2/K3 tensor-core sweeps with the gallery prepared once: 360 / 90 degrees, 10k x 10k
for fov in (360, 90):
    sw = int(fov / 360 * 512) // 8
    ov = torch.randn(10000, 16, 4, 64, device=dev, generator=gen) * 0.06
    su = torch.randn(10000, 16, 4, sw, device=dev, generator=gen) * 0.06
    d_true, _ = ops.true_match_distances(ov, su)
    t32 = torch.arange(10000, dtype=torch.int32, device=dev)
    cnt = torch.zeros(10000, dtype=torch.int32, device=dev)
    flop = 2.0 * 64 * 64 * sw * 1e8
    for impl, kname in (("spectral", "match_spec_kernel"), ("hankel", "match_tc_kernel")):
        gal, qry = ops.GalleryIndex(ov, sw, impl=impl), ops.QueryBatch(su, impl=impl)
        ms = timeit(lambda: ops.sweep_tc(gal, qry, d_true=d_true, true_idx=t32, rank_count=cnt, topk=10), iters=5)
        emit("%s fov=%d 10k x 10k (+rank count, top-10, merge)" % (kname, fov), ms, "TFLOP/s", flop / ms / 1e9, peaks["bf16_tflops_sustained"],
             {"queries_per_s": 1e4 / ms * 1e3, "flop_count": "direct form, 8192*sw per pair"})
        ms = timeit(lambda: ops.sweep_tc(gal, qry, want_dist=True, want_ori=True), iters=5)
        emit("%s fov=%d 10k x 10k (full dist+ori matrices)" % (kname, fov), ms, "TFLOP/s", flop / ms / 1e9, peaks["bf16_tflops_sustained"])
        ms = timeit(lambda: ops.GalleryIndex(ov, sw, impl=impl), iters=5)
        emit("gallery prep (%s, incl. fp32 spectra for spectral) fov=%d 10k items" % (impl, fov), ms, "GB/s",
             (gal.operand.numel() + ov.numel() * 4 * (2 if impl == "spectral" else 1)) / ms / 1e6, peaks["hbm_gbs"])
        del gal, qry
    del ov, su

# BASELINE configs[4] sweep: 100k-tile gallery, 10k queries, 90 degrees (gallery operand 7.2 GB)
ov = torch.randn(100000, 16, 4, 64, device=dev, generator=gen) * 0.06
su = torch.randn(10000, 16, 4, 16, device=dev, generator=gen) * 0.06
d_true = torch.full((10000,), 1.0, device=dev)
cnt = torch.zeros(10000, dtype=torch.int32, device=dev)
for impl, kname in (("spectral", "match_spec_kernel"), ("hankel", "match_tc_kernel")):
    gal, qry = ops.GalleryIndex(ov, 16, keep_fp32=False, impl=impl), ops.QueryBatch(su, keep_fp32=False, impl=impl)
    ms = timeit(lambda: ops.sweep_tc(gal, qry, d_true=d_true, rank_count=cnt, topk=10), iters=3, warm=1)
    emit("%s fov=90 100k gallery x 10k queries (+rank count, top-10, merge)" % kname, ms, "TFLOP/s", 2.0 * 64 * 64 * 16 * 1e9 / ms / 1e9,
         peaks["bf16_tflops_sustained"], {"queries_per_s": 1e4 / ms * 1e3})
    del gal, qry
del ov, su

# exact fp32 path, 2k x 2k at 360 degrees
ov = torch.randn(2048, 16, 4, 64, device=dev, generator=gen) * 0.06
su = torch.randn(2048, 16, 4, 64, device=dev, generator=gen) * 0.06
ms = timeit(lambda: W.match(ov, su, path="fp32"), iters=3, warm=1)
emit("match_tile_kernel fp32 2k x 2k fov=360", ms, "TFLOP/s", 2.0 * 64 * 64 * 64 * 2048 * 2048 / ms / 1e9, 70.0, {"peak_note": "nominal fp32 FMA peak ~70 TFLOP/s"})
