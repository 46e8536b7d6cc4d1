#!/usr/bin/env python
"""Benchmark of the WITW retrieval hot path on B200 (contract: see the task description).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): cvig_fov.py 360-degree eval, 10k queries x 10k gallery,
orientation-searched distance + rank counting + top-10, synthetic feature maps [N,16,4,64].
One step = one full pass from fp32 feature maps resident in HBM to per-query ranks and top-k:
gallery/query operand prep -> exact fp32 true-match distances -> tcgen05 sweep (fused argmax,
crop-normalise, distance, rank count, top-k; decisions its fp16 operands cannot settle are deferred)
-> top-k merge -> fp32 finish of the deferred pairs and of the top-k.
--sweep spectral (default: what the library picks) evaluates the circular correlation through the
correlation theorem (csrc/match_spec.cu: per-frequency tcgen05 products + in-register inverse FFT);
--sweep hankel is the dense contraction over all 64 shifts (csrc/match_tc.cu).
N > 1: the gallery is sharded, one 10k-item shard per GPU (weak scaling: gallery_total = N*10k),
queries replicated; the [Q] true distances before the sweep and [counts | top-k] after it are exchanged
over peer memory (NVLink stores from the library's own kernels, witw_b200/peer.py; --no-peer-exchange: an
NCCL all-reduce and one all-gather).  value = N*Q / t: queries swept per second, each
against a 10k-item gallery shard (see config.unit_note for the whole-gallery figure).

Next to the timed step the line carries side measurements (N = 1; --no-extras skips them): the same step on data
whose ranks are non-trivial (`hard`), checked against the oracle on the same inputs; the other BASELINE configs
(90 degrees, 100k and 1M galleries); the dense sweep's tensor roofline; the HBM-bound kernels (polar resample, rank
count, top-k select); the reference's literal one-query loop through the drop-in functions; a gallery-resident
end-to-end figure.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

G_PER_GPU = 10000
Q_TOTAL = 10000
FOV = 360
TOPK = 10
NOISE = 0.5                            # planted matches: query = rolled, cropped gallery item + NOISE x its own scale of noise
HARD_NOISE = {64: 13.0, 16: 6.0}       # `hard` arm: noise at which about half of the queries are rank 1 (measured, 10k gallery)
SW = 64                                # query columns: int(fov/360*512)//8 (cvig_fov.py:22, 8 image pixels per feature column)
FLOP_PER_PAIR = 2 * 64 * 16 * 4 * SW   # 2*W*C*H*sw = 524 288 at 360 deg, 131 072 at 90 deg
METRIC = "queries/sec vs gallery size (orientation-searched distance + top-k)"
SPEC_TC_FLOP_PER_PAIR = 2.0 * 128 * 32 * 16 * 256 / 2048.0    # issued as tcgen05.mma: 256 MMAs of 128x32x16 per CTA pair and 2048 pairs
SPEC_TMA_BYTES_PER_PAIR = 32.0 * 20480 / 1024.0               # operand bytes TMA delivers: 32 slots x (16 KB queries + 4 KB items) per 1024 pairs


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed regions."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None
        self.first = 0

    def mark(self):
        """Samples taken before this call (sampler start-up, warm-up steps) are not reported."""
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            # wait until it has attached to the driver and delivers samples steadily: on a fresh box the first polls stall
            # kernel launches for tens of milliseconds (seen as one 40 ms step among 7.5 ms ones, 60 ms after the first sample)
            deadline = time.time() + 5.0
            while len(self.lines) < 4 and time.time() < deadline and self.proc.poll() is None:
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines[self.first:]:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def profile_json(name):
    path = os.path.join(ROOT, "profiles", name)
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f)
    return {}


def measured_peaks():
    """(hbm GB/s, bf16 TFLOP/s burst, sustained, source) from the driver-written MEASURED_PEAKS.json, else the recipe's fallbacks."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("hbm_gbs", 6550.0), p.get("bf16_tflops", 1645.0), p.get("bf16_tflops_sustained", 1364.0), "MEASURED_PEAKS.json"
    return 6400.0, 1650.0, 1400.0, "fallback (B200_PROFILING.md)"


def config_name(g_total):
    if (FOV, Q_TOTAL, G_PER_GPU) == (360, 10000, 10000):
        return "configs[1]"
    if (FOV, Q_TOTAL, G_PER_GPU) == (90, 10000, 10000):
        return "configs[2]"
    if FOV == 360 and g_total == 1000000:
        return "configs[3]"
    if FOV == 90 and g_total == 100000:
        return "configs[4], correlation-sweep half"
    return "a variation of configs[1]"


def workload_config(world):
    """The `config` object of the JSON line: identical for both arms (the driver compares them)."""
    g_total = world * G_PER_GPU
    return {
        "workload": "cvig_fov %ddeg eval: %d queries x %d-item gallery per GPU, orientation-searched distance + rank count + top-%d "
                    "(BASELINE %s)" % (FOV, Q_TOTAL, G_PER_GPU, TOPK, config_name(g_total)),
        "gallery_total": g_total, "gallery_per_gpu": G_PER_GPU, "queries": Q_TOTAL, "fov": FOV, "feature_shape": [16, 4, 64],
        "planted_noise": NOISE,
        "l2_policy": "inputs larger than L2: every step re-reads %d MB of fp32 features and re-writes their operands and fp32 spectra "
                     "(126 MB L2); no flush between steps" % ((G_PER_GPU * 64 * 64 + Q_TOTAL * 64 * SW) * 4 // 1000000),
    }


# ----------------------------------------------------------------------------- CPU arm (the reference's PyTorch path)
def synth_cpu(n_gallery, n_query, noise, seed):
    """Host copy of the workload's construction (make_data below) for the reference arm, which has no GPU."""
    import torch

    gen = torch.Generator().manual_seed(seed)
    ov = torch.randn(n_gallery, 16, 4, 64, generator=gen) * 0.06
    su = torch.randn(n_query, 16, 4, SW, generator=gen) * 0.06
    n = min(n_gallery, n_query)
    shifts = torch.randint(0, 64, (n,), generator=gen)
    cols = (shifts.view(n, 1) + torch.arange(SW).view(1, SW)) % 64
    su[:n] = torch.gather(ov[:n], 3, cols.view(n, 1, 1, SW).expand(n, 16, 4, SW)) + noise * su[:n]
    return ov, su


def cpu_rank_sample(ov, su, budget_s, true_idx=None, max_queries=64):
    """The reference's CPU path (oracle port: conv2d / argmax / gather / norm, cvig_fov.py:545-552) on the host cores for
    as many of the first queries as fit the budget, each against the whole gallery.  Returns (queries/s, cores, ranks,
    ties, seconds): ties[i] = gallery items whose fp32 distance is within 3e-6 of query i's threshold (rank decisions that
    fp32 round-off alone can flip)."""
    import numpy as np
    import torch

    from oracle import witw_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)

    def one(i):
        _, dist = O.match(ov, su[i: i + 1])
        dist = dist.squeeze()
        thr = dist[i if true_idx is None else int(true_idx[i])]
        return int(torch.sum(torch.le(dist, thr)).item()), int(((dist - thr).abs() <= 3e-6).sum().item()) - 1

    t0 = time.perf_counter()
    first = one(0)                                   # warm-up, also the calibration sample
    per_q = time.perf_counter() - t0
    n = int(max(2, min(max_queries, su.shape[0], budget_s / max(per_q, 1e-3))))
    t0 = time.perf_counter()
    res = [one(i) for i in range(n)]
    dt = time.perf_counter() - t0
    del first
    return n / dt, cores, np.array([r[0] for r in res]), np.array([r[1] for r in res]), dt


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    world = max(1, args.gpus)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # each step is a bounded sample; the whole run stays within a few minutes
    per_step_budget = max(2.0, min(15.0, 150.0 / (steps + warmup)))
    ov, su = synth_cpu(G_PER_GPU, min(Q_TOTAL, 64), NOISE, seed=100)
    vals, n, cores = [], 0, 1
    for i in range(warmup + steps):
        v, cores, ranks, _, dt = cpu_rank_sample(ov, su, per_step_budget)
        n = len(ranks)
        if i >= warmup:
            vals.append(v)
    value = statistics.mean(vals)
    sample = "first %d queries of the %d-query set against the full %d-item gallery per step" % (n, Q_TOTAL, G_PER_GPU)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1000.0 * Q_TOTAL / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(world),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "CPU port of the reference's PyTorch path (oracle/witw_oracle.py: the same F.conv2d / argmax / gather / norm calls); "
                                 "a Python reference cannot travel to the GPU box"},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- GPU arm
def make_data(torch, device, n_gallery, n_query, seed, noise, sw=None, true_idx=None):
    """Synthetic feature maps of the BASELINE shapes generated on the device: gallery ~ N(0, 0.06^2); query i
    is its matching gallery item (i, or true_idx[i]) rolled by a random azimuth, cropped to the field of view, plus
    noise x N(0, 0.06^2)."""
    sw = SW if sw is None else sw
    gen = torch.Generator(device=device).manual_seed(seed)
    ov = torch.randn(n_gallery, 16, 4, 64, generator=gen, device=device) * 0.06
    su = torch.randn(n_query, 16, 4, sw, generator=gen, device=device) * 0.06
    if true_idx is None:
        n = min(n_gallery, n_query)
        src = ov[:n]
    else:
        n = n_query
        src = ov[true_idx]
    shifts = torch.randint(0, 64, (n,), generator=gen, device=device)
    cols = (shifts.view(n, 1) + torch.arange(sw, device=device).view(1, sw)) % 64
    su[:n] = torch.gather(src, 3, cols.view(n, 1, 1, sw).expand(n, 16, 4, sw)) + noise * su[:n]
    return ov, su


def device_step_ms(torch, ops, ov, su, true_idx=None, iters=5, warm=3, sink=None):
    """Average device time of the whole step (prep -> ranks + top-k) on resident features, pipelined one step deep."""
    def launch():
        sw = su.shape[3]
        ev = None
        if sink is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            sink.append(ev)
        return ops.RankEvaluation(ops.GalleryIndex(ov, sw), ops.QueryBatch(su), true_idx=true_idx, topk=TOPK, events=ev)

    out = None
    prev = None
    for _ in range(warm + 1):             # pipelined like the timed loop, so that both sets of buffers exist before it starts
        cur = launch()
        if prev is not None:
            out = prev.result()
        prev = cur
    out = prev.result()
    torch.cuda.synchronize()
    if sink is not None:
        del sink[:]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    prev = None
    for _ in range(iters):
        cur = launch()
        if prev is not None:
            out = prev.result()
        prev = cur
    out = prev.result()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters, out


def safe(fn, *a, **kw):
    """A side measurement must not take the bench line down with it."""
    import torch
    try:
        return fn(*a, **kw)
    except Exception as exc:
        return {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    finally:
        torch.cuda.empty_cache()


def parity_against_cpu(torch, ranks, ov, su, budget_s, true_idx=None):
    """The GPU ranks of the first queries against the oracle's rank loop on the very same inputs (copied to the host)."""
    import numpy as np

    qps, cores, want, ties, dt = cpu_rank_sample(ov.cpu(), su.cpu(), budget_s, true_idx=None if true_idx is None else true_idx.cpu())
    got = ranks[: len(want)].cpu().numpy()
    diff = np.abs(got - want)
    return {"queries": int(len(want)), "ranks_equal": bool(np.all(diff <= ties)), "ranks_identical": int(np.sum(diff == 0)),
            "max_rank_diff": int(diff.max()), "fp32_ties_at_threshold": int(ties.sum()),
            "oracle": "oracle.match per query on the host, same inputs (cvig_fov.py:545-552)", "cpu_seconds": dt}, qps, cores


def dense_sweep_roofline(torch, ops, ov, su, burst, sustained, true_idx=None, iters=5):
    """The north star's kernel (2): the shift search as one dense fp16 contraction on tcgen05 (csrc/match_tc.cu), timed alone
    on the bench workload with CUDA events -- the tensor-pipe roofline figure that the spectral default cannot show, because
    the spectral sweep does 31x fewer tensor FLOPs per pair.  Outside the timed step; N=1 only."""
    gallery = ops.GalleryIndex(ov, SW, impl="hankel")
    queries = ops.QueryBatch(su, impl="hankel")
    evs = []
    for i in range(2 + iters):
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ops.evaluate_ranks_prepared(gallery, queries, true_idx=true_idx, topk=TOPK, events=ev)
        if i >= 2:
            evs.append(ev)
    torch.cuda.synchronize()
    ms = statistics.mean(a.elapsed_time(b) for a, b in evs)
    achieved = FLOP_PER_PAIR * float(ov.shape[0]) * float(su.shape[0]) / (ms / 1000.0) / 1e12
    del gallery, queries
    return {"kernel": "match_tc_kernel", "bound": "tensor", "kernel_ms": ms, "achieved": achieved, "unit": "TFLOP/s",
            "peak": burst, "frac": achieved / burst, "peak_sustained": sustained, "frac_sustained": achieved / sustained, "launches": iters,
            "note": "dense contraction over all 64 shifts, %d tensor FLOP per pair as executed, fp16 operands; same epilogue outputs (rank "
                    "count, deferral, top-k candidates) as the step's spectral sweep; peak = cuBLAS bf16 burst (the kernel is timed alone), "
                    "peak_sustained = the figure under a long step; ncu: profiles/match_tc_r1c.txt" % FLOP_PER_PAIR}


def config_step(torch, ops, device, n_gallery, fov, noise=NOISE, iters=5, warm=2, n_query=Q_TOTAL):
    """The device-resident step at another BASELINE config (field of view / gallery size)."""
    sw = int(fov / 360 * 512) // 8
    true_idx = None if n_query <= n_gallery else torch.arange(n_query, device=device) % n_gallery
    ov, su = make_data(torch, device, n_gallery, n_query, seed=7, noise=noise, sw=sw, true_idx=true_idx)
    sink = []
    ms, out = device_step_ms(torch, ops, ov, su, true_idx=true_idx, iters=iters, warm=warm, sink=sink)
    torch.cuda.synchronize()
    kms = statistics.mean(a.elapsed_time(b) for a, b in sink[-iters:])
    stats = ops.evaluate_ranks_prepared.last_stats
    rec = ops.recall_from_ranks(out[0])
    return {"gallery": n_gallery, "queries": n_query, "fov": fov, "noise": noise, "ms_per_step": ms, "sweep_kernel_ms": kms,
            "queries_per_s": n_query / (ms / 1000.0), "pairs_per_s": float(n_gallery) * n_query / (ms / 1000.0),
            "recall_top_one": float(rec["top_one"]), "deferred_pairs": int(stats["deferred"].sum()), "flagged_queries": int(stats["flagged"])}


def hard_arm(torch, ops, device, budget_s):
    """The timed step on data whose ranks are not trivial (about half of the queries rank 1): how many pairs the sweep defers
    to fp32, what that costs, and the ranks against the oracle on the same inputs."""
    noise = HARD_NOISE.get(SW, 13.0 * SW / 64.0)
    ov, su = make_data(torch, device, G_PER_GPU, Q_TOTAL, seed=300, noise=noise)
    sink = []
    ms, out = device_step_ms(torch, ops, ov, su, iters=5, warm=3, sink=sink)
    torch.cuda.synchronize()
    kms = statistics.mean(a.elapsed_time(b) for a, b in sink[-5:])
    stats = ops.evaluate_ranks_prepared.last_stats
    deferred = stats["deferred"]
    rec = ops.recall_from_ranks(out[0])
    parity, _, _ = parity_against_cpu(torch, out[0], ov, su, budget_s)
    return {"noise": noise, "ms_per_step": ms, "sweep_kernel_ms": kms, "queries_per_s": Q_TOTAL / (ms / 1000.0),
            "recall": {k: float(v) for k, v in rec.items()},
            "deferral": {"deferred_pairs": int(deferred.sum()), "fraction_of_pairs": float(deferred.sum()) / (float(G_PER_GPU) * Q_TOTAL),
                         "max_per_query": int(deferred.max()), "list_capacity_per_query": int(stats["list_cap"]),
                         "dropped": 0, "queries_redone_in_fp32": int(stats["flagged"])},
            "parity_check": parity}


def hbm_kernels(torch, W, hbm):
    """The HBM-bound kernels of the path against the measured copy bandwidth: K1 polar resample (3- and 5-channel tiles),
    K4 rank count and top-k select on a materialised 10k x 10k matrix."""
    def timeit(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    out = {}
    for name, n, c in (("polar_c3", 1024, 3), ("polar_c5", 600, 5)):
        tiles = torch.randn(n, c, 256, 256, device="cuda")
        ms = timeit(lambda: W.polar_transform(tiles))
        nbytes = 4.0 * n * c * (256 * 256 + 128 * 512)
        out[name] = {"kernel": "polar_quadrant_kernel", "tiles": n, "channels": c, "ms": ms, "tiles_per_s": n / (ms / 1000.0),
                     "achieved": nbytes / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": nbytes / ms / 1e6 / hbm}
        del tiles
    tiles = torch.randint(0, 256, (2048, 3, 256, 256), device="cuda", dtype=torch.uint8)
    ms = timeit(lambda: W.normalized_polar(tiles))
    nbytes = 2048.0 * 3 * (256 * 256 + 4 * 128 * 512)
    out["polar_u8_c3"] = {"kernel": "polar_quadrant_kernel<u8>", "tiles": 2048, "ms": ms, "tiles_per_s": 2048 / (ms / 1000.0),
                          "achieved": nbytes / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": nbytes / ms / 1e6 / hbm}
    del tiles
    dist = torch.rand(10000, 10000, device="cuda") + 1.0
    ms = timeit(lambda: W.rank_from_distances(dist))
    out["rank_count"] = {"kernel": "rank_count_kernel", "matrix": [10000, 10000], "ms": ms, "achieved": 4e8 / ms / 1e6, "peak": hbm, "unit": "GB/s",
                         "frac": 4e8 / ms / 1e6 / hbm}
    ms = timeit(lambda: W.topk_from_distances(dist, 10))
    out["topk_select"] = {"kernel": "topk threshold + filter + select", "matrix": [10000, 10000], "k": 10, "ms": ms, "achieved": 4e8 / ms / 1e6,
                          "peak": hbm, "unit": "GB/s", "frac": 4e8 / ms / 1e6 / hbm}
    return out


def streamed_polar(torch, W, hbm, n_tiles=100000, batch=512):
    """BASELINE configs[4], first half: 100k five-channel tiles (131 GB as fp32: they do not fit the device at once) stream from
    pinned host memory through the fused ImageNormalization + PolarTransform kernel as uint8, batch i+1 uploading while batch i
    is transformed (witw_b200.streamed_polar).  The host side is a ring of 8 pinned batches reused cyclically."""
    mean, std, div = (0.485, 0.456, 0.406, 0.5, 0.5), (0.229, 0.224, 0.225, 0.25, 0.25), (255.0, 255.0, 255.0, 1.0, 1.0)
    ring = [torch.randint(0, 256, (batch, 5, 256, 256), dtype=torch.uint8).pin_memory() for _ in range(4)]
    n_batches = (n_tiles + batch - 1) // batch

    def batches(n):
        for i in range(n):
            yield ring[i % len(ring)]

    def run(n):
        last = None
        for polar in W.streamed_polar(batches(n), mean, std, div):
            last = polar                      # the encoder would consume it here; keeping one reference lets the allocator reuse the rest
        return last

    run(4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = run(n_batches)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # the upload alone, for the floor under it
    dev = torch.empty_like(ring[0], device="cuda")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(16):
        dev.copy_(ring[i % len(ring)], non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    h2d_gbs = 16 * ring[0].numel() / a.elapsed_time(b) / 1e6
    tiles = n_batches * batch
    in_bytes = float(tiles) * 5 * 65536
    return {"tiles": tiles, "channels": 5, "batch": batch, "seconds": dt, "tiles_per_s": tiles / dt, "h2d_gbs_while_streaming": in_bytes / dt / 1e9,
            "h2d_alone_gbs": h2d_gbs, "h2d_floor_s": in_bytes / (h2d_gbs * 1e9), "output_shape": list(out.shape),
            "device_kernel_gbs": None,
            "note": "uint8 tiles over PCIe (a quarter of the fp32 bytes), normalised polar images [n,5,128,512] fp32 left on the device for the "
                    "encoder; end-to-end rate is the upload's: the kernel needs %.2f ms per batch" % (1e3 * batch * 5 * (65536 + 4 * 65536) / (0.57 * hbm * 1e9))}


def dropin_loop(torch, W, ov, su, n_queries=64):
    """The reference's own loop body (cvig_fov.py:545-552), unmodified, running on the rebound names after install():
    one query at a time against the whole gallery, a device -> host read per query."""
    import types

    import numpy as np

    mod = types.ModuleType("cvig_like")
    for n in ("bilinear_interpolate", "PolarTransform", "correlation", "crop_overhead", "l2_distance"):
        setattr(mod, n, None)
    W.install(mod)
    correlation, crop_overhead, l2_distance = mod.correlation, mod.crop_overhead, mod.l2_distance
    surface_embed, overhead_embed = su, ov

    def loop(count):
        ranks = np.zeros([count], dtype=int)
        for idx in range(count):
            this_surface_embed = torch.unsqueeze(surface_embed[idx, :], 0)
            orientation_estimate = correlation(overhead_embed, this_surface_embed)
            overhead_cropped_all = crop_overhead(overhead_embed, orientation_estimate, this_surface_embed.shape[3])
            distances = l2_distance(overhead_cropped_all, this_surface_embed)
            distances = torch.squeeze(distances)
            distance = distances[idx]
            ranks[idx] = torch.sum(torch.le(distances, distance)).item()
        return ranks

    loop(4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ranks = loop(n_queries)
    dt = time.perf_counter() - t0
    W.ops.clear_cache()
    return {"queries": n_queries, "gallery": int(ov.shape[0]), "seconds": dt, "queries_per_s": n_queries / dt,
            "ranks_head": [int(r) for r in ranks[:8]],
            "note": "literal body of cvig_fov.py:545-552 on witw_b200.install()'s names: per query one fp32 column sweep from the cached "
                    "gallery spectra (correlation), a [G,1,16,4,sw] gather (crop_overhead), l2_distance, .item()"}


def resident_e2e(torch, W, ops, ov, su_host, true_idx, steps):
    """Gallery prepared once and resident (what GalleryIndex is for); per step the query set is uploaded from pinned host
    memory, prepared, swept and finished, and ranks + top-k are read back -- uploads on a copy stream one step ahead, results
    through pinned buffers on a third stream, as in the full end-to-end loop."""
    device = ov.device
    gallery = ops.GalleryIndex(ov, SW)
    bufs = [torch.empty(su_host.shape, device=device) for _ in range(2)]
    copy_stream, d2h_stream = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    computed = [torch.cuda.Event() for _ in range(2)]
    landed = [torch.cuda.Event() for _ in range(2)]
    host_out = [None, None]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[i % 2])
            bufs[i % 2].copy_(su_host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def download(b, tensors):
        computed[b].record()
        if host_out[b] is None:
            host_out[b] = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in tensors]
        else:
            landed[b].synchronize()
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(computed[b])
            for h, t in zip(host_out[b], tensors):
                h.copy_(t, non_blocking=True)
                t.record_stream(d2h_stream)
            landed[b].record(d2h_stream)

    def run(n):
        upload(0)
        pending = None
        for i in range(n):
            b = i % 2
            if i + 1 < n:
                upload(i + 1)
            torch.cuda.current_stream().wait_event(ready[b])
            cur = ops.RankEvaluation(gallery, ops.QueryBatch(bufs[b]), true_idx=true_idx, topk=TOPK)
            free[b].record()
            if pending is not None:
                download(pending[0], pending[1].result())
            pending = (b, cur)
        download(pending[0], pending[1].result())
        landed[pending[0]].synchronize()
        return tuple(h.clone() for h in host_out[pending[0]])

    run(4)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    out = run(steps)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    return {"ms_per_step": ms, "queries_per_s": Q_TOTAL / (ms / 1000.0), "h2d_bytes_per_step": su_host.numel() * 4,
            "d2h_bytes_per_step": sum(t.numel() * t.element_size() for t in out),
            "note": "gallery operand built once and kept in HBM; every step uploads the %d queries from pinned host memory (double-buffered "
                    "behind the previous step), prepares, sweeps, finishes and reads ranks + top-%d back" % (Q_TOTAL, TOPK)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import witw_b200 as W
    from witw_b200 import ops
    from witw_b200.sharded import evaluate_ranks_sharded

    if args.sweep != "auto":
        ops.TC_IMPL = args.sweep
    ops.L2_WINDOW = bool(args.l2_window)
    if args.no_peer_exchange:
        W.sharded.PEER_EXCHANGE = False
    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if world != args.gpus and world > 1:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries the one JSON line only (NCCL_DEBUG=VERSION prints there)
        dist.init_process_group("nccl", device_id=device)
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    g_offset = rank * G_PER_GPU
    g_total = world * G_PER_GPU
    # query i matches gallery item i; with more queries than gallery items the match indices wrap around
    true_idx = None if Q_TOTAL <= g_total else torch.arange(Q_TOTAL, device=device) % g_total
    # rank 0's shard holds the planted matches of the first min(G, Q) queries; the query set is the same on every rank
    ov, su = make_data(torch, device, G_PER_GPU, Q_TOTAL, seed=100 + rank, noise=NOISE,
                       true_idx=None if Q_TOTAL <= G_PER_GPU else torch.arange(Q_TOTAL, device=device) % G_PER_GPU)
    if world > 1:
        dist.broadcast(su, src=0)
    sweep_events = []
    timed_local = W.sharded.CudaLocal(event_sink=sweep_events)

    def launch_step():
        """fp32 features in HBM -> ranks (+ top-k), enqueued; .result() finishes it."""
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        sweep_events.append(ev)
        return ops.RankEvaluation(ops.GalleryIndex(ov, SW), ops.QueryBatch(su), true_idx=true_idx, topk=TOPK, events=ev)

    def run_steps(n):
        """n steps; on one GPU pipelined one step deep (step i+1 is enqueued before the host asks step i for its result)."""
        out = None
        if world == 1:
            prev = None
            t_prev = time.perf_counter()
            for i in range(n):
                cur = launch_step()
                if prev is not None:
                    out = prev.result()
                prev = cur
                if os.environ.get("WITW_BENCH_TRACE") == "1":
                    now = time.perf_counter()
                    sys.stderr.write("step %d: %.2f ms host\n" % (i, 1e3 * (now - t_prev)))
                    t_prev = now
            return prev.result()
        prev = None
        for _ in range(n):                             # same pipelining; the collectives of step i+1 are enqueued behind step i's
            cur = W.ShardedEvaluation(ov, su, g_offset, g_total, true_idx=true_idx, topk=TOPK, local=timed_local)
            if prev is not None:
                out = prev.result()
            prev = cur
        return prev.result()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        out = fn()
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    # the sampler starts before the warm-up: the first nvidia-smi on a fresh box takes a while to attach to the driver and
    # stalls kernel launches while it does (seen once as 12.6 instead of 7.65 ms per step when it started with the timed region)
    sampler = ClockSampler(local_rank)
    if rank == 0 and os.environ.get("WITW_BENCH_NO_SAMPLER") != "1":      # (debug: is a stall the sampler's doing?)
        sampler.start()
    # Spin-up before the W warm-up steps.  The first CUDA process on a box that has been idle sees one stall of 40 - 50 ms some
    # 60 ms into its first stretch of continuous work, with or without the sampler (measured: one 49.6 ms step among 7.5 ms
    # ones, in the first process only; later processes on the same box never) -- a power-state transition of the idle GPU.  A
    # second of untimed steps puts it behind us; the second of rest that follows lets the board's power average recover, so
    # that the timed steps start from the state a fresh run starts from (after a second of continuous work the power cap
    # lowers the SM clock and the sweep takes 7.1 instead of 6.8 ms: `spin_up.ms_per_step` reports that sustained figure).
    run_steps(3)                                       # first launches: module loads, allocator pools
    torch.cuda.synchronize()
    t_spin = time.perf_counter()
    run_steps(3)
    torch.cuda.synchronize()
    spin_s = float(os.environ.get("WITW_BENCH_SPIN", "1.0"))
    n_spin = torch.tensor([int(min(200, max(0, spin_s / max((time.perf_counter() - t_spin) / 3, 1e-4))))], device=device)
    if world > 1:
        dist.broadcast(n_spin, src=0)                  # every rank must run the same number of steps: they hold collectives
    n_spin = (int(n_spin.item()) // 5) * 5
    t_spin = time.perf_counter()
    for _ in range(n_spin // 5):
        run_steps(5)
    torch.cuda.synchronize()
    spin_ms = 1e3 * (time.perf_counter() - t_spin) / max(n_spin, 1)
    time.sleep(float(os.environ.get("WITW_BENCH_REST", "1.0")))
    run_steps(warmup)
    torch.cuda.synchronize()
    sweep_events.clear()
    sampler.mark()
    ms_total, out = timed(lambda: run_steps(steps))
    ms_step = ms_total / steps
    value = world * Q_TOTAL / (ms_step / 1000.0)
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in sweep_events[:steps])
    ranks = out[0]
    recall = W.recall_from_ranks(ranks)
    step_stats = ops.evaluate_ranks_prepared.last_stats

    # End to end through the public API with host buffers: every step copies its feature maps from pinned host
    # memory to the device, evaluates, and reads ranks and top-k back to the host.  The copies of step i+1 are
    # issued on a second stream into the other of two device buffers while step i computes (double buffering);
    # all of it, including the first copy, is inside the timed region.
    ov_host = ov.cpu().pin_memory()
    su_host = su.cpu().pin_memory()
    bufs = [(torch.empty_like(ov), torch.empty_like(su)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=device)

    # The replicated query set crosses PCIe once per step in the whole job: every rank uploads 1/N of it (with one rank uploading
    # all of it that rank carries twice the bytes of the others and paces the job) and the slices are all-gathered over NVLink.
    q_split = world > 1 and Q_TOTAL % world == 0
    q_lo, q_hi = (rank * (Q_TOTAL // world), (rank + 1) * (Q_TOTAL // world)) if q_split else (0, Q_TOTAL)
    q_part = [torch.empty_like(su[q_lo:q_hi]) for _ in range(2)] if q_split else None

    def upload(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[b])            # the step that last used this buffer has finished with it
            bufs[b][0].copy_(ov_host, non_blocking=True)
            if q_split:
                q_part[b].copy_(su_host[q_lo:q_hi], non_blocking=True)
            elif rank == 0:
                bufs[b][1].copy_(su_host, non_blocking=True)
            ready[b].record(copy_stream)

    # results leave the device the same way: step i's ranks and top-k are copied into pinned host buffers on a third stream
    # and read by the host while step i+1 is already running, so the launch latency of a step hides behind the previous one
    d2h_stream = torch.cuda.Stream(device=device)
    host_out = [None, None]
    computed = [torch.cuda.Event() for _ in range(2)]
    landed = [torch.cuda.Event() for _ in range(2)]

    def download(b, tensors):
        computed[b].record()
        if host_out[b] is None:
            host_out[b] = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in tensors]
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(computed[b])
            for h, t in zip(host_out[b], tensors):
                h.copy_(t, non_blocking=True)
                t.record_stream(d2h_stream)
            landed[b].record(d2h_stream)

    trace = os.environ.get("WITW_BENCH_TRACE") == "1"

    def run_e2e(n):
        out = None
        upload(0)
        pending = None                                 # (buffer index, RankEvaluation) of the step enqueued last
        t_prev = time.perf_counter()
        for i in range(n):
            b = i % 2
            if trace:
                now = time.perf_counter()
                sys.stderr.write("e2e iter %d: %.2f ms host\n" % (i, 1e3 * (now - t_prev)))
                t_prev = now
            if i + 1 < n:
                upload(i + 1)
            torch.cuda.current_stream().wait_event(ready[b])
            if q_split:
                dist.all_gather_into_tensor(bufs[b][1], q_part[b])     # the other ranks' slices arrive over NVLink
            elif world > 1:
                dist.broadcast(bufs[b][1], src=0)
            if world == 1:
                cur = ops.RankEvaluation(ops.GalleryIndex(bufs[b][0], SW), ops.QueryBatch(bufs[b][1]), true_idx=true_idx, topk=TOPK)
                free[b].record()
                if pending is not None:                # step i-1: finish it, send its result to the host
                    pb, pev = pending
                    if host_out[pb] is not None:
                        landed[pb].synchronize()       # the host buffers of step i-3 have been read below; reuse them
                    download(pb, pev.result())
                pending = (b, cur)
                if i >= 2:
                    landed[b].synchronize()            # step i-2 is on the host now
                    out = tuple(h.clone() for h in host_out[b])
            else:
                if i >= 2:
                    landed[b].synchronize()
                res = evaluate_ranks_sharded(bufs[b][0], bufs[b][1], g_offset, g_total, true_idx=true_idx, topk=TOPK)
                free[b].record()
                download(b, res)                       # device -> host read of this step's result (ranks, top-k)
                if i >= 1:
                    landed[b ^ 1].synchronize()        # step i-1 is on the host now
                    out = tuple(h.clone() for h in host_out[b ^ 1])
        if pending is not None:
            pb, pev = pending
            if host_out[pb] is not None:
                landed[pb].synchronize()
            download(pb, pev.result())
        last = (n - 1) % 2
        landed[last].synchronize()
        out = tuple(h.clone() for h in host_out[last])
        return out

    run_e2e(max(4, warmup))                             # untimed: fills the pipeline once (allocator pools, pinned buffers, all code paths)
    e2e_ms, e2e_out = timed(lambda: run_e2e(steps))
    clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions (device-resident and end-to-end)
    e2e_value = world * Q_TOTAL / (e2e_ms / steps / 1000.0)
    # the same step's upload on an idle device: the PCIe floor under the end-to-end step time
    torch.cuda.synchronize()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    for _ in range(3):
        bufs[0][0].copy_(ov_host, non_blocking=True)
        if q_split:
            q_part[0].copy_(su_host[q_lo:q_hi], non_blocking=True)
        elif rank == 0:
            bufs[0][1].copy_(su_host, non_blocking=True)
    u1.record()
    torch.cuda.synchronize()
    h2d_alone_ms = u0.elapsed_time(u1) / 3
    h2d = ov_host.numel() * 4 + (q_hi - q_lo) * su_host[0].numel() * 4     # per rank: its gallery shard + its slice of the queries
    d2h = sum(t.numel() * t.element_size() for t in e2e_out)
    del bufs

    extras = {}
    if world > 1 and not args.no_extras:
        extras.update(multi_gpu_extras(torch, dist, W, ops, device, world, rank))

    if world > 1:
        from witw_b200 import peer
        peer_used = any(v is not None for v in peer._cache.values())
        torch.cuda.synchronize()
        dist.barrier()
        peer.shutdown()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    hbm, burst, sustained, peak_src = measured_peaks()
    sweep_impl = ops._pick_impl(None, 64, 64, SW)
    kernel = "match_spec_kernel" if sweep_impl == "spectral" else "match_tc_kernel"
    pairs = float(G_PER_GPU) * float(Q_TOTAL)
    dense_equiv = FLOP_PER_PAIR * pairs / (kernel_ms / 1000.0) / 1e12
    traffic = profile_json("roofline_traffic.json")
    at_baseline = FOV == 360 and G_PER_GPU == 10000 and Q_TOTAL == 10000
    if sweep_impl == "spectral":
        # The spectral sweep is a three-stage pipeline per tile -- TMA: fp16 operands L2 -> shared memory; tcgen05.mma: 64
        # accumulators per pair in TMEM; epilogue: inverse FFT + maximum + rank count + top-k on the CUDA cores -- and TMEM holds
        # exactly one tile, so the first half of a tile's epilogue cannot overlap the next tile's MMAs.  Its roof is the slowest
        # stage running alone, measured on the kernel itself (tools/ring_roof.py: the hooks build switches the other stages off):
        # the operand ring, bound by the L2 -> SM delivery of 640 B per pair.  HBM and the tensor pipe are far from their peaks
        # by construction (operands are L2-resident; the correlation theorem removed 97 % of the tensor work).
        roof = profile_json("sweep_roof_r2.json")
        stage = roof.get("kernel_ms", {})
        ring_alone_ms = stage.get("ring_alone", 3.235)   # fallbacks: the round-2 measurement
        sec = kernel_ms / 1000.0
        achieved = SPEC_TMA_BYTES_PER_PAIR * pairs / sec / 1e9
        peak = SPEC_TMA_BYTES_PER_PAIR * 1e8 / (ring_alone_ms / 1000.0) / 1e9
        exec_tf = SPEC_TC_FLOP_PER_PAIR * pairs / sec / 1e12
        roofline = {
            "bound": "l2_to_sm_operand_delivery", "kernel": kernel, "kernel_ms": kernel_ms, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak,
            "peak_source": roof.get("source", "round-2 measurement quoted in DESIGN.md 4.2s (profiles/sweep_roof_r2.json not found)"),
            "bound_detail": "%d B of fp16 operands per (query, item) pair delivered from L2 into the shared-memory ring by TMA (a CTA pair stages "
                            "64 queries + 8 items per CTA and frequency slot for 1 024 pairs per CTA; TMEM capacity fixes the tile); peak = the same "
                            "bytes over the time of the kernel's own TMA + MMA ring running alone at 10k x 10k" % int(SPEC_TMA_BYTES_PER_PAIR),
            "stages_alone_ms": {k: stage.get(k) for k in ("ring_alone", "tma_alone", "mma_alone", "epilogue_alone", "full")},
            "stages_note": "each stage of the kernel's pipeline timed alone on the 10k x 10k workload (hooks build); `full` is the hooks build's "
                           "whole kernel.  The gap between the slowest stage and the kernel is the half of the epilogue that holds TMEM.",
            "traffic": traffic.get(kernel + "_10k_x_10k_fov360_dram_bytes") if at_baseline else None,
            "traffic_source": traffic.get(kernel + "_source", traffic.get("source")) if at_baseline else None,
            "hbm": {"algorithmic_bytes": (G_PER_GPU * 2 + Q_TOTAL) * 64 * 64 * 2,
                    "achieved": (G_PER_GPU * 2 + Q_TOTAL) * 64 * 64 * 2 / sec / 1e9, "peak": hbm, "unit": "GB/s",
                    "frac": (G_PER_GPU * 2 + Q_TOTAL) * 64 * 64 * 2 / sec / 1e9 / hbm,
                    "note": "the fp16 operands (16 KB per gallery item, 8 KB per query) need to come from HBM once; the sweep re-reads them from "
                            "L2 (`traffic` = measured DRAM bytes per launch).  HBM does not bound this kernel"},
            "tensor": {"executed_tflops": exec_tf, "peak": sustained, "frac": exec_tf / sustained, "flop_per_pair_executed": int(SPEC_TC_FLOP_PER_PAIR),
                       "note": "tensor pipe as executed; the roofline-bound tensor kernel of the path is `dense_sweep`"},
            "algorithmic_speedup": {"flop_per_pair_direct": FLOP_PER_PAIR, "flop_per_pair_executed": int(SPEC_TC_FLOP_PER_PAIR),
                                    "dense_equivalent_tflops": dense_equiv,
                                    "note": "correlation theorem: the same correlation from %.0fx fewer tensor FLOPs plus a 64-point inverse FFT "
                                            "per pair on the CUDA cores; not a roofline fraction" % (FLOP_PER_PAIR / SPEC_TC_FLOP_PER_PAIR)},
        }
    else:
        roofline = {"bound": "tensor", "kernel": kernel, "kernel_ms": kernel_ms, "achieved": dense_equiv, "peak": sustained, "unit": "TFLOP/s",
                    "frac": dense_equiv / sustained, "peak_source": peak_src + " bf16_tflops_sustained (kernel timed inside a long step)",
                    "traffic": traffic.get(kernel + "_10k_x_10k_fov360_dram_bytes") if at_baseline else None,
                    "traffic_source": traffic.get(kernel + "_source", traffic.get("source")) if at_baseline else None, "flop_per_pair": FLOP_PER_PAIR}
    detail = {}
    detail.update({
        "unit_note": "one unit = one query swept over one %d-item gallery shard; queries/s over the whole %d-item gallery = %.1f, pairs/s = %.4g"
                     % (G_PER_GPU, g_total, value / world, value * G_PER_GPU),
        "sweep": sweep_impl,
        "step": "fp32 features in HBM -> operand prep (fp16 %s of the norm-scaled features, scale / error-bound tables, fp32 spectra) -> fp32 "
                "true-match distances -> tcgen05 sweep (argmax, distance, rank count, top-k candidates; decisions inside the fp16 error bound "
                "deferred) -> top-k merge -> fp32 finish (deferred rank decisions, top-k re-rank, completeness proof)%s"
                % ("azimuth spectra in UMMA layout" if sweep_impl == "spectral" else "Hankel blocks",
                   "" if world == 1 else ("; gallery sharded, one shard per GPU; thresholds, counts and top-k exchanged over peer memory (NVLink stores "
                                          "from the library's kernels + flags)" if peer_used else
                                          "; gallery sharded, one shard per GPU, NCCL all-reduce of the thresholds + one all-gather of counts and top-k")),
        "pipelining": "one step deep: step i+1 is enqueued before the host reads step i's 4-byte finish flag",
        "l2_window": bool(args.l2_window),
        "spin_up": {"steps": n_spin, "ms_per_step": spin_ms, "rest_s": float(os.environ.get("WITW_BENCH_REST", "1.0")),
                    "note": "untimed: ~1 s of continuous steps (its wall-clock rate is the sustained, power-capped figure), then ~1 s of rest, "
                            "then the W warm-up steps and the timed steps"},
    })
    line = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": workload_config(world),
        "config_detail": detail,
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / steps,
                "h2d_alone_ms": h2d_alone_ms, "h2d_alone_gbs": h2d / h2d_alone_ms / 1e6, "h2d_bytes_per_step_all_ranks": h2d * world,
                "note": "gallery and queries uploaded from pinned host memory every step, W.GalleryIndex / QueryBatch / RankEvaluation on them; H2D of step "
                        "i+1 double-buffered on a copy stream behind step i, D2H of step i's ranks and top-k into pinned host memory on a third stream"
                        + ("" if world == 1 else "; every rank uploads its gallery shard and 1/N of the replicated query set (h2d_bytes_per_step is per rank), "
                           "the query slices are all-gathered over NCCL")},
        # per step: gallery prep, query prep, spectral_pairs (true match), the sweep, topk_merge, finish_scan / _pairs / _topk
        # (hankel: item_stats + gallery_blocks instead of one gallery prep, + spectral_rows x2)
        "gpu_launches": (8 if sweep_impl == "spectral" else 11) * steps,
        "clocks": clocks,
        "recall": {k: float(v) for k, v in recall.items()},
        "deferral": {"deferred_pairs": int(step_stats["deferred"].sum()), "queries_redone_in_fp32": int(step_stats["flagged"]),
                     "list_capacity_per_query": int(step_stats["list_cap"])} if step_stats else None,
    }
    line.update(extras)
    if world == 1:
        # the CPU leg runs on the very inputs of the timed step, so its ranks double as the parity check of the bench line
        parity, qps, cores = parity_against_cpu(torch, ranks, ov, su, 15.0, true_idx=true_idx)
        line["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                                "sample": "first %d queries of the %d-query set against the full %d-item gallery (%.1f s)"
                                          % (parity["queries"], Q_TOTAL, G_PER_GPU, parity["cpu_seconds"])}
        line["parity_check"] = parity
    if world == 1 and not args.no_extras:
        line["hard"] = safe(hard_arm, torch, ops, device, 8.0)
        line["dense_sweep"] = safe(dense_sweep_roofline, torch, ops, ov, su, burst, sustained, true_idx)
        line["e2e_resident"] = safe(resident_e2e, torch, W, ops, ov, su_host, true_idx, steps)
        line["dropin_loop"] = safe(dropin_loop, torch, W, ov, su)
        del ov_host
        line["configs"] = {
            "configs[2] 90deg 10k x 10k": safe(config_step, torch, ops, device, 10000, 90),
            "configs[4] 90deg 100k gallery (sweep half)": safe(config_step, torch, ops, device, 100000, 90, iters=3),
            "gallery 1k": safe(config_step, torch, ops, device, 1000, 360, iters=20, warm=5),
            "gallery 100k": safe(config_step, torch, ops, device, 100000, 360, iters=3),
            "configs[3] 1M gallery on one GPU": safe(config_step, torch, ops, device, 1000000, 360, iters=2, warm=1),
        }
        line["hbm_kernels"] = safe(hbm_kernels, torch, W, hbm)
        line["configs"]["configs[4] streamed polar transform, 100k x 5 channels"] = safe(streamed_polar, torch, W, hbm)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def multi_gpu_extras(torch, dist, W, ops, device, world, rank):
    """N > 1 side measurements (every rank takes part, rank 0 reports): BASELINE configs[3] -- the 1M-tile gallery sharded
    over the N GPUs (strong scaling) -- and the cost of the exchange alone."""
    from witw_b200.sharded import CudaLocal, evaluate_ranks_sharded

    out = {}
    try:
        g_total = 1000000
        lo, hi = W.shard_bounds(g_total, world, rank)
        true_idx = torch.arange(Q_TOTAL, device=device) * (g_total // Q_TOTAL)      # matches spread over all shards
        gen = torch.Generator(device=device).manual_seed(500 + rank)
        ov = torch.randn(hi - lo, 16, 4, 64, generator=gen, device=device) * 0.06
        su = torch.randn(Q_TOTAL, 16, 4, SW, generator=torch.Generator(device=device).manual_seed(77), device=device) * 0.06
        mine = (true_idx >= lo) & (true_idx < hi)
        # the owner of a query's match plants it; the others' contributions are zero in the sum
        shifts = torch.randint(0, 64, (Q_TOTAL,), generator=torch.Generator(device=device).manual_seed(78), device=device)
        cols = (shifts.view(-1, 1) + torch.arange(SW, device=device).view(1, SW)) % 64
        planted = torch.gather(ov[(true_idx - lo).clamp(0, hi - lo - 1)], 3, cols.view(Q_TOTAL, 1, 1, SW).expand(Q_TOTAL, 16, 4, SW))
        planted = torch.where(mine.view(-1, 1, 1, 1), planted, torch.zeros_like(planted))
        dist.all_reduce(planted)
        su = planted + NOISE * su
        sink = []
        local = CudaLocal(event_sink=sink)

        def step():
            return evaluate_ranks_sharded(ov, su, lo, g_total, true_idx=true_idx, topk=TOPK, local=local)

        step()
        dist.barrier()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        t0.record()
        for _ in range(n):
            res = step()
        t1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([t0.elapsed_time(t1) / n], device=device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        kms = statistics.mean(a.elapsed_time(b) for a, b in sink[-n:])
        rec = W.recall_from_ranks(res[0])
        out["configs[3] 1M gallery sharded"] = {
            "gallery_total": g_total, "gallery_per_gpu": hi - lo, "queries": Q_TOTAL, "scaling": "strong", "ms_per_step": float(ms.item()),
            "queries_per_s": Q_TOTAL / (float(ms.item()) / 1000.0), "pairs_per_s": float(g_total) * Q_TOTAL / (float(ms.item()) / 1000.0),
            "sweep_kernel_ms_rank0": kms, "recall_top_one": float(rec["top_one"])}
        del ov, su, planted
    except Exception as exc:
        out["configs[3] 1M gallery sharded"] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    torch.cuda.empty_cache()
    try:
        # the exchange alone: thresholds all-reduce + packed counts / top-k all-gather + merge, on buffers of the step's sizes
        from witw_b200 import sharded
        counts = torch.zeros(Q_TOTAL, dtype=torch.int64, device=device)
        td = torch.rand(Q_TOTAL, TOPK, device=device)
        ti = torch.zeros(Q_TOTAL, TOPK, dtype=torch.int32, device=device)
        d_true = torch.zeros(Q_TOTAL, device=device)
        local = CudaLocal()
        res = {}
        for name, packed in (("packed", True), ("separate", False)):
            def exch():
                dist.all_reduce(d_true)
                if packed:
                    c, ad, ai, _ = sharded._exchange_packed(counts, td, ti, world, None)
                else:
                    c = counts.clone()
                    dist.all_reduce(c)
                    ad = [torch.empty_like(td) for _ in range(world)]
                    ai = [torch.empty_like(ti) for _ in range(world)]
                    dist.all_gather(ad, td)
                    dist.all_gather(ai, ti)
                    ad, ai = torch.stack(ad), torch.stack(ai)
                local.merge(ad, ai, TOPK)
            for _ in range(3):
                exch()
            dist.barrier()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(20):
                exch()
            t1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([t0.elapsed_time(t1) / 20], device=device)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            res[name + "_ms"] = float(ms.item())
        # the same exchange over peer memory (witw_b200/peer.py): two kernels of the library, NVLink stores + flags
        from witw_b200 import peer
        px = peer.get(Q_TOTAL, TOPK, device=device) if sharded.PEER_EXCHANGE else None
        if px is not None:
            c32 = counts.to(torch.int32)
            t_idx = torch.arange(Q_TOTAL, device=device) % (world * G_PER_GPU)
            flagged = torch.zeros(1, dtype=torch.int32, device=device)

            def exch_peer():
                px.thresholds(d_true, t_idx, rank * G_PER_GPU, G_PER_GPU)
                px.results(c32, td, ti, flagged)
            for _ in range(3):
                exch_peer()
            dist.barrier()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(20):
                exch_peer()
            t1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([t0.elapsed_time(t1) / 20], device=device)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            res["peer_ms"] = float(ms.item())
        res["used"] = "peer" if px is not None else "packed"
        res["note"] = ("per step: [Q] thresholds to every rank + exchange of [Q] counts and [Q,%d] top-k + merge; peer = NVLink stores from the library's "
                       "own kernels + flags (default), packed = NCCL all-reduce + one all-gather, separate = all-reduce + all-reduce + two all-gathers" % TOPK)
        out["exchange"] = res
    except Exception as exc:
        out["exchange"] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    return out


def main():
    global G_PER_GPU, FOV, Q_TOTAL, SW, FLOP_PER_PAIR, NOISE
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--sweep", choices=["auto", "spectral", "hankel"], default="auto",
                    help="tensor-core sweep: spectral (correlation theorem, default where supported) or hankel (dense contraction over the shifts)")
    ap.add_argument("--gallery-per-gpu", type=int, default=G_PER_GPU,
                    help="gallery items per GPU (default 10000 = BASELINE configs[1]; 125000 on 8 GPUs = configs[3], the 1M-tile gallery)")
    ap.add_argument("--fov", type=int, default=FOV, help="field of view of the queries in degrees (default 360; 90 = BASELINE configs[2] / [4])")
    ap.add_argument("--queries", type=int, default=Q_TOTAL, help="number of queries (default 10000)")
    ap.add_argument("--noise", type=float, default=NOISE, help="noise of the planted matches (default 0.5: every match is rank 1; see the `hard` key)")
    ap.add_argument("--no-peer-exchange", action="store_true", help="N > 1: exchange over NCCL collectives instead of peer memory")
    ap.add_argument("--l2-window", action="store_true",
                    help="experiment: an L2 access-policy window over the query operand while the sweep runs (ops.L2_WINDOW)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the side measurements after the timed regions: what the ncu launch list of the step is taken with")
    args = ap.parse_args()
    G_PER_GPU = args.gallery_per_gpu
    FOV, Q_TOTAL, NOISE = args.fov, args.queries, args.noise
    SW = int(FOV / 360 * 512) // 8
    if not 1 <= SW <= 64:
        raise SystemExit("bench.py: --fov must give 1..64 feature columns, got %d" % SW)
    FLOP_PER_PAIR = 2 * 64 * 16 * 4 * SW
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
