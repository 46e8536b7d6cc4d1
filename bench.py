#!/usr/bin/env python
"""Benchmark of the WITW retrieval hot path on B200 (contract: see the task description).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): cvig_fov.py 360-degree eval, 10k queries x 10k gallery,
orientation-searched distance + rank counting + top-10, synthetic feature maps [N,16,4,64].
One step = one full pass from fp32 feature maps resident in HBM to per-query ranks and top-k:
gallery/query operand prep -> exact fp32 true-match distances -> tcgen05 sweep (fused argmax,
crop-normalise, distance, rank count, top-k) -> fp32 re-check -> top-k merge -> fp32 re-rank.
--sweep spectral (default: what the library picks) evaluates the circular correlation through the
correlation theorem (csrc/match_spec.cu: per-frequency tcgen05 products + in-register inverse FFT);
--sweep hankel is the dense contraction over all 64 shifts (csrc/match_tc.cu).
N > 1: the gallery is sharded, one 10k-item shard per GPU (weak scaling: gallery_total = N*10k),
queries replicated; the exchange is an all-reduce of [Q] true distances and [Q] counts and an
all-gather of [Q,10] top-k candidates over NCCL.  value = N*Q / t: queries swept per second, each
against a 10k-item gallery shard.
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
SW = 64                                # query columns: int(fov/360*512)//8 (cvig_fov.py:22, 8 image pixels per feature column)
FLOP_PER_PAIR = 2 * 64 * 16 * 4 * SW   # 2*W*C*H*sw = 524 288 at 360 deg, 131 072 at 90 deg
METRIC = "queries/sec vs gallery size (orientation-searched distance + top-k)"


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
            deadline = time.time() + 5.0                 # wait until it has attached to the driver and delivers samples
            while not self.lines and time.time() < deadline and self.proc.poll() is None:
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


def measured_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.isfile(path):
        with open(path) as f:
            t = json.load(f)
        return t.get(kernel + "_10k_x_10k_fov360_dram_bytes"), t.get(kernel + "_source", t.get("source"))
    return None, None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return p.get("bf16_tflops_sustained", 1397.8), "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
    return 1400.0, "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"


def cpu_reference_queries_per_s(budget_s=15.0):
    """The reference's CPU path (oracle port: conv2d / argmax / gather / norm, cvig_fov.py:545-552) on the host
    cores: a bounded sample of queries against the full 10k gallery."""
    import torch

    from oracle import witw_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ov, su, _ = O.synth_features(G_PER_GPU, 64, fov=FOV, noise=0.5, seed=1234)
    t0 = time.perf_counter()
    O.rank_loop(ov, su, query_indices=[0])          # warm-up, also the calibration sample
    per_q = time.perf_counter() - t0
    n = int(max(2, min(64, budget_s / max(per_q, 1e-3))))
    t0 = time.perf_counter()
    O.rank_loop(ov, su, query_indices=list(range(n)))
    dt = time.perf_counter() - t0
    return n / dt, cores, "first %d queries of the %d-query set against the full %d-item gallery (%.1f s)" % (n, Q_TOTAL, G_PER_GPU, dt)


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # each step is a bounded sample; the whole run stays within a few minutes
    per_step_budget = max(2.0, min(15.0, 150.0 / (steps + warmup)))
    vals, sample, cores = [], "", 1
    for i in range(warmup + steps):
        v, cores, sample = cpu_reference_queries_per_s(per_step_budget)
        if i >= warmup:
            vals.append(v)
    value = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1000.0 * Q_TOTAL / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "cvig_fov %ddeg eval, %d queries x %d gallery, rank loop (BASELINE %s); CPU port of the reference's PyTorch path"
                               % (FOV, Q_TOTAL, G_PER_GPU, config_name(G_PER_GPU)),
                   "gallery": G_PER_GPU, "queries": Q_TOTAL, "fov": FOV},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def make_data(torch, device, n_gallery, n_query, seed, planted):
    """Synthetic feature maps of the BASELINE shapes generated on the device."""
    gen = torch.Generator(device=device).manual_seed(seed)
    ov = torch.randn(n_gallery, 16, 4, 64, generator=gen, device=device) * 0.06
    qgen = torch.Generator(device=device).manual_seed(1234)
    su = torch.randn(n_query, 16, 4, SW, generator=qgen, device=device) * 0.06
    if planted:
        n = min(n_gallery, n_query)
        shifts = torch.randint(0, 64, (n,), generator=qgen, device=device)
        cols = (shifts.view(n, 1) + torch.arange(SW, device=device).view(1, SW)) % 64
        su[:n] = torch.gather(ov[:n], 3, cols.view(n, 1, 1, SW).expand(n, 16, 4, SW)) + 0.5 * su[:n]
    return ov, su


def dense_sweep_roofline(torch, ops, ov, su, peak, true_idx=None, iters=5):
    """The north star's kernel (2): the shift search as one dense bf16 contraction on tcgen05 (csrc/match_tc.cu), timed alone
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
    return {"kernel": "match_tc_kernel", "bound": "tensor", "kernel_ms": ms, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "launches": iters,
            "note": "dense contraction over all 64 shifts, %d tensor FLOP per pair as executed; same operands and epilogue outputs "
                    "(rank count + top-k candidates) as the step's spectral sweep; ncu: profiles/match_tc_r1c.txt" % FLOP_PER_PAIR}


def gallery_size_sweep(torch, ops, device, value_10k, sizes=(1000, 100000), iters=5):
    """The metric's other axis: queries/s of the same device-resident step (10k queries, 360 deg) at other gallery sizes."""
    rows = []
    for g in sizes:
        try:
            ov, su = make_data(torch, device, g, Q_TOTAL, seed=7, planted=False)
            true_idx = torch.arange(Q_TOTAL, device=device) % g
            gen = torch.Generator(device=device).manual_seed(11)      # every query is a rolled, noised copy of its gallery item
            cols = (torch.randint(0, 64, (Q_TOTAL, 1), generator=gen, device=device) + torch.arange(SW, device=device).view(1, SW)) % 64
            su = torch.gather(ov[true_idx], 3, cols.view(Q_TOTAL, 1, 1, SW).expand(Q_TOTAL, 16, 4, SW)) + 0.5 * su

            def step():
                return ops.evaluate_ranks_prepared(ops.GalleryIndex(ov, SW), ops.QueryBatch(su), true_idx=true_idx, topk=TOPK)

            for _ in range(3):
                step()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(iters):
                step()
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / iters
            rows.append({"gallery": g, "queries": Q_TOTAL, "ms_per_step": ms, "queries_per_s": Q_TOTAL / (ms / 1000.0),
                         "pairs_per_s": float(g) * Q_TOTAL / (ms / 1000.0)})
            del ov, su
        except Exception as exc:      # a side measurement must not take the bench line down with it
            rows.append({"gallery": g, "error": "%s: %s" % (type(exc).__name__, exc)})
        torch.cuda.empty_cache()
    rows.append({"gallery": G_PER_GPU, "queries": Q_TOTAL, "queries_per_s": value_10k, "note": "the timed step above"})
    return sorted(rows, key=lambda r: r["gallery"])


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


def run_ours(args):
    import torch
    import torch.distributed as dist

    import witw_b200 as W
    from witw_b200 import ops
    from witw_b200.sharded import evaluate_ranks_sharded

    if args.sweep != "auto":
        ops.TC_IMPL = args.sweep
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

    ov, su = make_data(torch, device, G_PER_GPU, Q_TOTAL, seed=100 + rank, planted=(rank == 0))
    g_offset = rank * G_PER_GPU
    g_total = world * G_PER_GPU
    sweep_events = []

    # query i matches gallery item i; with more queries than gallery items the match indices wrap around
    true_idx = None if Q_TOTAL <= g_total else torch.arange(Q_TOTAL, device=device) % g_total

    def step_device():
        """fp32 features in HBM -> ranks (+ top-k)."""
        if world == 1:
            gallery = ops.GalleryIndex(ov, SW)
            queries = ops.QueryBatch(su)
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            sweep_events.append(ev)
            return ops.evaluate_ranks_prepared(gallery, queries, true_idx=true_idx, topk=TOPK, events=ev)
        return evaluate_ranks_sharded(ov, su, g_offset, g_total, true_idx=true_idx, topk=TOPK, local=timed_local)

    timed_local = W.sharded.CudaLocal(event_sink=sweep_events)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        out = None
        for _ in range(n):
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
    if rank == 0:
        sampler.start()
    for _ in range(warmup):
        out = step_device()
    torch.cuda.synchronize()
    sweep_events.clear()
    sampler.mark()
    ms_total, out = timed(step_device, steps)
    ms_step = ms_total / steps
    value = world * Q_TOTAL / (ms_step / 1000.0)
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in sweep_events[:steps])
    ranks = out[0]
    recall = W.recall_from_ranks(ranks)

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

    def upload(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[b])            # the step that last used this buffer has finished with it
            bufs[b][0].copy_(ov_host, non_blocking=True)
            if rank == 0:                              # the replicated query set crosses PCIe once, on rank 0 ...
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

    def run_e2e(n):
        out = None
        upload(0)
        for i in range(n):
            b = i % 2
            if i + 1 < n:
                upload(i + 1)
            torch.cuda.current_stream().wait_event(ready[b])
            if world > 1:
                dist.broadcast(bufs[b][1], src=0)      # ... and reaches the other ranks over NVLink
            if i >= 2:
                landed[b].synchronize()                # the host buffers of step i-2 have been read below; reuse them
            if world == 1:
                res = W.evaluate_ranks(bufs[b][0], bufs[b][1], true_idx=true_idx, path="tc", topk=TOPK)
            else:
                res = evaluate_ranks_sharded(bufs[b][0], bufs[b][1], g_offset, g_total, true_idx=true_idx, topk=TOPK)
            free[b].record()
            download(b, res)                           # device -> host read of this step's result (ranks, top-k)
            if i >= 1:
                landed[b ^ 1].synchronize()            # step i-1 is on the host now
                out = tuple(h.clone() for h in host_out[b ^ 1])
        landed[(n - 1) % 2].synchronize()
        out = tuple(h.clone() for h in host_out[(n - 1) % 2])
        return out

    run_e2e(2)
    e2e_ms, e2e_out = timed(lambda: run_e2e(steps), 1)
    clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions (device-resident and end-to-end)
    e2e_value = world * Q_TOTAL / (e2e_ms / steps / 1000.0)
    # the same step's upload on an idle device: the PCIe floor under the end-to-end step time
    torch.cuda.synchronize()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    for _ in range(3):
        bufs[0][0].copy_(ov_host, non_blocking=True)
        if rank == 0:
            bufs[0][1].copy_(su_host, non_blocking=True)
    u1.record()
    torch.cuda.synchronize()
    h2d_alone_ms = u0.elapsed_time(u1) / 3
    h2d = ov_host.numel() * 4 + su_host.numel() * 4     # rank 0; the other ranks upload their gallery shard only
    d2h = sum(t.numel() * t.element_size() for t in e2e_out)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peaks()
    achieved = FLOP_PER_PAIR * float(G_PER_GPU) * float(Q_TOTAL) / (kernel_ms / 1000.0) / 1e12
    sweep_impl = ops._pick_impl(None, 64, 64, SW)
    kernel = "match_spec_kernel" if sweep_impl == "spectral" else "match_tc_kernel"
    traffic, traffic_src = measured_traffic(kernel) if (FOV == 360 and G_PER_GPU == 10000 and Q_TOTAL == 10000) else (None, None)
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "kernel": kernel, "kernel_ms": kernel_ms, "flop_per_pair": FLOP_PER_PAIR, "peak_source": peak_src}
    if sweep_impl == "spectral":
        # the algorithmic count above is the direct form's; the spectral sweep executes 33 bins x 64 rows of
        # complex MACs per pair on the tensor cores and a 64-point inverse FFT per pair on the CUDA cores
        exec_tc = 2.0 * 128 * 16 * 16 * 256 / 1024.0          # FLOP per pair issued as tcgen05.mma (256 MMAs of 128x16x16 per 1024 pairs)
        smem_bytes = 2.0 * (256 * 4608) / 1024.0              # per pair: operand bytes written by TMA + read by the MMAs
        roofline.update({
            "note": "frac > 1: achieved counts the direct form's %d FLOP per pair; the kernel evaluates the same correlation "
                    "through the correlation theorem with %d tensor FLOP + ~1 000 CUDA-core FLOP per pair" % (FLOP_PER_PAIR, int(exec_tc)),
            "executed_tensor_tflops": exec_tc * float(G_PER_GPU) * float(Q_TOTAL) / (kernel_ms / 1000.0) / 1e12,
            "bound_detail": "shared-memory bandwidth feeding N=16 UMMAs (query stage written once by TMA, read once per 8 gallery items)",
            "smem_gbs_per_sm": smem_bytes * float(G_PER_GPU) * float(Q_TOTAL) / 148.0 / (kernel_ms / 1000.0) / 1e9,
        })
    line = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": "cvig_fov %ddeg eval: %d queries x %d-item gallery per GPU, orientation-searched distance + rank count + top-%d "
                        "(BASELINE %s%s)" % (FOV, Q_TOTAL, G_PER_GPU, TOPK, config_name(g_total),
                                             "" if world == 1 else "; gallery sharded, one shard per GPU, NCCL count all-reduce + top-k all-gather"),
            "gallery_total": g_total, "gallery_per_gpu": G_PER_GPU, "queries": Q_TOTAL, "fov": FOV, "feature_shape": [16, 4, 64],
            "unit_note": "one unit = one query swept over one %d-item gallery shard; queries/s over the whole %d-item gallery = %.1f" % (G_PER_GPU, g_total, value / world),
            "l2_policy": "inputs larger than L2 (fp32 features %d MB + their fp32 spectra + the bf16 operands per step vs 126 MB L2)"
                         % ((G_PER_GPU * 64 * 64 + Q_TOTAL * 64 * SW) * 4 // 1000000),
            "sweep": sweep_impl,
            "step": ("fp32 features in HBM -> operand prep (bf16 azimuth spectra in UMMA layout, norms, fp32 spectra) -> fp32 true-match distances -> "
                     "tcgen05 per-frequency products + in-register inverse FFT, argmax, distance, rank count, top-k -> fp32 re-check of near-threshold "
                     "rank decisions -> top-k merge -> fp32 re-rank of the top-k") if sweep_impl == "spectral" else
                    ("fp32 features in HBM -> operand prep (bf16 Hankel blocks, norms, fp32 azimuth spectra) -> fp32 true-match distances -> tcgen05 sweep -> "
                     "fp32 re-check of near-threshold rank decisions -> top-k merge -> fp32 re-rank of the top-k"),
        },
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / steps,
                "h2d_alone_ms": h2d_alone_ms, "h2d_alone_gbs": h2d / h2d_alone_ms / 1e6,
                "note": "W.evaluate_ranks() on pinned-host inputs; H2D of step i+1 double-buffered on a copy stream behind step i, "
                        "D2H of step i's ranks and top-k into pinned host memory on a third stream, read by the host during step i+1"
                        + ("" if world == 1 else "; every rank uploads its gallery shard, rank 0 also the replicated query set, which is then broadcast over NCCL")},
        # per step: gallery prep, crop_norm, query prep, (hankel: spectral_rows x2,) spectral_pairs (true match), the sweep,
        # topk_merge, spectral_pairs (re-check), recheck_apply, topk_refine_pairs, spectral_pairs (top-k), topk_refine_sort
        "gpu_launches": (11 if sweep_impl == "spectral" else 14) * steps,
        "clocks": clocks,
        "recall": {k: float(v) for k, v in recall.items()},
    }
    if world == 1 and not args.no_extras:
        line["dense_sweep"] = dense_sweep_roofline(torch, ops, ov, su, peak, true_idx)
        line["gallery_sweep"] = gallery_size_sweep(torch, ops, device, value)
    if world == 1:
        v, cores, sample = cpu_reference_queries_per_s(15.0)
        line["cpu_baseline"] = {"value": v, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    global G_PER_GPU, FOV, Q_TOTAL, SW, FLOP_PER_PAIR
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
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the side measurements after the timed regions (dense-sweep roofline, gallery-size sweep): "
                         "what the ncu launch list of the step is taken with")
    args = ap.parse_args()
    G_PER_GPU = args.gallery_per_gpu
    FOV, Q_TOTAL = args.fov, args.queries
    SW = int(FOV / 360 * 512) // 8
    if not 1 <= SW <= 64:
        raise SystemExit("bench.py: --fov must give 1..64 feature columns, got %d" % SW)
    FLOP_PER_PAIR = 2 * 64 * 16 * 4 * SW
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
