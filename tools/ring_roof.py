#!/usr/bin/env python
"""The measured roofs of the spectral sweep (csrc/match_spec.cu): each stage of the kernel's own pipeline running alone.

The kernel is a three-stage pipeline per tile (1 024 pairs per CTA): TMA brings the fp16 operands from L2 into the
shared-memory ring, tcgen05.mma turns them into 64 accumulators per pair in TMEM, and the epilogue warps turn those into
distances (inverse FFT, maximum, rank count, top-k) on the CUDA cores.  This tool runs that same kernel at 10k x 10k with parts
of it switched off, through the WITW_SPEC_DEBUG switches of the hooks build (make -C witw_b200/csrc HOOKS=1 ->
libwitw_b200_hooks.so; the switches give wrong results by design and do not exist in the shipped library):

    full            everything                                                          the kernel's time
    ring_alone      TMA + MMAs, no epilogue, no per-tile TMEM hand-off (bits 0|1|7)      the operand ring's roof: what bounds it is
                                                                                        the L2 -> SM delivery of the operands
    tma_alone       TMA only (bits 0|1|2|7)                                             L2 -> shared memory at the ring's own depth
    mma_alone       MMAs + gallery loads only (bits 0|1|3|7)                             the tensor pipe on N=16-per-CTA MMAs
    epilogue_alone  the epilogue and the per-tile hand-off, no ring (bit 5)              the CUDA-core roof
    ring            TMA + MMAs + per-tile hand-off with an (almost) empty epilogue (bits 0|1)
    protocol_only   barriers and 4 KB gallery loads only (bits 0|1|2|3)
    epilogue_no_ldtm / _no_ifft / _neither   the epilogue alone without its TMEM loads / its inverse FFT / both

One subprocess per mode (the switches are read once per process).  WITW_RING_VARIANT = 2 (CTA pairs, default) or 1.  Writes
gpurun_out/sweep_roof_v<variant>.json; profiles/sweep_roof_r2.json is a copy of the variant-2 file and is what bench.py's
`roofline.peak` quotes.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODES = (("full", 0), ("ring_alone", 3 + 128), ("tma_alone", 7 + 128), ("mma_alone", 11 + 128), ("epilogue_alone", 32), ("ring", 3),
         ("protocol_only", 15), ("epilogue_no_ldtm", 2 + 32), ("epilogue_no_ifft", 1 + 32), ("epilogue_neither", 3 + 32),
         ("full_half_b", 256), ("ring_alone_half_b", 3 + 128 + 256))
# operand bytes TMA delivers per (query, item) pair: a CTA pair stages 16 KB of query spectra + 4 KB of gallery spectra per slot
# and CTA for 1 024 pairs per CTA; one CTA per tile stages 32 + 4 KB
TMA_BYTES_PER_PAIR = {1: 32.0 * 36864 / 1024.0, 2: 32.0 * 20480 / 1024.0}
VARIANT = int(os.environ.get("WITW_RING_VARIANT", "2"))


def child():
    sys.path.insert(0, ROOT)
    import torch
    from witw_b200 import _lib
    _lib.LIB_PATH = os.environ.get("WITW_RING_LIB", os.path.join(ROOT, "witw_b200", "libwitw_b200_hooks.so"))
    from witw_b200 import ops
    _lib.call("witw_match_spec_variant", VARIANT)
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(0)
    n = 10000
    ov = torch.randn(n, 16, 4, 64, device=dev, generator=gen) * 0.06
    su = torch.randn(n, 16, 4, 64, device=dev, generator=gen) * 0.06
    gal, qry = ops.GalleryIndex(ov, 64, impl="spectral", keep_fp32=False), ops.QueryBatch(su, impl="spectral", keep_fp32=False)
    d_true = torch.full((n,), 1.0, device=dev)
    t32 = torch.arange(n, dtype=torch.int32, device=dev)
    cnt = torch.zeros(n, dtype=torch.int32, device=dev)
    times = []
    for i in range(8):
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ops.sweep_tc(gal, qry, d_true=d_true, true_idx=t32, rank_count=cnt, topk=16, events=ev)
        torch.cuda.synchronize()
        if i >= 3:
            times.append(ev[0].elapsed_time(ev[1]))
    print(json.dumps({"kernel_ms": sum(times) / len(times)}))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child()
    hooks = os.path.join(ROOT, "witw_b200", "libwitw_b200_hooks.so")
    if not os.path.isfile(hooks):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "witw_b200", "csrc"), "HOOKS=1", "-j", "8"], stdout=subprocess.DEVNULL)
    res = {}
    only = os.environ.get("WITW_RING_MODES")
    for name, bits in MODES:
        if only and name not in only.split(",") and name != "full":
            continue
        env = dict(os.environ, WITW_SPEC_DEBUG=str(bits))
        out = subprocess.check_output([sys.executable, os.path.abspath(__file__), "--child"], env=env, text=True)
        res[name] = json.loads(out.strip().splitlines()[-1])["kernel_ms"]
    pairs = 1e8
    bpp = TMA_BYTES_PER_PAIR[VARIANT]
    rec = {"workload": "10k x 10k, 360 deg, spectral sweep, counts + top-16", "variant": VARIANT, "kernel_ms": res,
           "tma_bytes_per_pair": bpp,
           "operand_delivery_tbs": {k: bpp * pairs / (v / 1000.0) / 1e12 for k, v in res.items() if k in ("full", "ring_alone", "tma_alone")},
           "source": "tools/ring_roof.py on B200: match_spec_kernel's own pipeline stages running alone (hooks build, WITW_SPEC_DEBUG); "
                     + ", ".join("%s %.3f ms" % (k, v) for k, v in res.items())}
    print(json.dumps(rec, indent=1))
    out_path = os.path.join(ROOT, "gpurun_out", "sweep_roof_v%d.json" % VARIANT)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()
