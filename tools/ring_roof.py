#!/usr/bin/env python
"""The measured roof of the spectral sweep's data flow (csrc/match_spec.cu): the shared-memory operand ring, written by TMA
and read by N=16 UMMAs.  Runs the kernel's own ring at 10k x 10k with the epilogue's work switched off, piece by piece,
through the WITW_SPEC_DEBUG switches of the hooks build (make -C witw_b200/csrc HOOKS=1 -> libwitw_b200_hooks.so; the
switches give wrong results by design and do not exist in the shipped library):

    full      everything                                              the shipped kernel's time
    ring      TMA writes + MMA reads, epilogue reduced to its barriers   the roof of the ring  (bits 0|1: no IFFT, no TMEM loads)
    tma_only  TMA writes only                                         (bits 0|1|2: no tcgen05.mma)
    mma_only  MMA reads + gallery loads only                          (bits 0|1|8: no query-stage loads)

One subprocess per mode (the switches are read once per process).  Writes profiles/smem_ring_roof.json, which bench.py's
`roofline.peak` quotes: bytes through the ring per SM per second in `ring` mode.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODES = (("full", 0), ("ring", 3), ("tma_only", 7), ("mma_only", 11), ("epilogue_only", 12), ("epilogue_alone", 32), ("protocol_only", 15), ("protocol_nocommit", 15 + 64), ("tma_nocommit", 7 + 64), ("ring_nohandshake", 3 + 128), ("protocol_nohandshake", 15 + 128))
# written by TMA + read by the MMAs, per (query, item) pair: one CTA per tile 36 + 36 KB per slot and 1 024 pairs; CTA pairs
# 20 KB written, 16 KB of queries + 8 KB of items (its own, read by both tensor cores) read per CTA, slot and 1 024 pairs
BYTES_PER_PAIR = {1: 2.0 * (256 * 4608) / 1024.0, 2: 32.0 * (20480 + 24576) / 1024.0}
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
        if only and name not in only.split(",") and name not in ("full", "ring"):
            continue
        env = dict(os.environ, WITW_SPEC_DEBUG=str(bits))
        out = subprocess.check_output([sys.executable, os.path.abspath(__file__), "--child"], env=env, text=True)
        res[name] = json.loads(out.strip().splitlines()[-1])["kernel_ms"]
    pairs = 1e8
    bpp = BYTES_PER_PAIR[VARIANT]
    gbs = {k: bpp * pairs / 148.0 / (v / 1000.0) / 1e9 for k, v in res.items()}
    rec = {"workload": "10k x 10k, 360 deg, spectral sweep, counts + top-16", "variant": VARIANT, "kernel_ms": res,
           "ring_gbs_per_sm_at_that_time": gbs, "gbs_per_sm": gbs["ring"],
           "bytes_per_pair": bpp,
           "source": "tools/ring_roof.py on B200: match_spec_kernel's own TMA-write + UMMA-read ring with the epilogue arithmetic and TMEM loads "
                     "switched off (hooks build, WITW_SPEC_DEBUG=3); full kernel %.3f ms, ring alone %.3f ms" % (res["full"], res["ring"])}
    print(json.dumps(rec, indent=1))
    out_path = os.path.join(ROOT, "gpurun_out", "smem_ring_roof_v%d.json" % VARIANT)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()
