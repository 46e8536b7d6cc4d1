#!/usr/bin/env python
"""Turn an .ncu-rep (--set full --import-source on; one or several kernels) into a short text summary for profiles/.
usage: python tools/summarize_ncu.py gpurun_out/x.ncu-rep [kernel-name-substring] > profiles/x.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
only = sys.argv[2] if len(sys.argv) > 2 else None
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "smsp__inst_executed_op_shared_ld.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
print("# ncu summary of", rep)
ik = h.index("Kernel Name")
seen = set()
for v in rows[2:]:
    name = v[ik]
    if (only and only not in name) or name in seen:
        continue
    seen.add(name)
    print("\n## kernel:", name[:120])
    for col, unit, val in zip(h, u, v):
        if col in KEYS:
            print("%-86s %-16s %s" % (col, unit, val))
if only or len(seen) == 1:
    args = ["ncu", "-i", rep, "--page", "source", "--csv"] + (["-k", "regex:" + only] if only else [])
    src = subprocess.run(args, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    try:
        hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
        body = rows[rows.index(hdr) + 1:]
        ist, isrc, iex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
        data = [(int(r[ist]) if r[ist].isdigit() else 0, r[isrc].strip(), r[iex]) for r in body if len(r) > max(ist, isrc, iex)]
        tot = sum(d[0] for d in data) or 1
        print("\n# top warp-stall sample locations (SASS), %d samples" % tot)
        for d in sorted(data, reverse=True)[:16]:
            print("%6.2f%%  exec=%-11s %s" % (100.0 * d[0] / tot, d[2], d[1][:110]))
        tens = [d for d in data if any(k in d[1] for k in ("UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM"))]
        if tens:
            print("\n# tensor / TMA instructions")
            for d in tens[:48]:
                print("%6.2f%%  exec=%-11s %s" % (100.0 * d[0] / tot, d[2], d[1][:110]))
    except StopIteration:
        pass
