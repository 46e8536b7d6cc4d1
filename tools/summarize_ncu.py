#!/usr/bin/env python
"""Turn an .ncu-rep (one kernel, --set full --import-source on) into a short text summary for profiles/.
usage: python tools/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "smsp__inst_executed_op_shared_ld.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
print("# ncu summary of", rep)
print("kernel:", v[h.index("Kernel Name")] if "Kernel Name" in h else "?")
for name, unit, val in zip(h, u, v):
    if name in KEYS or any(name.endswith("." + k) for k in KEYS):
        print("%-86s %-16s %s" % (name, unit, val))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ist, isrc, iex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
data = [(int(r[ist]) if r[ist].isdigit() else 0, r[isrc].strip(), r[iex]) for r in rows[2:] if len(r) > ist]
tot = sum(d[0] for d in data) or 1
print("\n# top warp-stall sample locations (SASS), %d samples" % tot)
for d in sorted(data, reverse=True)[:20]:
    print("%6.2f%%  exec=%-11s %s" % (100.0 * d[0] / tot, d[2], d[1][:110]))
print("\n# tensor / TMA instructions")
for d in data:
    if any(k in d[1] for k in ("UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM")):
        print("%6.2f%%  exec=%-11s %s" % (100.0 * d[0] / tot, d[2], d[1][:110]))
