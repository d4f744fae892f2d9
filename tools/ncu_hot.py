#!/usr/bin/env python
"""Hot spots of one kernel from an .ncu-rep: key raw metrics + the most stall-sampled SASS instructions.
  python tools/ncu_hot.py gpurun_out/x.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct",
        "smsp__average_warps_issue_stalled", "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct", "lts__throughput.avg.pct", "dram__throughput.avg.pct", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "smsp__inst_executed.sum ", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "sm__inst_executed_pipe_tmem", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active")
for r in rows[2:3]:
    print(r[hdr.index("Kernel Name")][:100])
    for h, u, v in zip(hdr, units, r):
        if any(k in (h + " ") for k in KEYS) and ".min" not in h and ".max" not in h:
            print(f"  {h:95s} {u:10s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = [(int(r[isamp]), r[isrc].strip(), int(r[iex]), i) for i, r in enumerate(rows[2:]) if len(r) > iex and r[isamp].isdigit()]
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for s, text, ex, i in sorted(data, reverse=True)[:top]:
    print(f"{s:7d} {100 * s / tot:5.1f}%  ex={ex:10d}  #{i:5d}  {text[:100]}")
