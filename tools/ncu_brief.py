#!/usr/bin/env python
"""Key numbers of the first kernel in an ncu report (raw page): duration, issue, pipes, shared-memory
wavefronts and bank conflicts, DRAM bytes, stall reasons, per-opcode executed counts.
usage: ncu_brief.py report.ncu-rep [codes]   (codes: per-code normalisation of the instruction mix)"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
codes = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in want:
    if k in m:
        print(f"{k:90s} {m[k]}")
for k, v in m.items():
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
        try:
            if float(v) >= 0.15:
                print(f"  stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {float(v):.2f}")
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
mix = collections.Counter()
n = 0
for r in csv.reader(src.splitlines()):
    if r and r[0] == "Kernel Name":
        n += 1
        if n > 1:
            break
        continue
    if len(r) < 7 or r[0] == "Address":
        continue
    try:
        ex = float(r[5])
    except ValueError:
        continue
    toks = r[1].strip().split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    mix[op] += ex
tot = sum(mix.values())
div = codes / 32 if codes else tot / 100
print(f"warp-instructions executed {tot:.0f}" + (f" = {tot / div:.2f} per code" if codes else ""))
for op, c in mix.most_common(22):
    print(f"  {op:34s} {c / div:8.2f}" + ("" if codes else " %"))
