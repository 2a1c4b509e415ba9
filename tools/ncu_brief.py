#!/usr/bin/env python
"""Brief of one ncu report: headline metrics, stall reasons, dynamic opcode mix (per unit if given).
usage: tools/ncu_brief.py report.ncu-rep [units]   (units = e.g. games x steps, to normalise counts)"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h, u, r = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_local_ld.sum',
        'smsp__inst_executed_op_local_st.sum', 'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
for k in keys:
    if k in h: print(f"{k:75s} {r[h.index(k)]:>16s} {u[h.index(k)]}")
st = [(float(r[i]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
      for i, k in enumerate(h) if 'issue_stalled' in k and 'per_issue_active' in k]
print("stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); hdr = rows[1]; data = rows[2:]
iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(int(x[iE]) for x in data); cls = collections.Counter()
for x in data:
    op = x[iS].split()
    if op[0].startswith('@'): op = op[1:]
    cls[op[0].split('.')[0]] += int(x[iE])
print(f"warp instructions {tot}" + (f" = {tot / units:.1f} per unit" if units else ""))
print("  ".join(f"{k} {v / (units or tot) * (1 if units else 100):.1f}" for k, v in cls.most_common(18)))
