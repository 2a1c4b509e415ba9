#!/usr/bin/env python
"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into profiles/<round>_<name>.md and
profiles/traffic.json (DRAM bytes per launch, keyed by the bench's kernel names).

usage: tools/ncu_summary.py r01 k_lq=gpurun_out/x.ncu-rep k_bwd=... ls_rollout=... ls_merit=...
"""
import csv
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" , "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "local_load_bytes", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]

SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, zip(units, r))) for r in rows[2:]]


def main():
    rnd = sys.argv[1]
    traffic = {}
    tpath = os.path.join("profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    for arg in sys.argv[2:]:
        name, rep = arg.split("=")
        for idx, row in enumerate(read(rep)):
            kname = row["Kernel Name"][1]
            lines = [f"# {rnd} {name}: `{kname.split('(')[0]}`", "",
                     f"source report: `{rep}` (ncu --set full --clock-control none, one launch, bench workload batch 4096)", "",
                     "| metric | value | unit |", "|---|---|---|"]
            for k in KEYS:
                if k in row:
                    lines.append(f"| {k} | {row[k][1]} | {row[k][0]} |")
            rd = float(row["dram__bytes_read.sum"][1]) * SCALE[row["dram__bytes_read.sum"][0]]
            wr = float(row["dram__bytes_write.sum"][1]) * SCALE[row["dram__bytes_write.sum"][0]]
            traffic[name] = rd + wr
            lines += ["", f"DRAM traffic per launch: {rd/1e9:.3f} GB read + {wr/1e9:.3f} GB written = {(rd+wr)/1e9:.3f} GB"]
            suffix = "" if idx == 0 else f"_{idx}"
            with open(os.path.join("profiles", f"{rnd}_{name}{suffix}.md"), "w") as f:
                f.write("\n".join(lines) + "\n")
            break
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
    print(traffic)


if __name__ == "__main__":
    main()
