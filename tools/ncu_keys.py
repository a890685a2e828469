#!/usr/bin/env python
"""Headline metrics + stall-reason sample counts of the first launch in an .ncu-rep.
Usage: python tools/ncu_keys.py gpurun_out/prof.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum"]


def main(rep, idx=0):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U, d = rows[0], rows[1], dict(zip(rows[0], rows[2 + idx]))
    print("kernel:", d.get("Kernel Name"))
    for k in KEYS:
        if k in d:
            print(f"  {k:66s} {d[k]:>16s} {U[H.index(k)]}")
    st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(d[k]) for k in H
          if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and d[k]}
    tot = sum(st.values()) or 1
    print("  stall samples:", ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v * 50 > tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
