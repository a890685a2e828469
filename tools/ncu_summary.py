#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`): headline metrics of each
captured launch plus the instruction / stall-sample share of each region of the step
kernel.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum"]
MARKS_TILE = [("tile loads", "cooperative, coalesced loads"), ("phaseA load/integrate", "phase A: thread = (world, agent)"),
         ("passA", "pass A: everything reward()"), ("passB", "pass B: terms that look"),
         ("outputs/writeback", "reward / done / info outputs"), ("reset", "reset (auto-reset of finished"),
         ("mask+obs copy", "phase B: element-parallel"), ("node_obs", "node_obs: [lw][j][k][8]"),
         ("staged adj/edges", "small worlds: the masked distance"), ("direct adj/edges", "large worlds (or reset-only"),
         ("end", "argmax over the one-hot")]
MARKS_SMALL = [("copy helpers", "warp-cooperative slab -> global copies"), ("place_small", "placement of one world in registers"),
               ("prologue/acttab", "the kernel ------"), ("loads", "loads: the whole world into registers"),
               ("physics", "stage 2 + stage 1: action decode"), ("headings", "pre-reward velocities and heading"),
               ("distances", "all pairwise distances of the world"), ("agent loop (reward/info)", "stages 3 + 4: agents in index order"),
               ("reward/done/info out", "reward / done / info outputs"), ("state write-back", "state write-back ----"),
               ("reset", "reset (auto-reset of finished worlds, or MODE_RESET)"), ("obs+adj", "emission: obs + adj"),
               ("edges", "emission: edge lists"), ("node_obs", "emission: node features")]


MARKS_COOP = [("copy helpers", "Contiguous span of n elements"), ("prologue", "gmpe_step_coop_kernel(const"),
              ("loads", "loads: one agent per lane"), ("physics", "stage 2 + stage 1: action decode"),
              ("passA", "pass A: what reward()"), ("passB", "pass B: terms that look at other agents"),
              ("reward/done/info out", "reward / done / info outputs"), ("state write-back", "state write-back ----"),
              ("reset", "reset (auto-reset of finished worlds, or MODE_RESET)"), ("node_obs", "emission: node features"),
              ("obs+adj", "emission: obs rows + adjacency"), ("edges", "emission: edge lists")]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def num(x):
    try:
        return int(x)
    except (TypeError, ValueError):
        return 0


def main(rep, src_path):
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    H = rows[0]
    units = dict(zip(H, rows[1]))
    for r in rows[2:]:
        d = dict(zip(H, r))
        print("== launch", d.get("ID"), d.get("Kernel Name", "")[:60])
        for w in WANT:
            if w in d:
                print(f"  {w:62s} {d[w]:>16s} {units.get(w, '')}")
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                                            "--launch-count", "1"]))))
    src = open(src_path).read().split("\n")
    marks = []
    for name, needle in (MARKS_SMALL if 'small' in src_path else (MARKS_COOP if 'coop' in src_path else MARKS_TILE)):
        for i, l in enumerate(src):
            if needle in l:
                marks.append((name, i + 1))
                break
    cur, Hh, out = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            Hh = r
            continue
        if Hh and r[0]:
            try:
                ln = int(r[0])
            except ValueError:
                continue
            d = dict(zip(Hh[4:], r[4:]))
            out.append((cur, ln, num(d.get("Instructions Executed")), num(d.get("# Samples")), r[1][:100]))
    tot = sum(o[2] for o in out) or 1
    ts = sum(o[3] for o in out) or 1
    reg = {}
    for f, ln, inst, samp, _ in out:
        key = "(device.cuh / intrinsics)"
        if f == src_path.split("/")[-1]:
            key = "helpers(top)"
            for name, start in marks:
                if ln >= start:
                    key = name
        a = reg.setdefault(key, [0, 0])
        a[0] += inst
        a[1] += samp
    print(f"== per-region share (warp instructions {tot}, stall samples {ts})")
    for k, v in reg.items():
        print(f"  {k:28s} inst {100 * v[0] / tot:5.1f}%   samples {100 * v[1] / ts:5.1f}%")
    print("== top source lines by stall samples")
    for o in sorted(out, key=lambda o: -o[3])[:25]:
        print(f"  {o[0][:18]:18s}:{o[1]:4d} inst {100 * o[2] / tot:5.1f}% samp {100 * o[3] / ts:5.1f}%  {o[4]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "contracts-marl-aam-corridors_b200/csrc/gmpe_step.cuh")
