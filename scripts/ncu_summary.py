#!/usr/bin/env python3
"""Condense an `ncu --page raw --csv` (+ optional `--page source --csv`) export into a markdown summary.

usage: ncu_summary.py raw.csv [source.csv] > profiles/<name>.md
"""
import csv
import re
import sys
from collections import Counter

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__sass_average_branch_targets_threads_uniform.pct",
    "smsp__average_warp_latency_per_inst_issued.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed_op_branch.sum",
]
STALLS = "smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio"


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print(f"# ncu summary: kernel `{d.get('Kernel Name', ('?',))[0]}`\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in d:
            print(f"| {k} | {d[k][0]} | {d[k][1]} |")
    print("\n## Warp stall reasons (warps per issue-active cycle)\n\n| reason | value |\n|---|---|")
    st = []
    for h in hdr:
        m = re.fullmatch(STALLS, h)
        if m:
            st.append((float(d[h][0].replace(",", "")), m.group(1)))
    for v, name in sorted(st, reverse=True)[:10]:
        print(f"| {name} | {v:.3f} |")
    if len(sys.argv) > 2:
        src = list(csv.reader(open(sys.argv[2])))[2:]
        tot = sum(int(r[5]) for r in src)
        ops, thr = Counter(), Counter()
        for r in src:
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1].strip())
            op = m.group(2).split(".")[0]
            ops[op] += int(r[5])
            thr[op] += int(r[6])
        print(f"\n## Executed warp instructions by opcode (total {tot:.4g})\n\n| opcode | share | avg active threads |\n|---|---|---|")
        for op, n in ops.most_common(16):
            print(f"| {op} | {n / tot * 100:.2f} % | {thr[op] / max(n, 1):.1f} |")
        print("\n## Top instructions by stall samples\n\n| samples | executed | avg threads | SASS |\n|---|---|---|---|")
        for r in sorted(src, key=lambda r: -int(r[4]))[:14]:
            print(f"| {r[4]} | {r[5]} | {r[8]} | `{r[1].strip()}` |")


if __name__ == "__main__":
    main()
