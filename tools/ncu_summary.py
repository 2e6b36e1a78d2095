#!/usr/bin/env python
"""Condenses ncu exports (tools/ncu_export.sh: <name>.raw.csv, <name>.source.csv) into the small
text/CSV summaries kept under profiles/.  usage: tools/ncu_summary.py gpurun_out/<name> [...] > profiles/x.md"""
import collections
import csv
import re
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def raw_summary(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for d in data:
        out.append(("kernel", d[hdr.index("Kernel Name")], ""))
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append((w, d[i], units[i]))
    return out


def source_summary(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    body = [r for r in rows[2:] if r and r[0] != "Kernel Name"]
    i_src, i_exec = hdr.index("Source"), hdr.index("Instructions Executed")
    ops = collections.Counter()
    for r in body:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[i_src])
        ops[m.group(2) if m else "?"] += int(r[i_exec])
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    stalls = collections.Counter()
    for r in body:
        for i in stall_cols:
            try:
                stalls[hdr[i]] += int(r[i])
            except ValueError:
                pass
    return ops, stalls, sum(int(r[i_exec]) for r in body)


def main():
    for base in sys.argv[1:]:
        print(f"## {base.split('/')[-1]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k, v, u in raw_summary(base + ".raw.csv"):
            print(f"| {k} | {v} | {u} |")
        ops, stalls, total = source_summary(base + ".source.csv")
        print(f"\nwarp-level instructions executed: {total}\n")
        print("| opcode | executed | share |\n|---|---|---|")
        for o, c in ops.most_common(16):
            print(f"| {o} | {c} | {100 * c / total:.1f}% |")
        st = sum(stalls.values())
        print("\n| warp stall (sampled) | samples | share |\n|---|---|---|")
        for k, c in stalls.most_common(8):
            print(f"| {k} | {c} | {100 * c / max(st, 1):.1f}% |")
        print()


if __name__ == "__main__":
    main()
