#!/usr/bin/env python
"""Summarises an ncu report (.ncu-rep) of one kernel into a small text file for profiles/.

usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/NAME.txt ["free-form note"]
Reads the raw page (launch metrics) and the source page (SASS opcode mix + stall samples) with `ncu -i`.
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    lines = [f"# ncu summary of {rep}", note, ""]
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"## kernel {d.get('Kernel Name', '?')}  (launch id {d.get('ID', '?')})")
        for k in KEYS:
            if k in d and d[k] != "":
                lines.append(f"{k:75s} {d[k]:>18s} {u[k]}")
        st = [(h, d[h]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
        lines.append("stall reasons (warps stalled per issue-active cycle):")
        for h, v in sorted(st, key=lambda t: -float(t[1] or 0))[:9]:
            lines.append(f"    {h.split('stalled_')[1].replace('_per_issue_active.ratio', ''):28s} {float(v):8.3f}")
        lines.append("")
    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    if len(src) > 2:
        hdr = src[1]
        ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        ops, samp = collections.Counter(), collections.Counter()
        for r in src[2:]:
            if len(r) < len(hdr):
                continue
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia].strip())
            if not m:
                continue
            op = m.group(2)
            base = op.split(".")[0]
            if base in ("MUFU", "F2F"):
                base = ".".join(op.split(".")[:2])
            ops[base] += int(r[ie] or 0)
            samp[base] += int(r[isamp] or 0)
        tot, ts = sum(ops.values()), max(1, sum(samp.values()))
        lines.append(f"SASS opcode mix (warp-level instructions executed, total {tot}):")
        for k, v in ops.most_common(22):
            lines.append(f"    {k:14s} {v:15d} {100 * v / tot:6.2f}%   stall samples {100 * samp[k] / ts:6.2f}%")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
