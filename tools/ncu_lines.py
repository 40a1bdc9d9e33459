#!/usr/bin/env python
"""tools/ncu_lines.py REPORT.ncu-rep OUT.txt [min_percent] -- per CUDA source line: warp instructions executed, stall samples,
active lanes, from the report's source page (`--import-source on` capture, code built with -lineinfo).  Small enough to come
back from the GPU box when the report itself is not."""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    floor = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    i_inst, i_thr = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    i_smp = hdr.index("Warp Stall Sampling (All Samples)")
    cur, fname = None, "?"
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) < 8:
            if r and r[0] == "File Path":
                fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            continue
        if r[0] != "":
            try:
                cur = (fname, int(r[0]), r[1])
            except ValueError:
                continue
            agg.setdefault(cur, [0, 0, 0])
            continue
        if cur is None:
            continue
        try:
            agg[cur][0] += int(r[i_inst]); agg[cur][1] += int(r[i_smp]); agg[cur][2] += int(r[i_thr])
        except (ValueError, IndexError):
            pass
    tot = max(1, sum(v[0] for v in agg.values()))
    tots = max(1, sum(v[1] for v in agg.values()))
    lines = [f"# per-line profile of {rep}: {tot} warp instructions, {tots} stall samples; lines with >= {floor} % of either"]
    for k, v in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        if v[0] >= tot * floor / 100 or v[1] >= tots * floor / 100:
            lines.append(f"{k[0]:18s} {k[1]:5d} inst {100 * v[0] / tot:5.2f}% stall {100 * v[1] / tots:5.2f}% lanes {v[2] / max(v[0], 1):5.1f} | {k[2].strip()[:120]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out, len(lines))


if __name__ == "__main__":
    main()
