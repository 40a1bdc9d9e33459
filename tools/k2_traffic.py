#!/usr/bin/env python
"""Writes profiles/k2_traffic.json from an ncu summary (tools/ncu_summary.py output) of the bench step's dispersion
kernel: DRAM bytes per launch + the fingerprint of the K2 sources it was captured from (bench.k2_source_sha), so that
bench.py drops the figure as soon as the kernel sources change.
usage: tools/k2_traffic.py profiles/r2_k2_final.txt"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    txt = open(sys.argv[1]).read()
    kern = re.search(r"## kernel (\w+)", txt).group(1)

    def val(name):
        m = re.search(name + r"\s+([\d.]+)\s+(\w+)", txt)
        v, u = float(m.group(1)), m.group(2)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    out = {"kernel": kern, "workload": "C2 x 32 models (131072 columns), bench default",
           "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
           "k2_source_sha": bench.k2_source_sha(), "source": f"{os.path.relpath(sys.argv[1], ROOT)} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "k2_traffic.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
