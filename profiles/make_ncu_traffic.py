#!/usr/bin/env python
"""Regenerates profiles/ncu_traffic.json from one `ncu --set full --clock-control none` capture of
`python profiles/ncu_target.py` (the bench workload: one full evaluation), and stamps it with the hash of the library the
capture ran (bench.library_build_id).  bench.py quotes roofline.traffic from this file only when the stamp, the workload and
the kernel's duration agree with its own run.

usage (on the GPU box, after the plain run exited 0):
    ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_leaf|k_colsum|k_reduce" -c 8 \\
        -o gpurun_out/final -f python profiles/ncu_target.py
    python profiles/make_ncu_traffic.py gpurun_out/final.ncu-rep
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def unit_scale(u):
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[u]


def main():
    import bench
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    kernels = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "").strip()
        key = "k_prune" if short.startswith("k_walk") or short.startswith("k_prune") else short
        ent = {
            "kernel": name.split("(")[0].replace("void ", "").strip(),
            "dram_bytes_read": int(float(r[ix["dram__bytes_read.sum"]]) * unit_scale(units[ix["dram__bytes_read.sum"]])),
            "dram_bytes_write": int(float(r[ix["dram__bytes_write.sum"]]) * unit_scale(units[ix["dram__bytes_write.sum"]])),
            "duration_ms_under_ncu": float(r[ix["gpu__time_duration.sum"]]) * unit_scale(units[ix["gpu__time_duration.sum"]]),
        }
        kernels[key] = ent          # the last launch of each kernel (first-touch effects gone)
    out = {
        "_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from one `ncu --set full --clock-control none` capture of "
                    "`python profiles/ncu_target.py` (the bench workload); written by profiles/make_ncu_traffic.py. 'k_prune' is the "
                    "dominant pruning kernel of the step, whatever its instantiation (k_walk<0> by default).",
        "library_build_id": bench.library_build_id(),
        "source_report": os.path.basename(rep),
        "workload": {"cells": bench.N_CELLS, "sites": bench.N_SITES, "schedule": "single-launch walk",
                     "transition_matrices": "distances", "const_twin": False},
        "kernels": kernels,
    }
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(out, f, indent=2)
        f.write("\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
