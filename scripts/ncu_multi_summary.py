#!/usr/bin/env python
"""Summarise a multi-kernel ncu report (build / angle / misc kernels of one pass) into markdown.

  python scripts/ncu_multi_summary.py gpurun_out/prof_r1_v5.ncu-rep profiles/r1_v5_kernels.md "title"
"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_op_shared_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic",
]
SCALE = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    kernels = raw[2:]
    names = [re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "") for r in kernels]
    lines = ["# " + (title or rep), "",
             "Source: `%s` (`ncu --set full --clock-control none --import-source on`, one launch of each kernel; "
             "per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes)." % rep, "",
             "| metric | " + " | ".join(names) + " | unit |", "|---|" + "---|" * (len(names) + 1)]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append("| `%s` | %s | %s |" % (k, " | ".join(r[i] for r in kernels), units[i]))
    traffic = {}
    for nme, r in zip(names, kernels):
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        traffic[nme] = float(r[rd]) * SCALE[units[rd]] + float(r[wr]) * SCALE[units[wr]]
    grid = int(kernels[0][hdr.index("launch__grid_size")])
    lines += ["", "DRAM bytes per launch (read + write): " + ", ".join("%s %.1f MB" % (k, v / 1e6) for k, v in traffic.items()),
              "= %.0f B per patch over the pass (grid %d patches); algorithmic bytes 8 936 B/patch." % (sum(traffic.values()) / grid, grid)]
    src = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"])
    for blk in re.split(r'(?m)^"File Path",', src)[1:]:
        rows = list(csv.reader(io.StringIO('"File Path",' + blk)))
        fpath = rows[0][1].split("/")[-1]
        fname = re.sub(r"\(.*", "", rows[1][1]).replace("void ", "")
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
        if not fpath.startswith("radb_"):
            continue
        h = rows[hi]
        iS, iI, iT = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
        L = []
        for r in rows[hi + 1:]:
            if r and r[0].isdigit():
                try:
                    L.append((int(r[0]), r[1].strip(), int(r[iS] or 0), int(r[iI] or 0), int(r[iT] or 0)))
                except ValueError:
                    pass
        totI = sum(l[3] for l in L) or 1
        totS = sum(l[2] for l in L) or 1
        lines += ["", "## %s -- %s: top source lines by executed warp instructions (%d in this file)" % (fname, fpath, totI), "",
                  "| line | inst % | samples % | threads/inst | source |", "|---|---|---|---|---|"]
        for l in sorted(L, key=lambda l: -l[3])[:12]:
            lines.append("| %d | %.1f | %.1f | %.1f | `%s` |" % (l[0], 100 * l[3] / totI, 100 * l[2] / totS, l[4] / max(l[3], 1),
                                                               l[1][:80].replace("|", "\\|")))
    with open(out, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print(json.dumps({"grid": grid, "dram_bytes_per_launch": traffic, "dram_bytes_per_pass": sum(traffic.values())}))


if __name__ == "__main__":
    main()
