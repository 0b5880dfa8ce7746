#!/usr/bin/env python
"""Summarise an ncu report (read here, on the GPU-less box) into a markdown file under profiles/.

  python scripts/ncu_summary.py gpurun_out/prof_r1_v2.ncu-rep profiles/r1_v2_kernel.md "title"
"""
import csv
import io
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
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines = ["# " + (title or rep), "", "Source: `%s` (`ncu --set full --clock-control none --import-source on`, one launch)." % rep, "",
             "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in m:
            lines.append("| `%s` | %s | %s |" % (k, m[k][0], m[k][1]))
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    hi = next(i for i, r in enumerate(src) if r and r[0] == "Line No")
    h = src[hi]
    iS, iI, iT = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    rows, cur_file = [], ""
    for r in src:
        if r and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        if r and r[0].isdigit():
            try:
                rows.append((cur_file, int(r[0]), r[1].strip(), int(r[iS] or 0), int(r[iI] or 0), int(r[iT] or 0)))
            except ValueError:
                pass
    totI = sum(r[4] for r in rows) or 1
    totS = sum(r[3] for r in rows) or 1
    lines += ["", "Warp instructions attributed to source lines: %d; stall samples: %d." % (totI, totS), "",
              "## Top source lines by executed warp instructions", "",
              "| file:line | inst % | samples % | threads/inst | source |", "|---|---|---|---|---|"]
    for r in sorted(rows, key=lambda r: -r[4])[:30]:
        lines.append("| %s:%d | %.1f | %.1f | %.1f | `%s` |" % (r[0], r[1], 100 * r[4] / totI, 100 * r[3] / totS,
                                                                  r[5] / max(r[4], 1), r[2][:80].replace("|", "\\|")))
    with open(out, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    dram = None
    try:
        rd, wr = float(m["dram__bytes_read.sum"][0]), float(m["dram__bytes_write.sum"][0])
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
        dram = rd * scale[m["dram__bytes_read.sum"][1]] + wr * scale[m["dram__bytes_write.sum"][1]]
    except Exception:
        pass
    print("wrote", out, "dram bytes per launch:", dram, "grid:", m.get("launch__grid_size", ("?",))[0])


if __name__ == "__main__":
    main()
