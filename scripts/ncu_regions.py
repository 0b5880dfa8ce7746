#!/usr/bin/env python
"""Aggregate an ncu source page by code region (function / phase markers found in the CURRENT sources;
the report must come from the same source revision).  usage: ncu_regions.py report.ncu-rep [patches]"""
import csv, io, os, re, subprocess, sys
rep = sys.argv[1]
npatch = float(sys.argv[2]) if len(sys.argv) > 2 else 8192.0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
marks = {}
for fn in ("radb_kernels.cuh", "radb_features.cuh"):
    m = []
    for ln, line in enumerate(open(os.path.join(root, "multimodal-isic_b200", "csrc", fn)), 1):
        g = re.match(r"\s*// ---- (phase [0-9a-z]+)", line) or re.match(r"__device__ (?:__forceinline__ |__noinline__ )?[\w ]*?\b(\w+)\(", line)
        if g:
            m.append((ln, g.group(1)))
    marks[fn] = m
for blk in re.split(r'(?m)^"File Path",', out)[1:]:
    rows = list(csv.reader(io.StringIO('"File Path",' + blk)))
    fpath = rows[0][1].split("/")[-1]
    kname = re.sub(r"\(.*", "", rows[1][1]).replace("void ", "")
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    h = rows[hi]
    iS, iI, iT = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    agg = {}
    for r in rows[hi + 1:]:
        if r and r[0].isdigit():
            try:
                ln, s_, i_, t_ = int(r[0]), int(r[iS] or 0), int(r[iI] or 0), int(r[iT] or 0)
            except ValueError:
                continue
            name = fpath
            for l0, nm in marks.get(fpath, []):
                if l0 <= ln:
                    name = nm
            a = agg.setdefault(name, [0, 0, 0])
            a[0] += s_; a[1] += i_; a[2] += t_
    tot = sum(a[1] for a in agg.values())
    if tot < 5e6:
        continue
    print("== %s / %s: %.1fk warp instructions per patch" % (kname, fpath, tot / npatch / 1000))
    for k, (s_, i_, t_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print("   %-18s %5.1f%%  threads/inst %4.1f  (%.1fk/patch)" % (k, 100 * i_ / tot, t_ / max(i_, 1), i_ / npatch / 1000))
