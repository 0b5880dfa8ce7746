#!/usr/bin/env python
"""Aggregate an ncu source page by code region (function / phase markers found in the sources)."""
import csv, io, re, subprocess, sys, os
rep = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout)))
hi = next(i for i, r in enumerate(src) if r and r[0] == "Line No")
h = src[hi]
iS, iI, iT = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
marks = {}
for fn in ("radb_kernels.cuh", "radb_features.cuh"):
    m = []
    for ln, line in enumerate(open(os.path.join(root, "multimodal-isic_b200", "csrc", fn)), 1):
        g = re.match(r"\s*// ---- (phase [0-9a-z]+)", line) or re.match(r"__device__ (?:__forceinline__ )?\w[\w ]*?\b(\w+)\(", line)
        if g:
            m.append((ln, g.group(1)))
    marks[fn] = m
agg, cur = {}, ""
for r in src:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    if r and r[0].isdigit():
        try:
            ln, s, i, t = int(r[0]), int(r[iS] or 0), int(r[iI] or 0), int(r[iT] or 0)
        except ValueError:
            continue
        name = cur
        for l0, nm in marks.get(cur, []):
            if l0 <= ln:
                name = nm
        a = agg.setdefault(name, [0, 0, 0])
        a[0] += s; a[1] += i; a[2] += t
totI = sum(a[1] for a in agg.values()); totS = sum(a[0] for a in agg.values())
print("| region | inst % | samples % | threads/inst |\n|---|---|---|---|")
for k, (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if i * 1000 > totI:
        print("| %s | %.1f | %.1f | %.1f |" % (k, 100 * i / totI, 100 * s / totS, t / max(i, 1)))
