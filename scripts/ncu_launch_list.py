#!/usr/bin/env python
"""Markdown summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (per-kernel launches, total time, share).

  python scripts/ncu_launch_list.py gpurun_out/launches_r1.csv profiles/launches_r1.md "command that was profiled"
"""
import csv
import re
import sys

src, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = {}
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[r[iu]]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) * scale
tot = sum(a[1] for a in agg.values())
with open(out, "w") as fh:
    fh.write("# ncu launch list of `%s`\n\n" % cmd)
    fh.write("`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:radb_` (engine kernels only). Per-launch times "
             "under ncu are cold-cache and serialised: compare SHARES with `roofline.kernel_ms_parts` of the bench line, not absolutes.\n\n")
    fh.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
        fh.write("| `%s` | %d | %.2f | %.1f %% |\n" % (k, n, ms, 100 * ms / tot))
    fh.write("\nRaw list: `profiles/%s`.\n" % src.split("/")[-1])
print(open(out).read())
