#!/usr/bin/env python
"""Small driver for ncu captures: a few launches of the extraction kernel on N synthetic
patches (default 8192 64x64, binWidth 25, in-plane angles).  Run it plain first, then under
`ncu --set full -k regex:radb_extract -s 2 -c 1` (see profiles/README.md)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import multimodal_isic_b200 as pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--patches", type=int, default=8192)
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--bin-width", type=float, default=25.0)
ap.add_argument("--launches", type=int, default=4)
ap.add_argument("--literal", action="store_true")
a = ap.parse_args()
imgs, masks = pkg.synth.make_patches_torch(a.patches, a.size, seed=1234, device="cuda")
ex = pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": a.bin_width, "force2D": a.literal}})
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.launches):
    e0.record()
    out, st = ex.extract_batch(imgs, masks)
    e1.record()
    torch.cuda.synchronize()
    print("launch %d: %.3f ms, %.0f patches/s" % (i, e0.elapsed_time(e1), a.patches / e0.elapsed_time(e1) * 1e3))
ex.engine.set_profiling(True)
for i in range(3):
    ex.extract_batch(imgs, masks)
torch.cuda.synchronize()
print("per-kernel ms (3 launches):", {k: round(v, 3) for k, v in ex.engine.kernel_ms().items()})
print("smem bytes/CTA:", ex.engine.smem_bytes(a.size, a.size), "status!=0:", int((st != 0).sum()))
