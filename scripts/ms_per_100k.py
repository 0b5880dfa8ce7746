#!/usr/bin/env python
"""Device-resident milliseconds per 100 k 64x64 patches (full feature set) -- one line, for A/B runs under different
RADB_* environment knobs / RADB_LIB builds.  python scripts/ms_per_100k.py [binWidth] [label]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import multimodal_isic_b200 as pkg  # noqa: E402

bw = float(sys.argv[1]) if len(sys.argv) > 1 else 25.0
B = 100000
imgs, masks = pkg.synth.make_patches_torch(B, 64, seed=1234, device="cuda")
eng = pkg.Engine(bw, 255, pkg.in_plane_angles())
out = torch.empty((B, eng.F), dtype=torch.float64, device="cuda")
st = torch.empty((B,), dtype=torch.int32, device="cuda")
for _ in range(4):
    eng.extract_device(imgs, masks, out, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8):
    eng.extract_device(imgs, masks, out, st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 8
print("%s binWidth %g: %.3f ms per 100k = %.2f M patches/s  [RADB_LIB=%s RADB_BUILD_PAD=%s RADB_CHUNK=%s] checksum %.6e"
      % (" ".join(sys.argv[2:]), bw, ms, B / ms / 1e3, os.environ.get("RADB_LIB", "-"), os.environ.get("RADB_BUILD_PAD", "-"),
         os.environ.get("RADB_CHUNK", "-"), float(torch.nan_to_num(out).sum())))
