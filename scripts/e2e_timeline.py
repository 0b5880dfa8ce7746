#!/usr/bin/env python
"""Per-chunk CUDA-event timeline of the end-to-end host pipeline (HostPipeline.run, 100 k 64x64 patches from pinned
host buffers): when each 8192-patch chunk's H2D copies, kernels and D2H copy start and end.  Shows which resource
the e2e rate sits on (the link: copies back to back, kernels hidden under them).  python scripts/e2e_timeline.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import multimodal_isic_b200 as pkg  # noqa: E402

B = 100000
ex = pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": 25.0, "force2D": False}})
imgs, masks = pkg.synth.make_patches_torch(B, 64, seed=1234, device="cuda")
h_img, h_msk = imgs.cpu().pin_memory(), masks.cpu().pin_memory()
h_out = torch.empty((B, ex.engine.F), dtype=torch.float64).pin_memory()
h_st = torch.empty((B,), dtype=torch.int32).pin_memory()
for mode, pack in (("masks packed on the host (default)", True), ("raw uint8 masks", False)):
    ex.pipeline.pack_masks = pack
    for _ in range(2):
        ex.pipeline.run(h_img, h_msk, h_out, h_st)
    ex.pipeline.timeline = []
    ex.pipeline.run(h_img, h_msk, h_out, h_st)
    rows = ex.pipeline.timeline_ms()
    print("# %s: %d bytes H2D per step, %d host pack threads" % (mode, ex.pipeline.h2d_bytes, ex.pipeline.pack_threads))
    print("# chunk patches  h2d_start  h2d_end/kernels_start  kernels_end  d2h_end   (ms since the first chunk's start)")
    for r in rows:
        print("%5d %7d %10.3f %10.3f %10.3f %10.3f" % r)
    end = max(r[5] for r in rows)
    link = sum(r[3] - r[2] for r in rows)
    print("# step %.3f ms = %.2f M patches/s; sum of H2D intervals %.3f ms (%.1f GB/s while copying), kernels %.3f ms in total\n"
          % (end, B / end / 1e3, link, ex.pipeline.h2d_bytes / link / 1e6, sum(r[4] - r[3] for r in rows)))
