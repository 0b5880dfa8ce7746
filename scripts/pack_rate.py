#!/usr/bin/env python
"""Host mask-packing rate (radb_pack_mask_host) on pinned memory vs thread count; decides whether the packed-mask
transfer path pays on this host (it must beat the ~50 GB/s host-to-device link it relieves)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import multimodal_isic_b200 as pkg
lib = pkg.load_library()
n = 100000 * 4096
m = (torch.rand(n) < 0.4).to(torch.uint8).mul_(255).pin_memory()
out = torch.empty(n // 8, dtype=torch.uint8).pin_memory()
print("cpus", os.cpu_count())
for th in (1, 2, 4, 8, 12, 16, 24, 32):
    best = 1e9
    for _ in range(3):
        t = time.perf_counter(); lib.radb_pack_mask_host(m.data_ptr(), n, 255, out.data_ptr(), th); best = min(best, time.perf_counter() - t)
    print(th, "threads: %.1f GB/s (%.1f ms per 100k 64x64 masks)" % (n / best / 1e9, best * 1e3))
