#!/usr/bin/env python
"""A/B of the host-to-host pipeline's chunk schedule (HostPipeline.chunk_schedule): 100 k 64x64 patches from pinned
host buffers, wall clock over STEPS runs per schedule, one JSON line each.  python scripts/e2e_ramp_ab.py [specs...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import multimodal_isic_b200 as pkg  # noqa: E402

B, STEPS = 100000, 10
specs = sys.argv[1:] or ["off", "on", "off", "on", "4/4", "4,2/4", "4/2,4", "4/2,4,8", "P:off", "P:on", "P:4/4", "P:4/2,4,8", "off@4096", "on@4096",
                         "off@16384", "4,2/2,4,8@16384"]
ex = pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": 25.0, "force2D": False}})
imgs, masks = pkg.synth.make_patches_torch(B, 64, seed=1234, device="cuda")
h_img, h_msk = imgs.cpu().pin_memory(), masks.cpu().pin_memory()
h_out = torch.empty((B, ex.engine.F), dtype=torch.float64).pin_memory()
h_st = torch.empty((B,), dtype=torch.int32).pin_memory()
pk = torch.empty((B, ex.engine.packed_stride(64, 64)), dtype=torch.uint8).pin_memory()
ex.engine.pack_masks_host(h_msk, pk.view(-1), 8)
ref = None
for spec in specs:
    chunk = 8192
    threads = 12
    prepacked = spec.startswith("P:")  # caller-packed masks
    spec = spec[2:] if prepacked else spec
    if "#" in spec:  # "spec#pack_threads"
        spec, t = spec.split("#")
        threads = int(t)
    slots = 6
    if "%" in spec:  # "spec%slots"
        spec, t = spec.split("%")
        slots = int(t)
    if ex.pipeline.slots != slots:
        ex.pipeline.slots = slots
        ex.pipeline._key = None  # new staging buffers
    if "@" in spec:  # "spec@chunk"
        spec, c = spec.split("@")
        chunk = int(c)
    ex.pipeline.chunk = chunk
    ex.pipeline.pack_threads = threads
    ex.pipeline.ramp = spec if "/" in spec else spec == "on"
    step = (lambda: ex.pipeline.run(h_img, pk, h_out, h_st, masks_packed=True)) if prepacked else \
        (lambda: ex.pipeline.run(h_img, h_msk, h_out, h_st))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(STEPS):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / STEPS
    rows = h_out.clone()
    if ref is None:
        ref = rows
    print(json.dumps({"ramp": spec, "caller_packed": prepacked, "chunk": chunk, "chunks": ex.pipeline.total_chunks, "ms_per_step": dt * 1e3,
                      "patches_per_s": B / dt, "h2d_gbs": ex.pipeline.h2d_bytes / dt / 1e9,
                      "pack_threads": threads, "slots": slots, "host_ms_last_run": {k: round(v * 1e3, 3) for k, v in ex.pipeline.stats.items()},
                      "rows_equal_first_schedule": bool(torch.equal(rows.view(torch.int64), ref.view(torch.int64)))}), flush=True)
