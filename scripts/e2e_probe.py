#!/usr/bin/env python
"""What bounds the host-to-host path at N = 1: the raw pinned H2D rate of this box (one big copy, chunked copies on
several streams, with the D2H of the rows running against it), with and without binding the process to the GPU's NUMA
node (BIND=1), and the pipeline itself in its three hand-over modes.  One JSON line per measurement."""
import glob
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import multimodal_isic_b200 as pkg  # noqa: E402

bind = os.environ.get("BIND", "0") == "1"
numa = pkg.numa.bind_to_gpu_node(0) if bind else None
print(json.dumps({"bind": bind, "numa": numa, "gpu_node": pkg.numa.gpu_numa_node(0), "cpu_count": os.cpu_count(),
                  "affinity": len(os.sched_getaffinity(0)), "nodes": len(glob.glob("/sys/devices/system/node/node[0-9]*"))}), flush=True)
B, STEPS = 100000, 10
ex = pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": 25.0, "force2D": False}})
imgs, masks = pkg.synth.make_patches_torch(B, 64, seed=1234, device="cuda")
h_img, h_msk = imgs.cpu().pin_memory(), masks.cpu().pin_memory()
h_out = torch.empty((B, ex.engine.F), dtype=torch.float64).pin_memory()
h_st = torch.empty((B,), dtype=torch.int32).pin_memory()
d_out = torch.empty((B, ex.engine.F), dtype=torch.float64, device="cuda")


def timed(fn, n=STEPS):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


nbytes = h_img.numel()
dt = timed(lambda: imgs.copy_(h_img, non_blocking=True))
print(json.dumps({"what": "one H2D copy of the pixels", "bytes": nbytes, "ms": dt * 1e3, "gbs": nbytes / dt / 1e9}), flush=True)
streams = [torch.cuda.Stream() for _ in range(6)]


def chunked(with_d2h):
    for k, s in enumerate(range(0, B, 8192)):
        n = min(8192, B - s)
        with torch.cuda.stream(streams[k % 6]):
            imgs[s:s + n].copy_(h_img[s:s + n], non_blocking=True)
            masks[s:s + n // 8].copy_(h_msk[s:s + n // 8], non_blocking=True)  # 1/8 of the mask bytes = packed size
            if with_d2h:
                h_out[s:s + n].copy_(d_out[s:s + n], non_blocking=True)


for wd in (False, True):
    dt = timed(lambda: chunked(wd))
    nb = nbytes + nbytes // 8
    print(json.dumps({"what": "13 chunks on 6 streams, pixels + packed-size masks" + (" + D2H of the rows" if wd else ""),
                      "bytes": nb, "ms": dt * 1e3, "gbs": nb / dt / 1e9}), flush=True)
pk = torch.empty((B, ex.engine.packed_stride(64, 64)), dtype=torch.uint8).pin_memory()
ex.engine.pack_masks_host(h_msk, pk.view(-1), 8)
for mode in ("caller-packed", "uint8"):
    ex.pipeline.pack_masks = mode != "uint8"
    fn = (lambda: ex.pipeline.run(h_img, pk, h_out, h_st, masks_packed=True)) if mode == "caller-packed" else \
        (lambda: ex.pipeline.run(h_img, h_msk, h_out, h_st))
    dt = timed(fn)
    print(json.dumps({"what": "pipeline, masks " + mode, "ms": dt * 1e3, "patches_per_s": B / dt,
                      "h2d_gbs": ex.pipeline.h2d_bytes / dt / 1e9}), flush=True)
ex.pipeline.pack_masks = True
for spawn in ("1", "0", "1", "0"):  # RADB_PACK_SPAWN=1: threads spawned per call (round 1); 0: persistent pool
    os.environ["RADB_PACK_SPAWN"] = spawn
    for thr in (6, 8, 12, 16):
        ex.pipeline.pack_threads = thr
        dt = timed(lambda: ex.pipeline.run(h_img, h_msk, h_out, h_st))
        print(json.dumps({"what": "pipeline, masks packed by the pipeline", "spawn_per_call": spawn == "1", "pack_threads": thr,
                          "ms": dt * 1e3, "patches_per_s": B / dt, "h2d_gbs": ex.pipeline.h2d_bytes / dt / 1e9,
                          "host_ms_last_run": {k: round(v * 1e3, 3) for k, v in ex.pipeline.stats.items()}}), flush=True)
for spawn in ("1", "0"):
    os.environ["RADB_PACK_SPAWN"] = spawn
    for thr in (4, 8, 16):
        n = 8192
        t0 = time.perf_counter()
        for _ in range(20):
            ex.engine.pack_masks_host(h_msk[:n], pk.view(-1), thr)
        dt = (time.perf_counter() - t0) / 20
        print(json.dumps({"what": "host packing of one 8192-patch chunk (link idle)", "spawn_per_call": spawn == "1", "threads": thr,
                          "ms": dt * 1e3, "gbs_read": n * 4096 / dt / 1e9}), flush=True)
