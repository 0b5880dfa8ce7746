#!/usr/bin/env python
"""End-to-end (host buffers -> rows on the host) rate of the host pipeline under different mask hand-over modes, on N
GPUs of one host (torchrun).  One JSON line per mode on rank 0: whole-job patches/s (max over ranks) and H2D GB/s per
rank.  Modes: raw uint8 masks, host-packed with T threads per rank, masks already packed by the caller.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 scripts/e2e_modes.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import multimodal_isic_b200 as pkg  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pkg.numa.bind_to_gpu_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = 100000
    ex = pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": 25.0, "force2D": False}}, device=local)
    F = ex.engine.F
    imgs, masks = pkg.synth.make_patches_torch(B, 64, seed=1234 + rank, device=dev)
    h_img, h_msk = imgs.cpu().pin_memory(), masks.cpu().pin_memory()
    h_out = torch.empty((B, F), dtype=torch.float64).pin_memory()
    h_st = torch.empty((B,), dtype=torch.int32).pin_memory()
    h_pk = torch.empty((B, ex.engine.packed_stride(64, 64)), dtype=torch.uint8).pin_memory()
    ex.engine.pack_masks_host(h_msk, h_pk.view(-1), 4)
    del imgs, masks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(name, fn, steps=4):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        if rank == 0:
            print(json.dumps({"mode": name, "n_gpus": world, "patches_per_s": world * B * steps / dt,
                              "h2d_gbs_per_rank": ex.pipeline.h2d_bytes * steps / dt / 1e9,
                              "packed_chunks": [ex.pipeline.packed_chunks, ex.pipeline.total_chunks],
                              "cores": os.cpu_count(), "numa": numa}), flush=True)

    ex.pipeline.adaptive = False
    ex.pipeline.pack_masks = False
    measure("raw uint8 masks (8192 B/patch over the link)", lambda: ex.pipeline.run(h_img, h_msk, h_out, h_st))
    ex.pipeline.pack_masks = True
    for t in (1, 2, 3, 4, 6, 8):
        if t * world > 2 * (os.cpu_count() or 1):
            continue
        ex.pipeline.pack_threads = t
        measure("host-packed, %d threads per rank" % t, lambda: ex.pipeline.run(h_img, h_msk, h_out, h_st))
    ex.pipeline.adaptive = True
    for t in (2, 4):
        ex.pipeline.pack_threads = t
        measure("adaptive packing, %d threads per rank" % t, lambda: ex.pipeline.run(h_img, h_msk, h_out, h_st))
    ex.pipeline.adaptive = False
    measure("masks already packed by the caller (4608 B/patch)", lambda: ex.pipeline.run(h_img, h_pk, h_out, h_st, masks_packed=True))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
