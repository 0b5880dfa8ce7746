#!/usr/bin/env python
"""BASELINE.json configs[2] and configs[3] at their stated scale: patch-sharded over N GPUs (torchrun, one
process per GPU, NCCL).  One JSON line per case on rank 0 with the rank count NCCL saw, per-GPU milliseconds
(CUDA events on the rank's stream) and the max/mean imbalance.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      scripts/bench_configs_multi.py [2] [3] >> gpurun_out/configs_multi.jsonl

configs[2]: 1024x1024 ISIC-like images tiled to 128x128 patches with the mask tile, sharded by tile list
            (cost = ROI pixels per tile), feature rows all-gathered in tile order.
configs[3]: patch sizes 32/64/224 mixed 1:1:1, mask coverage U(2 %, 100 %): contiguous cost-balanced shards
            (cost = ROI pixels + a per-patch constant), one ragged call per rank, padded all-gather.
Reference for the ordered fan-out: /root/reference/RadiomicExtractor.py:60-65."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import multimodal_isic_b200 as pkg  # noqa: E402


def emit(rank, **kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


def per_rank(ms, world, dev):
    t = torch.zeros(world, dtype=torch.float64, device=dev)
    t[dist.get_rank()] = ms
    dist.all_reduce(t)
    return [round(float(x), 4) for x in t.cpu().tolist()]


def timed_sharded(extract_fn, n, costs, world, dev, reps=5, warm=2):
    """Returns (best whole-step ms = max over ranks incl. the all-gather, per-rank extraction ms of that step, bounds)."""
    best, best_ranks, bounds = 1e30, None, None
    for it in range(warm + reps):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()

        def fn(lo, hi):
            r = extract_fn(lo, hi)
            e1.record()
            return r

        full, st, bounds = pkg.sharded_extract(fn, n, costs)
        e2.record()
        torch.cuda.synchronize()
        step = torch.tensor([e0.elapsed_time(e2)], device=dev)
        dist.all_reduce(step, op=dist.ReduceOp.MAX)
        ranks = per_rank(e0.elapsed_time(e1), world, dev)
        if it >= warm and float(step.item()) < best:
            best, best_ranks = float(step.item()), ranks
    return best, best_ranks, bounds, full, st


def config2(rank, world, dev, images_per_gpu=48):
    n_images = images_per_gpu * world
    # every rank generates only the images of its neighbourhood?  No: the shard boundaries depend on the global
    # cost vector, so every rank builds the global tile list (device generator, same seed) and keeps it resident
    imgs, masks = pkg.synth.make_patches_torch(n_images, 1024, seed=7, device=dev, chunk=8)
    t = lambda x: x.view(n_images, 8, 128, 8, 128).permute(0, 1, 3, 2, 4).reshape(-1, 128, 128).contiguous()
    ti, tm = t(imgs), t(masks)
    del imgs, masks
    n = len(ti)
    costs = (tm == 255).flatten(1).sum(1).double().cpu().numpy() + 512.0  # ROI pixels + fixed per-tile work
    ex = pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": 25}}, device=dev.index)

    def extract(lo, hi):
        return ex.engine.extract_device(ti[lo:hi], tm[lo:hi])

    ms, ranks, bounds, full, st = timed_sharded(extract, n, costs, world, dev)
    one_ms = None
    if rank == 0:  # the same list on one GPU: rows must agree bit for bit
        want, _ = ex.engine.extract_device(ti, tm)
        torch.cuda.synchronize()
        same = bool(torch.equal(full.view(torch.int64), want.view(torch.int64)))
    else:
        same = None
    emit(rank, config="configs[2]: 1024x1024 images tiled to 128x128 patches with lesion masks, sharded by tile list",
         n_gpus=world, nccl_ranks=dist.get_world_size(), images=n_images, tiles=n, valid_tiles=int((st == 0).sum()),
         shard_tiles=[bounds[r + 1] - bounds[r] for r in range(world)], ms_step=ms, tiles_per_s=n / ms * 1e3,
         ms_per_gpu=ranks, imbalance_max_over_mean=max(ranks) / (sum(ranks) / len(ranks)),
         gathered_equals_single_gpu_rows=same, bytes_per_patch=128 * 128 * 2 + 93 * 8,
         hbm_gbs_aggregate=n * (128 * 128 * 2 + 744) / ms / 1e6)


def config3(rank, world, dev, n_each_per_gpu=1500, sorted_list=False):
    n_each = n_each_per_gpu * world
    images, masks = [], []
    for k, H in enumerate((32, 64, 224)):
        g, m = pkg.synth.make_patches(64, H, seed=50 + k, coverage=(0.02, 1.0))
        for i in range(n_each):
            images.append(g[i % 64])
            masks.append(m[i % 64])
    if not sorted_list:
        order = np.random.default_rng(0).permutation(len(images))
        images = [images[i] for i in order]
        masks = [masks[i] for i in order]
    n = len(images)
    ip, mp, io, mo, hw = pkg.pack_ragged(images, masks)
    dip, dmp = torch.as_tensor(ip).to(dev), torch.as_tensor(mp).to(dev)
    ex = pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": 25}}, device=dev.index)
    roi = np.array([(m == 255).sum() for m in masks], dtype=np.float64)
    area = hw[:, 0].astype(np.float64) * hw[:, 1]
    for name, costs in (("equal counts (no cost model)", np.ones(n)), ("cost = H*W", area),
                        ("cost = ROI pixels + 0.15 * H*W + 600", roi + 0.15 * area + 600.0)):
        def extract(lo, hi):
            return ex.engine.extract_ragged(dip, dmp, io[lo:hi], mo[lo:hi], hw[lo:hi])

        ms, ranks, bounds, full, st = timed_sharded(extract, n, costs, world, dev, reps=3, warm=1)
        emit(rank, config="configs[3]: mixed patch sizes 32/64/224 (1:1:1), mask coverage 2-100 %, cost-balanced shards",
             list_order="sorted by size (all 32s, then 64s, then 224s): the load-imbalance stress" if sorted_list else "shuffled",
             n_gpus=world, nccl_ranks=dist.get_world_size(), patches=n, cost_model=name,
             shard_patches=[bounds[r + 1] - bounds[r] for r in range(world)], ms_step=ms, patches_per_s=n / ms * 1e3,
             mpixels_per_s=float(area.sum()) / ms / 1e3, ms_per_gpu=ranks,
             imbalance_max_over_mean=max(ranks) / (sum(ranks) / len(ranks)), invalid=int((st != 0).sum()))


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    which = [a for a in sys.argv[1:] if a in ("2", "3")] or ["2", "3"]
    if "2" in which:
        config2(rank, world, dev)
    if "3" in which:
        config3(rank, world, dev)
        config3(rank, world, dev, sorted_list=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
