#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs[2..4] (the parity-test configurations; bench.py's line is
configs[1]).  Device-resident inputs, CUDA events on the launching stream, 2 warm-up + best of 5; one JSON
line per case on stdout.  Run on a GPU box:  python scripts/bench_configs.py > gpurun_out/configs.jsonl"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import multimodal_isic_b200 as pkg  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def emit(**kw):
    print(json.dumps(kw), flush=True)


def config2_tiled_images(n_images=48):
    """1024x1024 ISIC-like images (one lesion spanning the image) tiled to 64 non-overlapping 128x128 patches
    with the mask tile; tiles without a valid ROI come back with status != 0."""
    imgs, masks = pkg.synth.make_patches_torch(n_images, 1024, seed=7, device="cuda", chunk=8)
    t = lambda x: x.view(n_images, 8, 128, 8, 128).permute(0, 1, 3, 2, 4).reshape(-1, 128, 128).contiguous()
    ti, tm = t(imgs), t(masks)
    ex = pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": 25}})
    out, st = ex.extract_batch(ti, tm)
    ms = timed(lambda: ex.extract_batch(ti, tm))
    emit(config="configs[2]: 1024x1024 images tiled to 128x128 patches with lesion masks", tiles=len(ti),
         valid_tiles=int((st == 0).sum()), ms=ms, patches_per_s=len(ti) / ms * 1e3,
         bytes_per_patch=128 * 128 * 2 + 93 * 8, hbm_gbs=len(ti) * (128 * 128 * 2 + 744) / ms / 1e6)


def config3_mixed_sizes(n_each=1500):
    """sizes 32 / 64 / 224 mixed 1:1:1, mask coverage U(2 %, 100 %), one radb_extract_ragged call"""
    images, masks = [], []
    for k, H in enumerate((32, 64, 224)):
        g, m = pkg.synth.make_patches(64, H, seed=50 + k, coverage=(0.02, 1.0))
        for i in range(n_each):
            images.append(g[i % 64])
            masks.append(m[i % 64])
    order = np.random.default_rng(0).permutation(len(images))
    images = [images[i] for i in order]
    masks = [masks[i] for i in order]
    ip, mp, io, mo, hw = pkg.pack_ragged(images, masks)
    ex = pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": 25}})
    dip, dmp = torch.as_tensor(ip).cuda(), torch.as_tensor(mp).cuda()
    ms = timed(lambda: ex.engine.extract_ragged(dip, dmp, io, mo, hw))
    px = int((hw[:, 0].astype(np.int64) * hw[:, 1]).sum())
    emit(config="configs[3]: mixed patch sizes 32/64/224 (1:1:1), mask coverage 2-100 %", patches=len(images), ms=ms,
         patches_per_s=len(images) / ms * 1e3, mpixels_per_s=px / ms / 1e3)
    for H in (32, 64, 224):
        idx = [i for i in range(len(images)) if images[i].shape[0] == H]
        di = torch.as_tensor(np.stack([images[i] for i in idx])).cuda()
        dm = torch.as_tensor(np.stack([masks[i] for i in idx])).cuda()
        ms = timed(lambda: ex.extract_batch(di, dm))
        emit(config="configs[3] size class %dx%d alone" % (H, H), patches=len(idx), ms=ms, patches_per_s=len(idx) / ms * 1e3)


def config4_binwidth_sweep(n_patches=1000000):
    """BASELINE.json configs[4]: 64x64 uint16 intensities in [0, 2048): binWidth 8..64 <-> 256..32 gray levels,
    `n_patches` patches per binWidth (1 M by default; the 256 distinct synthetic patches are tiled on the device)."""
    g, m = pkg.synth.make_patches(256, 64, seed=60, dtype=np.uint16, vmax=2047)
    g0 = torch.as_tensor(g.view(np.int16)).cuda()
    m0 = torch.as_tensor(m).cuda()
    reps = (n_patches + 255) // 256
    gi = g0.repeat(reps, 1, 1)[:n_patches].contiguous().view(torch.uint16)
    mi = m0.repeat(reps, 1, 1)[:n_patches].contiguous()
    for bw in (64, 32, 16, 8):
        n = n_patches
        eng = pkg.Engine(bw, 255, pkg.in_plane_angles(), max_ng=2048 // bw)
        probe = min(n, 4096)
        eng.extract_device(gi[:probe], mi[:probe])
        eng.set_profiling(True)
        out, st = eng.extract_device(gi[:probe], mi[:probe])
        torch.cuda.synchronize()
        parts = eng.kernel_ms()
        eng.set_profiling(False)
        ms = timed(lambda: eng.extract_device(gi, mi), reps=2, warm=1)
        emit(config="configs[4]: binWidth sweep on 64x64 uint16 patches", binWidth=bw, max_ng=2048 // bw, patches=n, ms=ms,
             patches_per_s=n / ms * 1e3, probe_patches=probe, kernel_ms_parts_probe=parts,
             smem_build_bytes=eng.smem_bytes(64, 64, pkg._abi.DTYPE_U16), invalid_in_probe=int((st != 0).sum()))


def reference_workload(n_images=128):
    """What extract_radiomics.py actually feeds (RadiomicExtractor.py:29-48, params.yml): whole 600x450 BGR dermoscopy
    images (HAM10000 size) with one lesion mask each -> gray / R / G / B planes on the device -> 4 executes per image,
    binWidth 10, force2D (literal: one along-row angle), shape2D + 93 features; wide mode (one CTA per plane)."""
    H, W = 450, 600
    g, m = pkg.synth.make_patches(8, H, W, seed=70)
    rng = np.random.default_rng(0)
    bgr = np.stack([np.stack([np.clip(g[i % 8].astype(int) + rng.integers(-20, 20), 0, 255).astype(np.uint8) for _ in range(3)], -1)
                    for i in range(n_images)])
    msk = np.stack([m[i % 8] for i in range(n_images)])
    ex = pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": 10, "force2D": True},
                                 "featureClass": {k: None for k in ("firstorder", "shape2D", "glcm", "gldm", "glrlm", "glszm", "ngtdm")}})
    db, dm = torch.as_tensor(bgr).cuda(), torch.as_tensor(msk).cuda()
    out, st = ex.engine.extract_bgr(db, dm)
    ms = timed(lambda: ex.engine.extract_bgr(db, dm), reps=3, warm=1)
    emit(config="reference workload: whole 600x450 BGR images, 4 executes per image (gray/R/G/B), binWidth 10, literal force2D, 102 features",
         images=n_images, ms=ms, images_per_s=n_images / ms * 1e3, executes_per_s=4 * n_images / ms * 1e3,
         mpixels_per_s=4 * n_images * H * W / ms / 1e3, invalid=int((st != 0).sum()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["ref", "2", "3", "4"]
    if "ref" in which:
        reference_workload()
    if "2" in which:
        config2_tiled_images()
    if "3" in which:
        config3_mixed_sizes()
    if "4" in which:
        n = [int(a.split("=")[1]) for a in which if a.startswith("n=")]
        config4_binwidth_sweep(n[0] if n else 1000000)
