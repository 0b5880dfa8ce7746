#!/usr/bin/env python
"""Headline benchmark: radiomic patches/sec (2-D 64x64, full feature set) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W]          # our CUDA engine
  python bench.py --impl reference [...]                        # the CPU path, all host cores

A step = one pass of the hot path (discretise -> 5 texture matrices -> 93 fp64 features)
over one batch of synthetic lesion-like patches (SURVEY.md section 8 d).  N=1 runs
BASELINE.json configs[1]: 100k 64x64 uint8 patches, full feature set.  With N>1 (torchrun) each
rank holds its own 100k-patch shard (weak scaling) and the step ends with the path's only
collective, the all-gather of the feature block.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "radiomic patches/sec (2D 64x64, full feature set)"
UNIT = "patches/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--patches", type=int, default=100000, help="patches per GPU per step")
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--bin-width", type=float, default=25.0)
    ap.add_argument("--literal-force2d", action="store_true",
                    help="pyradiomics' literal angle set for a 2-D array with force2D (1 angle) instead of in-plane (4)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target wall time of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the secondary (binWidth 10 / literal) measurements")
    ap.add_argument("--no-e2e", action="store_true", help="device-resident measurement only (parameter sweeps)")
    ap.add_argument("--e2e-pack", default="auto", choices=["auto", "on", "off"], help="host-side mask packing of the e2e path")
    ap.add_argument("--pack-threads", type=int, default=0)
    ap.add_argument("--e2e-ramp", default="off",
                    help="graded chunk sizes at both ends of the host-to-host pipeline (HostPipeline.chunk_schedule)")
    ap.add_argument("--gather", default="slices", choices=["chunks", "slices"],
                    help="N > 1: all-gather per engine chunk of ONE extraction call (completion events), or per separately "
                         "extracted slice (--slices)")
    ap.add_argument("--gather-chunks", type=int, default=4, help="chunks per shard in --gather chunks mode")
    ap.add_argument("--slices", default="0.45,0.3,0.15,0.1",
                    help="N > 1: fractions of the shard extracted per slice (the all-gather of a slice overlaps the next slice)")
    return ap.parse_args()


def settings_dict(args):
    return {"label": 255, "binWidth": args.bin_width, "force2D": bool(args.literal_force2d)}


def bytes_per_patch(H, W, F, pix=1):
    # SURVEY.md section 8 d: pixels + uint8 mask + fp64 feature row
    return H * W * pix + H * W + 8 * F


# ----------------------------------------------------------------------------- CPU arm
def _cpu_init():
    os.environ["OMP_NUM_THREADS"] = "1"


def _cpu_chunk(job):
    imgs, masks, settings = job
    from oracle import cmatrices, radiomics_oracle as orc

    out = []
    for b in range(len(imgs)):
        out.append(list(orc.execute(imgs[b], masks[b], settings, matrix_backend=cmatrices).values()))
    return np.asarray(out)


class CpuArm:
    """The reference's CPU path for this metric.  pyradiomics is not installable here (SURVEY.md
    fact 2), so the engine timed is the oracle: C matrix builders + NumPy feature formulas, the
    same split pyradiomics has (kind = "port"), fanned out over all host cores the way
    RadiomicExtractor.py:60-65 fans records over a process pool."""

    def __init__(self, settings, cores=None):
        import multiprocessing as mp

        from oracle import cmatrices

        cmatrices.build()
        self.settings = settings
        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init)

    def run(self, imgs, masks):
        n = len(imgs)
        per = max(1, min(16, n // (self.cores * 4) or 1))
        jobs = [(imgs[s:s + per], masks[s:s + per], self.settings) for s in range(0, n, per)]
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_chunk, jobs)
        dt = time.perf_counter() - t0
        return np.concatenate(res), dt

    def calibrate(self, imgs, masks, target_s):
        n0 = min(len(imgs), self.cores * 4)
        _, dt = self.run(imgs[:n0], masks[:n0])
        rate = n0 / dt
        return int(max(self.cores * 4, min(len(imgs), rate * target_s)))

    def close(self):
        self.pool.close()
        self.pool.join()


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from multimodal_isic_b200 import synth

    st = settings_dict(args)
    arm = CpuArm(st)
    pool_n = 4096
    imgs, masks = synth.make_patches(pool_n, args.size, seed=1234)
    n = arm.calibrate(imgs, masks, args.cpu_seconds / max(1, args.steps + args.warmup) * 1.0)
    n = max(arm.cores * 4, min(n, pool_n))
    for _ in range(args.warmup):
        arm.run(imgs[:n], masks[:n])
    tot = 0.0
    for _ in range(args.steps):
        _, dt = arm.run(imgs[:n], masks[:n])
        tot += dt
    arm.close()
    value = n * args.steps / tot
    F = 93
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, n, F), l2_policy="n/a (CPU arm)"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port",
                         "sample": "%d of the workload's 64x64 patches per step, oracle (C matrices + NumPy features), "
                                   "multiprocessing pool over %d cores" % (n, arm.cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, patches_per_step, F):
    return {
        "workload": "BASELINE.json configs[1]: full 2D feature set (firstorder/GLCM/GLDM/GLRLM/GLSZM/NGTDM, F=%d) on "
                    "%dx%d uint8 synthetic lesion-like patches" % (F, args.size, args.size),
        "patches_per_gpu_per_step": patches_per_step, "patch": [args.size, args.size], "features": F,
        "binWidth": args.bin_width,
        "angles": "literal force2D on 2-D input (1 angle, 2 neighbours)" if args.literal_force2d
        else "in-plane (4 angles, 8 neighbours)",
        "label": 255, "parallelism": ("patch-sharded (block-cyclic), one process per GPU; one extraction call per step, the all-gather of each of "
                                        "its %d chunks starts on the chunk's completion event and lands in place" % args.gather_chunks)
        if args.gather == "chunks" else
        ("patch-sharded (block-cyclic), one process per GPU, all-gather of the feature block in separately extracted slices "
         "(%s) overlapped with the extraction, written in place" % args.slices),
        "l2_policy": "inputs (%.0f MB per step) larger than the 126 MB L2" % (patches_per_step * args.size * args.size * 2 / 1e6),
    }


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nme, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- GPU arm
def gpu_arm(args):
    import torch
    import torch.distributed as dist

    import multimodal_isic_b200 as pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # N > 1: keep this rank's pinned staging buffers on the NUMA node of its GPU (before anything is allocated)
    numa = pkg.numa.bind_to_gpu_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    st = settings_dict(args)
    ex = pkg.RadiomicsExtractor({"setting": st}, device=local)
    if args.e2e_pack != "auto":
        ex.pipeline.pack_masks = args.e2e_pack == "on"
    if args.pack_threads:
        ex.pipeline.pack_threads = args.pack_threads
    ex.pipeline.ramp = args.e2e_ramp if "/" in args.e2e_ramp else args.e2e_ramp == "on"
    F = ex.engine.F
    B, H = args.patches, args.size
    # every rank owns a different shard of the (virtual) global patch list
    imgs, masks = pkg.synth.make_patches_torch(B, H, seed=1234 + rank, device=dev)
    out = torch.empty((B, F), dtype=torch.float64, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    gathered = torch.empty((world * B, F), dtype=torch.float64, device=dev) if world > 1 else None
    # N > 1: the shard is extracted in `pieces` slices and the all-gather of slice k (NCCL on a side stream) overlaps
    # the extraction of slice k+1; the global patch list is dealt to the ranks in blocks (block-cyclic), so every
    # slice's all-gather lands in its final rows; the step ends when the full [world * B, F] matrix is on every rank
    pieces = tuple(float(x) for x in args.slices.split(","))
    og = None
    if world > 1 and args.gather == "chunks":
        # one extraction call per step; the all-gather of chunk k waits for the engine's completion event of chunk k
        ex.engine.set_chunk(-(-B // args.gather_chunks))
        og = pkg.OverlappedGather(B, F, world, dev, bounds=pkg.OverlappedGather.chunk_bounds(ex.engine, B, H, H))
    elif world > 1:
        og = pkg.OverlappedGather(B, F, world, dev, pieces=pieces)  # block-cyclic: zero-copy gather

    def extract_slice(lo, hi, o, s):
        ex.engine.extract_device(imgs[lo:hi], masks[lo:hi], o, s)

    def step():
        if world > 1 and args.gather == "chunks":
            og.run_chunked(ex.engine, imgs, masks, out, status, gathered)
        elif world > 1:
            og.run(extract_slice, out, status, gathered)
        else:
            ex.engine.extract_device(imgs, masks, out, status)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ex.engine.launches
    ms = timed(step, args.steps)
    launches = ex.engine.launches - l0
    # per-kernel device time (CUDA events recorded by the library on the launching stream around the
    # build / angle / misc kernels of every chunk), averaged over the same number of steps
    ex.engine.set_profiling(True)
    for _ in range(args.steps):
        ex.engine.extract_device(imgs, masks, out, status)
    torch.cuda.synchronize()
    kparts = {k: v / args.steps for k, v in ex.engine.kernel_ms().items()}
    ex.engine.set_profiling(False)
    kms = sum(kparts.values())
    bad = int((status != 0).sum().item())

    # ---- end to end through the public API with HOST buffers (pinned), copies in the timed region
    h_img = imgs.cpu().pin_memory()
    h_msk = masks.cpu().pin_memory()
    h_out = torch.empty((B, F), dtype=torch.float64).pin_memory()
    h_st = torch.empty((B,), dtype=torch.int32).pin_memory()

    e2e_dev = torch.empty((B, F), dtype=torch.float64, device=dev) if world > 1 else None
    e2e_gathered = torch.empty((world * B, F), dtype=torch.float64, device=dev) if world > 1 else None

    def e2e_step():
        # host buffers in, host rows out; N > 1: the rows THIS run produced are also kept on the device and
        # all-gathered (the path's only collective), so every rank ends the step holding the e2e result
        ex.pipeline.run(h_img, h_msk, h_out, h_st, device_out=e2e_dev)
        if world > 1:
            dist.all_gather_into_tensor(e2e_gathered, e2e_dev)

    for _ in range(0 if args.no_e2e else 2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(1 if args.no_e2e else args.steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(2)
    same = bool(torch.equal(h_out, out.cpu()))
    e2e_stats = dict(h2d=ex.pipeline.h2d_bytes, packed=ex.pipeline.packed_chunks, chunks=ex.pipeline.total_chunks,
                     pack_on=ex.pipeline.pack_masks, threads=ex.pipeline.pack_threads, adaptive=ex.pipeline.adaptive)
    # ---- N > 1: one step outside the timed region, checked end to end -- rank 0 receives every rank's input
    # shard (broadcast), recomputes it on its own GPU and compares the gathered blocks bit for bit
    # (/root/reference/RadiomicExtractor.py:63-65: the fan-out is order-preserving)
    gather_ok = None
    if world > 1:
        step()
        torch.cuda.synchronize()
        ok = True
        t_img, t_msk = torch.empty_like(imgs), torch.empty_like(masks)
        t_out = torch.empty((B, F), dtype=torch.float64, device=dev)
        for r in range(world):
            if rank == r:
                t_img.copy_(imgs)
                t_msk.copy_(masks)
            dist.broadcast(t_img, src=r)
            dist.broadcast(t_msk, src=r)
            if rank == 0:
                ex.engine.extract_device(t_img, t_msk, t_out, status)
                gidx = torch.as_tensor(og.global_index(r, np.arange(B)), device=dev)  # rows of rank r's patches in `gathered`
                ok = ok and bool(torch.equal(gathered.index_select(0, gidx).view(torch.int64), t_out.view(torch.int64)))
                ok = ok and bool(torch.equal(e2e_gathered[r * B:(r + 1) * B].view(torch.int64), t_out.view(torch.int64)))
        del t_img, t_msk, t_out
        gather_ok = ok
        ex.engine.extract_device(imgs, masks, out, status)  # restore this rank's status / rows
        torch.cuda.synchronize()
    # ---- N > 1: where the step's time beyond one GPU's goes (outside the headline region, same timing rules):
    # one un-sliced call, the sliced schedule without the collective, and the full step measured above
    attribution = None
    if world > 1:
        def single_call():
            ex.engine.extract_device(imgs, masks, out, status)
        t_single = timed(single_call, args.steps) / args.steps
        og.skip_collective = True
        t_sliced = timed(step, args.steps) / args.steps
        og.skip_collective = False
        attribution = {"single_call_ms": t_single, "sliced_without_collective_ms": t_sliced,
                       "sliced_with_collective_ms": ms / args.steps,
                       "note": "max over ranks each: rank skew shows in single_call_ms vs the 1-GPU run, the slicing of the "
                               "shard in the second figure, the exposed part of the all-gather (and its SM / memory "
                               "interference) in the third"}
    # per-rank kernel time (the scaling tail: is it rank skew or the collective?)
    rank_kms = torch.zeros(world, dtype=torch.float64, device=dev)
    rank_kms[rank] = kms
    if world > 1:
        dist.all_reduce(rank_kms)

    # ---- the discretise / histogram / first-order stage on its own (class mask: firstorder): the one stage of the path
    # that is HBM-shaped (north_star: "achieved HBM GB/s ... for the discretise/first-order stage")
    stage1 = None
    if world == 1:
        ex1 = pkg.RadiomicsExtractor({"setting": st, "featureClass": {"firstorder": []}}, device=local)
        o1 = torch.empty((B, ex1.engine.F), dtype=torch.float64, device=dev)
        for _ in range(3):
            ex1.engine.extract_device(imgs, masks, o1, status)
        m1 = timed(lambda: ex1.engine.extract_device(imgs, masks, o1, status), args.steps) / args.steps
        b1 = bytes_per_patch(H, H, ex1.engine.F)
        same_cols = bool(torch.equal(o1, out[:, :ex1.engine.F])) if ex.engine.names[:18] == ex1.engine.names else None
        stage1 = {"classes": "firstorder only (18 features): TMA staging, ROI histogram, bin edges, level histogram, fp64 reductions",
                  "ms_per_step": m1, "patches_per_s": B / (m1 / 1e3), "bytes_per_patch": b1,
                  "achieved_gbs": B * b1 / (m1 / 1e3) / 1e9, "equals_full_pass_columns": same_cols}
        ex.engine.extract_device(imgs, masks, out, status)  # restore status of the full pass
        torch.cuda.synchronize()
        del ex1, o1

    alt = {}
    if not args.no_alt:
        # callers that already hold bit-packed masks (radb_pack_masks_host's layout) skip the host packing entirely
        h_pk = torch.empty((B, ex.engine.packed_stride(H, H)), dtype=torch.uint8).pin_memory()
        ex.engine.pack_masks_host(h_msk, h_pk.view(-1), 8)
        h_out2 = torch.empty((B, F), dtype=torch.float64).pin_memory()
        for _ in range(2):
            ex.pipeline.run(h_img, h_pk, h_out2, h_st, masks_packed=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ex.pipeline.run(h_img, h_pk, h_out2, h_st, masks_packed=True)
        barrier()
        tp = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        alt["e2e_masks_already_packed"] = {"value": world * B * args.steps / float(tp.item()), "unit": UNIT,
                                           "h2d_bytes_per_step": int(world * ex.pipeline.h2d_bytes),
                                           "rows_equal": bool(torch.equal(h_out2, h_out))}
        del h_pk, h_out2
    if not args.no_alt and world == 1:
        for name, s2 in (("binWidth10_inplane", {"label": 255, "binWidth": 10.0, "force2D": False}),
                         ("binWidth10_literal_force2D", {"label": 255, "binWidth": 10.0, "force2D": True})):
            ex2 = pkg.RadiomicsExtractor({"setting": s2}, device=local)
            o2 = torch.empty((B, ex2.engine.F), dtype=torch.float64, device=dev)
            for _ in range(2):
                ex2.engine.extract_device(imgs, masks, o2, status)
            m2 = timed(lambda: ex2.engine.extract_device(imgs, masks, o2, status), 3) / 3
            alt[name] = {"value": B / (m2 / 1e3), "unit": UNIT, "ms_per_step": m2}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        arm = CpuArm(st)
        n_pool = min(B, 8192)
        ci, cm = h_img[:n_pool].numpy(), h_msk[:n_pool].numpy()
        n = arm.calibrate(ci, cm, args.cpu_seconds)
        ref, dt = arm.run(ci[:n], cm[:n])
        arm.close()
        got = out[:n].cpu().numpy()
        ok = bool(np.allclose(got, ref, rtol=1e-6, atol=1e-9, equal_nan=True))
        cpu = {"value": n / dt, "unit": UNIT, "cores": arm.cores, "kind": "port",
               "sample": "first %d patches of the workload, oracle (C matrices + NumPy features) on a %d-process pool; "
                         "GPU rows match it within rtol 1e-6/atol 1e-9: %s" % (n, arm.cores, ok)}

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        bpp = bytes_per_patch(H, H, F)
        achieved = B * bpp / (kms / 1e3) / 1e9
        dom = max(kparts, key=kparts.get)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh)["dram_bytes_per_patch"] * B  # ncu capture, scaled to the patches of one pass
        except Exception:
            pass
        total = world * B * args.steps
        line = {
            "metric": METRIC, "value": total / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, B, F),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "radb_build_kernel + radb_angle_lane_kernel + radb_misc_lane_kernel (+ warp-level "
                                                       "radb_misc_kernel residual); one pass of the hot path, chunks of 65536 patches",
                         "kernel_ms": kms, "kernel_ms_parts": kparts, "dominant_kernel": "radb_%s_kernel" % dom,
                         "dominant_share": kparts[dom] / kms,
                         "dominant_achieved": B * bpp / (kparts[dom] / 1e3) / 1e9,
                         "bytes_per_patch": bpp, "peak_source": peak_src,
                         "stage1_firstorder_only": None if stage1 is None else dict(stage1, frac=stage1["achieved_gbs"] / peak),
                         "note": "algorithmic bytes = H*W px + H*W mask + 8*F per patch over the summed duration of "
                                 "the kernels of a pass (build / angle reductions / misc reductions); dominant_achieved uses the dominant kernel's duration "
                                 "alone. The pass is issue/latency bound (shared-memory atomics, fp64), not HBM "
                                 "bound (DESIGN.md, profiles/)"},
            "cpu_baseline": cpu,
            "e2e": {"value": total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(world * e2e_stats["h2d"]),
                    "d2h_bytes_per_step": int(world * B * (F * 8 + 4)), "matches_device_path": same,
                    "h2d_gbs_per_rank": e2e_stats["h2d"] * args.steps / e2e_s / 1e9,
                    "masks": "uint8 masks handed over (as the reference does); %d of %d chunks per step packed to 1 bit per "
                             "pixel by %d host threads%s and read packed by the kernels (radb_extract_packed)"
                             % (e2e_stats["packed"], e2e_stats["chunks"], e2e_stats["threads"],
                                " (adaptive: only while the host keeps ahead of the link)" if e2e_stats["adaptive"] else ""),
                    "numa": numa},
            "gpu_launches": int(launches),
            "multi_gpu": None if world == 1 else {
                "gathered_equals_single_gpu_rows": gather_ok,
                "kernel_ms_per_rank": [round(float(x), 4) for x in rank_kms.cpu().tolist()],
                "attribution": attribution,
                "note": "gathered (device-resident step) and e2e_gathered (host-to-host step) compared bit for bit on rank 0 "
                        "with a single-GPU recomputation of every rank's shard, outside the timed region"},
            "clocks": sampler.summary(),
            "invalid_rows": bad,
            "alt": alt,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
