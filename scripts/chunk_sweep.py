#!/usr/bin/env python
"""Chunk-size sweep of the device-resident pass (records in L2 vs HBM): patches per chunk vs ms per 100 k patches,
with the two-stream build / reduce pipeline.  python scripts/chunk_sweep.py [binWidth]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import multimodal_isic_b200 as pkg  # noqa: E402


def main():
    bw = float(sys.argv[1]) if len(sys.argv) > 1 else 25.0
    B = 100000
    imgs, masks = pkg.synth.make_patches_torch(B, 64, seed=1234, device="cuda")
    eng = pkg.Engine(bw, 255, pkg.in_plane_angles())
    out = torch.empty((B, eng.F), dtype=torch.float64, device="cuda")
    st = torch.empty((B,), dtype=torch.int32, device="cuda")
    for chunk in (2048, 4096, 6144, 8192, 12288, 16384, 32768, 65536):
        eng.set_chunk(chunk)
        for _ in range(3):
            eng.extract_device(imgs, masks, out, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.extract_device(imgs, masks, out, st)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"binWidth": bw, "chunk": chunk, "ms_per_100k": e0.elapsed_time(e1) / 5,
                          "patches_per_s": B / (e0.elapsed_time(e1) / 5e3)}), flush=True)


if __name__ == "__main__":
    main()
