#!/usr/bin/env python
"""Small cases for compute-sanitizer (memcheck / racecheck): narrow + wide mode, shape2D, uint8 + float64,
edge cases.  Run plain first, then `compute-sanitizer --tool memcheck python scripts/sanitize_case.py`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import multimodal_isic_b200 as pkg  # noqa: E402
from tests.emu_runner import edge_case_batch  # noqa: E402

classes = ("shape2D",) + tuple(pkg.CLASS_ORDER)
ang4 = [(1, 1), (0, 1), (-1, 1), (1, 0)]
for (H, W, n) in ((64, 64, 6), (37, 53, 3), (130, 150, 2), (300, 280, 1)):
    imgs, masks = pkg.synth.make_patches(n, H, W, seed=1)
    eng = pkg.Engine(10, 255, ang4, classes=classes)
    out, st = eng.extract_device(torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda())
    torch.cuda.synchronize()
    print(H, W, "ok", float(out[0, 0]), int(st.sum()))
imgs, masks = edge_case_batch()
eng = pkg.Engine(10, 255, [(0, 1)], classes=classes)
out, st = eng.extract_device(torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda())
torch.cuda.synchronize()
print("edge", st.tolist())
f = torch.as_tensor(np.sqrt(pkg.synth.make_patches(3, 48, 48, seed=2)[0].astype(np.float64)) * 9.0).cuda()
m = torch.as_tensor(pkg.synth.make_patches(3, 48, 48, seed=2)[1]).cuda()
eng = pkg.Engine(5.0, 255, ang4, max_ng=40)
out, st = eng.extract_device(f, m)
torch.cuda.synchronize()
print("f64 ok", int(st.sum()))
bgr = torch.randint(0, 256, (2, 40, 44, 3), dtype=torch.uint8, device="cuda")
eng = pkg.Engine(10, 255, [(0, 1)], classes=classes)
out, st = eng.extract_bgr(bgr, torch.as_tensor(pkg.synth.make_patches(2, 40, 44, seed=3)[1]).cuda())
torch.cuda.synchronize()
print("bgr ok", int(st.sum()))
# thread-level reduction kernels (binWidth 25: Ng <= 11, MCC in the thread), multi-chunk two-stream pipeline
imgs, masks = pkg.synth.make_patches(45, 64, 64, seed=4)
eng = pkg.Engine(25, 255, ang4, classes=classes)
eng.set_chunk(12)
out, st = eng.extract_device(torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda())
torch.cuda.synchronize()
print("lane + pipeline ok", int(st.sum()))
# ragged batch (mixed sizes, odd offsets), packed-mask transfer
lst_i = [imgs[0], imgs[1][:33, :29].copy(), imgs[2], pkg.synth.make_patches(1, 128, 128, seed=5)[0][0]]
lst_m = [masks[0], masks[1][:33, :29].copy(), masks[2], pkg.synth.make_patches(1, 128, 128, seed=5)[1][0]]
ip, mp, io, mo, hw = pkg.pack_ragged(lst_i, lst_m)
out, st = eng.extract_ragged(torch.as_tensor(ip).cuda(), torch.as_tensor(mp).cuda(), io, mo, hw)
torch.cuda.synchronize()
print("ragged ok", st.tolist())
o2, s2 = pkg.HostPipeline(eng, chunk=16, pack_masks=True, pack_threads=2).run(imgs, masks)
print("packed masks ok", int(s2.sum()))
# 256 gray levels: big mode (GLCM + MCC workspace in global memory, u16 level image)
g, m = pkg.synth.make_patches(2, 32, 32, seed=6, dtype=np.uint16, vmax=2047)
eng = pkg.Engine(8, 255, ang4, max_ng=256)
out, st = eng.extract_device(torch.as_tensor(g.view(np.int16)).cuda().view(torch.uint16), torch.as_tensor(m).cuda())
torch.cuda.synchronize()
print("big mode ok", int(st.sum()))
