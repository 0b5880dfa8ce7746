"""Cases run under AddressSanitizer by tests/test_emu_kernel.py::test_emu_address_sanitizer (compute-sanitizer is
closed on the GPU pool): narrow / wide mode, every pixel type, shape2D, edge cases."""
import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from tests.emu_runner import EmuRunner, edge_case_batch, compare_with_oracle
from multimodal_isic_b200 import synth
from oracle import radiomics_oracle as orc
emu=EmuRunner(sys.argv[1])
ALL=("shape2D",)+tuple(orc.CLASS_ORDER)
INPLANE=orc.angles(2)[0]; LIT=orc.angles(2,force2D=True)[0]
for (H,W,n,ang) in ((64,64,2,INPLANE),(37,53,2,LIT),(120,150,1,INPLANE),(270,300,1,INPLANE),(121,151,3,INPLANE)):  # last: wide mode, odd pixel count, several patches
    im,mk=synth.make_patches(n,H,W,seed=1)
    r=emu.run(im,mk,10,255,ang,classes=ALL); print(H,W,'ok',r['status'])
im,mk=edge_case_batch(); r=emu.run(im,mk,10,255,INPLANE,classes=ALL); print('edge',r['status'])
g,mk=synth.make_patches(2,40,36,seed=6); f=(np.sqrt(g.astype(np.float64))*11.3-40.0)
r=emu.run(f,mk,7.5,255,INPLANE,max_ng=40); print('f64',r['status'])
r=emu.run(f.astype(np.float32),mk,7.5,255,INPLANE,max_ng=40); print('f32',r['status'])
r=emu.run((g.astype(np.uint16)*7),mk,64,255,INPLANE,max_ng=40); print('u16',r['status'])
# thread-level reduction kernels with the in-thread MCC (binWidth 25 -> Ng <= 11), ragged batch, 256 gray levels (big mode)
im,mk=synth.make_patches(5,40,44,seed=7); r=emu.run(im,mk,25,255,INPLANE,classes=ALL); print('lane',r['status'])
imgs=[im[0],im[1][:33,:29].copy(),im[2],synth.make_patches(1,20,20,seed=8)[0][0]]; mks=[mk[0],mk[1][:33,:29].copy(),mk[2],synth.make_patches(1,20,20,seed=8)[1][0]]
out,st=emu.run_ragged(imgs,mks,25,255,INPLANE); print('ragged',st)
g16,m16=synth.make_patches(1,20,20,seed=9,dtype=np.uint16,vmax=2047); r=emu.run(g16,m16,8,255,INPLANE,max_ng=256); print('big',r['status'])
