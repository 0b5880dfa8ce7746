#!/usr/bin/env python
"""Per-phase clock64() timing of the build kernel (experiment tool, not part of the product).

Makes an instrumented copy of multimodal-isic_b200/csrc under build/clk/: thread 0 of every CTA stores clock64()
after each phase barrier into a side buffer, and the patched launch() prints the average cycles per phase when
RADB_CLK is set.  The instrumented library is slower than the product (extra barriers); the SHARES are what it is for.

  python scripts/phase_clocks.py            # here: writes + compiles build/clk/libradb_clk.so
  gpurun -- 'RADB_CLK=1 RADB_LIB=build/clk/libradb_clk.so python scripts/profile_kernel.py --patches 32768 --launches 2'
"""
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "multimodal-isic_b200", "csrc")
DST = os.path.join(ROOT, "build", "clk")
os.makedirs(DST, exist_ok=True)
for f in os.listdir(SRC):
    if f.endswith((".h", ".cuh", ".cu", ".cpp")):
        shutil.copy(os.path.join(SRC, f), os.path.join(DST, f))

# ---- kernel: clock marks after the barrier that closes each phase
lines = open(os.path.join(SRC, "radb_kernels.cuh")).read().split("\n")
sync = [i for i, l in enumerate(lines) if l == "    __syncthreads();"]  # function-level barriers only
ph = {}
for i, l in enumerate(lines):
    m = re.match(r"\s*// ---- (phase [0-9a-z]+)", l)
    if m:
        ph[m.group(1)] = i
before = lambda i: max(j for j in sync if j < i)
after = lambda i: min(j for j in sync if j > i)
between = [j for j in sync if ph["phase 1"] < j < ph["phase 2"]]
marks = [before(ph["phase 1"]), between[0], between[-1], before(ph["phase 3a"]), before(ph["phase 3b"]),
         before(ph["phase 4"]), before(ph["phase 5"]), after(ph["phase 5"])]
out = []
for i, l in enumerate(lines):
    out.append(l)
    if i in marks:
        out.append("    if (tid == 0 && p.clk) p.clk[patch * 16 + %d] = clock64();" % (marks.index(i) + 1))
s = "\n".join(out)
split = "    if (glcm_pad) {\n        // the padded counters become the record's compact"
assert split in s
s = s.replace(split, "    __syncthreads();\n    if (tid == 0 && p.clk) p.clk[patch * 16 + 12] = clock64();\n" + split, 1)
s = s.replace("    // ---- phase 0: stage the patch, zero the counters",
              "    if (tid == 0 && p.clk) p.clk[patch * 16 + 0] = clock64();\n    // ---- phase 0: stage the patch, zero the counters", 1)
tail = "            if (i < g0 || i >= g1) dst[i] = src[i];\n    }\n}"
assert tail in s
s = s.replace(tail, tail[:-1] + "    __syncthreads();\n    if (tid == 0 && p.clk) p.clk[patch * 16 + 9] = clock64();\n}", 1)
# experiment switches (RADB_SKIP bits): 1 = no size atomics in the fold, 2 = no pointer jumping in the fold
x1 = "        atomicAdd(&lab[r], (UW)(len - own) << US);"
x2 = "        unsigned r;\n        while (true) {\n            const unsigned pa = (unsigned)(ws & ULO);"
assert x1 in s and x2 in s
x3 = "radb_red_add_if(endp && g, w, val);"
x4 = "atomicAdd(has ? &ngn_p[c * NB + cnt] : trash, num);"
assert s.count(x3) == 2 and x4 in s
s = s.replace(x3, "radb_red_add_if(endp && g && !(p.dbg_skip & 4), w, val);")   # 4 = no GLRLM updates in the line walks
s = s.replace(x4, "if (!(p.dbg_skip & 8)) atomicAdd(has ? &ngn_p[c * NB + cnt] : trash, num);")  # 8 = no NGTDM numerator updates
x5 = "    for (int k = tid; k < nruns; k += RADB_NTB) runs[k] = (unsigned short)fold_run(runs[k]);"
assert x5 in s
s = s.replace(x5, x5.replace("k < nruns", "k < ((p.dbg_skip & 16) ? 0 : nruns)"))   # 16 = no fold at all (timing floor of the phase)
s = s.replace(x1, "        if (!(p.dbg_skip & 1)) atomicAdd(&lab[r], (UW)(len - own) << US);", 1)
s = s.replace(x2, "        unsigned r = st;\n        while (!(p.dbg_skip & 2)) {\n            const unsigned pa = (unsigned)(ws & ULO);", 1)
open(os.path.join(DST, "radb_kernels.cuh"), "w").write(s)

# ---- params: the side buffer
s = open(os.path.join(SRC, "radb_params.h")).read()
anchor = "    unsigned char* ws;        // [B][rec_bytes]"
assert anchor in s
open(os.path.join(DST, "radb_params.h"), "w").write(s.replace(anchor, "    long long* clk;\n    int dbg_skip;\n" + anchor))

# ---- launch(): allocate, pass, dump
s = open(os.path.join(SRC, "radb_api.cu")).read()
a1 = "    cudaStream_t rs = piped ? h->red_stream : st;"
a2 = "        q.B = n;\n"
a3 = "    e = cudaGetLastError();\n    if (e != cudaSuccess) return cuda_fail(e, \"radb kernel launch\");"
assert a1 in s and a2 in s and a3 in s
s = s.replace(a1, a1 + "\n    static long long* g_clk = nullptr;\n    if (!g_clk) cudaMalloc(&g_clk, (size_t)200000 * 16 * 8);\n"
                       "    cudaMemsetAsync(g_clk, 0, (size_t)200000 * 16 * 8, st);", 1)
s = s.replace(a2, a2 + "        q.clk = (n <= 200000) ? g_clk : nullptr;\n        q.dbg_skip = getenv(\"RADB_SKIP\") ? atoi(getenv(\"RADB_SKIP\")) : 0;\n", 1)
dump = r'''    if (getenv("RADB_CLK")) {
        cudaDeviceSynchronize();
        const long long nb = p.B < chunk ? p.B : chunk;
        std::vector<long long> hc((size_t)nb * 16);
        cudaMemcpy(hc.data(), g_clk, hc.size() * 8, cudaMemcpyDeviceToHost);
        double acc[10] = {0};
        long long cnt = 0;
        for (long long i = 0; i < nb; i++) {
            const long long* c = &hc[i * 16];
            if (!c[0] || !c[9]) continue;
            cnt++;
            for (int k = 1; k <= 9; k++) acc[k] += (double)(c[k] - c[k - 1]);
        }
        static const char* nm[10] = {"", "p0 stage+zero", "p1 hist/bbox", "validity+lut", "p2 discretise", "p3a walks",
                                     "p3b neighbourhood", "p4 fold+glcm out", "p5 zones", "p6 record copy"};
        {
            double f = 0;
            long long c2 = 0;
            for (long long i = 0; i < nb; i++) {
                const long long* c = &hc[i * 16];
                if (!c[0] || !c[9] || !c[12]) continue;
                f += (double)(c[12] - c[6]);
                c2++;
            }
            if (c2) fprintf(stderr, "  (p4 split: fold %.0f cycles, GLCM write-out = rest)\n", f / c2);
        }
        double tot = 0;
        for (int k = 1; k <= 9; k++) tot += acc[k];
        fprintf(stderr, "phase clocks (avg cycles per CTA over %lld patches, total %.0f):\n", cnt, tot / cnt);
        for (int k = 1; k <= 9; k++) fprintf(stderr, "  %-20s %8.0f  %5.1f%%\n", nm[k], acc[k] / cnt, 100 * acc[k] / tot);
    }
'''
s = s.replace(a3, dump + a3, 1)
open(os.path.join(DST, "radb_api.cu"), "w").write(s)
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
                       "-shared", "-o", "libradb_clk.so", "radb_api.cu", "radb_hostpack.cpp"], cwd=DST)
print("built", os.path.join(DST, "libradb_clk.so"))
