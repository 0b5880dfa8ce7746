// radb -- B200-native radiomic feature kernels (sm_100a).  A pass over a chunk of patches is a build
// kernel followed by small reduction kernels, so that every kernel's code stays small enough for the SM
// instruction caches (a fused single kernel measured 55 % "no instruction" stalls, profiles/):
//   radb_build_kernel : one CTA per patch.  TMA bulk copy (cp.async.bulk + mbarrier) of the raw patch + mask
//                       into shared memory -> ROI histogram / bbox -> value->level LUT -> padded level image
//                       -> line walks (GLRLM runs + row-run labels) -> one neighbourhood pass (GLCM, GLDM,
//                       NGTDM, run-adjacency unions for GLSZM) -> zone sizes.  Every matrix is a
//                       shared-memory privatised integer counter array; the finished RECORD (header +
//                       matrices) is copied to a global workspace.
//   radb_angle_lane_kernel / radb_mcc_g8_kernel / radb_misc_lane_kernel (radb_lane.cuh): thread-level fp64
//                       reductions of the records -- one thread per (patch, angle) for GLCM / GLRLM / MCC, one
//                       thread per patch for first-order / GLDM / NGTDM / GLSZM; the eigenproblems of 15-40
//                       gray levels on one warp per patch, 8 lanes per angle.
//   radb_angle_kernel : warp a reduces GLRLM angle a (16 features) and GLCM angle a (24 features, MCC through
//                       Householder tridiagonalisation + Sturm multisection), then the CTA forms the nanmean
//                       over angles.  Kept for asymmetric GLCMs and more than 40 gray levels.
//   radb_misc_kernel  : one warp each for GLSZM, GLDM, NGTDM and first-order features.  Kept for long GLSZM
//                       overflow lists and many gray levels.
// Semantics follow pyradiomics 3.1.0 as called from /root/reference/RadiomicExtractor.py:38-48
// (settings /root/reference/params.yml:93-119); the algorithm restated is SURVEY.md Appendix A.
// This header also compiles as plain C++ under tests/emu/cuda_emu.h (RADB_EMU) so that the same
// source is exercised against the oracle on a GPU-less box.  It is not a CPU fallback: the
// product (radb_api.cu) only ever launches it on the device.
#pragma once
#include "radb_params.h"
#ifndef RADB_EMU
#include <cuda_runtime.h>
#endif
#include <math.h>

#define RADB_EPS 2.220446049250313e-16
#define RADB_EPS_LN2 3.203426503814917e-16   // eps / ln(2): first-order term of log2(x + eps)
#define FULLMASK 0xffffffffu

// ------------------------------------------------------------------ warp helpers
__device__ __forceinline__ double warp_sum(double v)
{
#ifdef RADB_EMU
    double all[32];
    emu::gather<double>(v, all);
    for (int m = 16; m >= 1; m >>= 1)
        for (int i = 0; i < 32; i++) all[i] = (i & m) ? all[i] : all[i] + all[i ^ m];
    return all[0];
#else
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(FULLMASK, v, m);
    return v;
#endif
}
// N sums at once through a per-warp shared-memory scratch (N x 33 doubles): every lane stores its
// N partials, lane i < N adds the 32 partials of value i in lane order, all lanes read the N
// totals back.  Compared with N shuffle butterflies this is ~3x fewer instructions, one short
// dependent chain instead of N, and a fraction of the code size (the inlined butterflies were
// 25-40 % of the reduction kernels' instruction footprint).  Fixed order => reproducible.
#define RADB_RED_MAX 14
#define RADB_RED_DOUBLES (RADB_RED_MAX * 33)
template <int N>
__device__ __forceinline__ void warp_sum_n(double (&v)[N], double* scr, int lane)
{
#pragma unroll
    for (int i = 0; i < N; i++) scr[i * 33 + lane] = v[i];
    __syncwarp();
    if (lane < N) {
        const double* r = scr + lane * 33;
        double s0 = 0, s1 = 0;
#pragma unroll 1
        for (int j = 0; j < 32; j += 2) { s0 += r[j]; s1 += r[j + 1]; }
        scr[lane * 33 + 32] = s0 + s1;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = scr[i * 33 + 32];
    __syncwarp();
}
__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#ifdef RADB_EMU
    long long all[32];
    emu::gather<long long>(v, all);
    long long s = 0;
    for (int i = 0; i < 32; i++) s += all[i];
    return s;
#else
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(FULLMASK, v, m);
    return v;
#endif
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#ifdef RADB_EMU
    return (int)warp_sum_ll(v);
#else
    return __reduce_add_sync(FULLMASK, v);
#endif
}
__device__ __forceinline__ int warp_max_i(int v)
{
#ifdef RADB_EMU
    int all[32];
    emu::gather<int>(v, all);
    int s = all[0];
    for (int i = 1; i < 32; i++) s = all[i] > s ? all[i] : s;
    return s;
#else
    return __reduce_max_sync(FULLMASK, v);
#endif
}
__device__ __forceinline__ int warp_min_i(int v) { return -warp_max_i(-v); }
// segmented sum: lanes are split into groups of `w` consecutive lanes (w = 8, 16 or 32, warp-uniform);
// every lane gets the sum of its group
__device__ __forceinline__ int group_sum_i(int v, int w, int lane)
{
#ifdef RADB_EMU
    int all[32];
    emu::gather<int>(v, all);
    int s = 0;
    for (int i = lane - (lane & (w - 1)), e = i + w; i < e; i++) s += all[i];
    return s;
#else
    (void)lane;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const int t = __shfl_xor_sync(FULLMASK, v, m);
        if (m < w) v += t;
    }
    return v;
#endif
}
// exclusive prefix sum over lanes
__device__ __forceinline__ int warp_excl_scan_i(int v, int lane)
{
#ifdef RADB_EMU
    int all[32];
    emu::gather<int>(v, all);
    int s = 0;
    for (int i = 0; i < lane; i++) s += all[i];
    return s;
#else
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(FULLMASK, x, d);
        if (lane >= d) x += y;
    }
    return x - v;
#endif
}

__device__ __forceinline__ double nan_f64()
{
    unsigned long long b = 0x7ff8000000000000ULL;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

// ------------------------------------------------------------------ staging (TMA bulk copy)
#ifndef RADB_EMU
__device__ __forceinline__ unsigned smem_u32(const void* p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(void* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, void* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
#endif

// ------------------------------------------------------------------ union-find over row runs
// One word per pixel: low half = parent pixel index, high half = size.  During the union pass the
// high half of a run start holds its run length (static); after it, the run lengths are folded
// into the zone root with one native atomic per run.  Narrow mode (patches that fit shared
// memory, <= 65535 pixels): 32-bit words (16 | 16) in shared memory.  Wide mode (whole images):
// 64-bit words (32 | 32) in the global scratch.
template <bool WIDE> struct UF { typedef unsigned W; enum { S = 16 }; };
template <> struct UF<true> { typedef unsigned long long W; enum { S = 32 }; };

template <typename W, int S>
__device__ __forceinline__ unsigned uf_find(volatile W* L, unsigned x)
{
    const W lo = (((W)1) << S) - 1;
    W w = L[x];
    unsigned p = (unsigned)(w & lo);
    while (p != x) {
        const W wp = L[p];
        const unsigned g = (unsigned)(wp & lo);
        if (g != p) L[x] = (w & ~lo) | (W)g;  // path halving: x is not a root, g is an ancestor
        x = p;
        w = wp;
        p = g;
    }
    return x;
}
template <typename W, int S>
__device__ __forceinline__ unsigned uf_find_ro(const volatile W* L, unsigned x)
{
    const W lo = (((W)1) << S) - 1;
    unsigned p;
    while ((p = (unsigned)(L[x] & lo)) != x) x = p;
    return x;
}
template <typename W, int S>
__device__ __forceinline__ void uf_union(W* L, unsigned a, unsigned b)
{
    const W lo = (((W)1) << S) - 1;
    while (true) {
        a = uf_find<W, S>(L, a);
        b = uf_find<W, S>(L, b);
        if (a == b) return;
        if (a < b) { unsigned t = a; a = b; b = t; }
        const W wa = ((volatile W*)L)[a];
        if ((unsigned)(wa & lo) != a) continue;  // lost the race: a is no longer a root
        if (atomicCAS(&L[a], wa, (wa & ~lo) | (W)b) == wa) return;
    }
}

// 32-bit shared-memory addresses (line walks of the build kernel): address arithmetic in one register, the
// update as a plain `red.shared` (the CPU emulation uses ordinary pointers)
#ifdef RADB_EMU
typedef uintptr_t radb_saddr;
__device__ __forceinline__ radb_saddr radb_to_saddr(const void* p) { return (radb_saddr)p; }
__device__ __forceinline__ void radb_red_add_if(bool on, radb_saddr a, unsigned v) { if (on) atomicAdd((unsigned*)a, v); }
#else
typedef unsigned radb_saddr;
__device__ __forceinline__ radb_saddr radb_to_saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// predicated (not branched-around, not redirected to a scratch word: lanes that are off do not touch shared memory)
__device__ __forceinline__ void radb_red_add_if(bool on, radb_saddr a, unsigned v)
{
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p red.shared.add.u32 [%0], %1; }" ::"r"(a), "r"(v), "r"((unsigned)on) : "memory");
}
#endif

// GLRLM counters: packed u16 pairs updated with one 32-bit atomic (narrow) or plain u32 (wide)
__device__ __forceinline__ void add_u16(unsigned* base, int cell)
{
    atomicAdd(&base[cell >> 1], (cell & 1) ? 0x10000u : 1u);
}
__device__ __forceinline__ int get_u16(const unsigned* base, int cell)
{
    return (int)((base[cell >> 1] >> ((cell & 1) * 16)) & 0xffffu);
}
template <bool WIDE>
__device__ __forceinline__ void add_run(unsigned* base, int cell)
{
    if (WIDE) atomicAdd(&base[cell], 1u);
    else add_u16(base, cell);
}
__device__ __forceinline__ int get_run(const unsigned* base, int cell, int wide)
{
    return wide ? (int)base[cell] : get_u16(base, cell);
}

// sum of the four bytes of a word (IDP.4A)
__device__ __forceinline__ unsigned bytes_sum4(unsigned w, unsigned acc)
{
#ifdef RADB_EMU
    return acc + (w & 0xffu) + ((w >> 8) & 0xffu) + ((w >> 16) & 0xffu) + (w >> 24);
#else
    return __dp4a(w, 0x01010101u, acc);
#endif
}
// dot product of the four bytes of w with the four bytes of m, plus acc (IDP.4A)
__device__ __forceinline__ unsigned bytes_dot4(unsigned w, unsigned m, unsigned acc)
{
#ifdef RADB_EMU
    for (int k = 0; k < 4; k++) acc += ((w >> (8 * k)) & 0xffu) * ((m >> (8 * k)) & 0xffu);
    return acc;
#else
    return __dp4a(w, m, acc);
#endif
}
// 0x80 in every byte position where the byte of w is non-zero
__device__ __forceinline__ unsigned bytes_nz4(unsigned w)
{
    return (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;
}

// per-byte equality of two packed 4-byte words: 0x80 in every byte position where a == b (exact)
__device__ __forceinline__ unsigned bytes_eq4(unsigned a, unsigned b)
{
    const unsigned t = a ^ b;
    return ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) & 0x80808080u;
}
__device__ __forceinline__ unsigned radb_roi4(const unsigned char* m, int q, unsigned label4, int packed)
{
    if (packed) {  // nibble q of the bit stream -> one bit per byte (multiply spreads bit k to bit 8k), as 0x80 flags
        const unsigned nib = ((unsigned)m[q >> 1] >> ((q & 1) * 4)) & 0xfu;
        return ((nib * 0x00204081u) & 0x01010101u) << 7;
    }
    return bytes_eq4(((const unsigned*)m)[q], label4);
}

// Python-style modulo (sign of the divisor), as numpy's % in imageoperations.getBinEdges
__device__ __forceinline__ double py_mod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}

// ROI membership of pixel i: byte masks compare with the label (imageoperations.getMask: mask == label); bit-packed
// masks (radb_extract_packed: bit i of the patch's stream, set = ROI) test the bit.  `packed` is kernel-uniform.
__device__ __forceinline__ bool radb_roi(const unsigned char* m, int i, int label, int packed)
{
    return packed ? (((unsigned)m[i >> 3] >> (i & 7)) & 1u) != 0u : ((int)m[i] == label);
}
// four consecutive pixels q*4 .. q*4+3 at once: 0x80 in byte k <=> pixel k is in the ROI
__device__ __forceinline__ unsigned radb_roi4(const unsigned char* m, int q, unsigned label4, int packed);

// dense batches: patch k is row k; ragged batches: the group-local patch k maps to output row rows[k]
__device__ __forceinline__ long long radb_row(const RadbParams& p, long long patch) { return p.rows ? p.rows[patch] : patch; }
__device__ __forceinline__ const unsigned char* radb_mask_ptr(const RadbParams& p, long long patch)
{
    return p.mask_off ? p.mask + p.mask_off[patch] : p.mask + (patch / p.mask_group) * p.mask_stride;
}

#include "radb_features.cuh"
#include "radb_lane.cuh"
#include "radb_lanczos.cuh"
#include "radb_filters.cuh"

// ------------------------------------------------------------------ first-order for non-uint8 pixels
// uint8 patches get all 18 first-order features from the 256-bin raw histogram (fo_task_u8, misc
// kernel).  For uint16 / float32 / float64 pixels the build kernel computes them from the ROI values
// themselves while the raw patch is still staged: fp64 moment passes with fixed-order CTA
// reductions, and the ten order statistics behind the 10/25/50/75/90 percentiles by a multi-rank
// radix select (8 key bits per pass over order-preserving integer keys).
template <typename PT> struct PixKey;
template <> struct PixKey<unsigned char> { enum { BITS = 8 }; };
template <> struct PixKey<unsigned short> {
    enum { BITS = 16 };
    static __device__ __forceinline__ unsigned long long key(unsigned short v) { return v; }
    static __device__ __forceinline__ double value(unsigned long long k) { return (double)k; }
};
template <> struct PixKey<float> {
    enum { BITS = 32 };
    static __device__ __forceinline__ unsigned long long key(float v)
    {
        unsigned b;
        memcpy(&b, &v, 4);
        return (b & 0x80000000u) ? (unsigned)~b : (b | 0x80000000u);
    }
    static __device__ __forceinline__ double value(unsigned long long k)
    {
        unsigned b = (unsigned)k;
        b = (b & 0x80000000u) ? (b & 0x7fffffffu) : ~b;
        float v;
        memcpy(&v, &b, 4);
        return (double)v;
    }
};
template <> struct PixKey<double> {
    enum { BITS = 64 };
    static __device__ __forceinline__ unsigned long long key(double v)
    {
        unsigned long long b;
        memcpy(&b, &v, 8);
        return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
    }
    static __device__ __forceinline__ double value(unsigned long long k)
    {
        unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
        double v;
        memcpy(&v, &b, 8);
        return v;
    }
};

// fixed-order CTA sum of K doubles per thread: warp butterflies, one slot per warp, everyone adds the slots
template <int K>
__device__ __forceinline__ void cta_sum(double (&v)[K], double* slots /*[K][RADB_NTB/32]*/, int tid)
{
    const int lane = tid & 31, warp = tid >> 5, NW = RADB_NTB / 32;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const double t = warp_sum(v[k]);
        if (lane == 0) slots[k * NW + warp] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) {
        double t = 0;
        for (int w = 0; w < NW; w++) t += slots[k * NW + w];
        v[k] = t;
    }
    __syncthreads();
}

#define RADB_FO_RANKS 10
// scratch layout (bytes from `scr`): slots double[8*8] | fr double[16] | prefix u64[10] | gprefix u64[10] |
//                                    rem int[10] | grp int[10] | ngroups int[2] | hist int[10][256]
#define RADB_FO_SCRATCH (64 * 8 + 16 * 8 + 10 * 8 + 10 * 8 + 10 * 4 + 10 * 4 + 8 + 10 * 256 * 4 + 64)

template <typename PT>
__device__ void fo_generic(const RadbParams& p, const PT* img, const unsigned char* msk, int HW, int N, double vmn,
                           double vmx, const int* lhist, int ng, unsigned char* scr, double* o, int tid)
{
    typedef PixKey<PT> KY;
    double* slots = (double*)scr;
    double* fr = slots + 64;                                  // [0..9] order statistics, [10..14] fractions
    unsigned long long* prefix = (unsigned long long*)(fr + 16);
    unsigned long long* gprefix = prefix + 10;
    int* rem = (int*)(gprefix + 10);
    int* grp = rem + 10;
    int* ngroups = grp + 10;
    int* hist = ngroups + 2;
    const double dN = (double)N, rN = 1.0 / dN, shift = p.shift;
    // ---- pass 1: mean; pass 2: central moments, MAD, energy
    double a1[2] = {0, 0};
    for (int i = tid; i < HW; i += RADB_NTB)
        if (radb_roi(msk, i, p.label, p.mask_bits)) { const double x = (double)img[i]; a1[0] += x; a1[1] += (x + shift) * (x + shift); }
    cta_sum(a1, slots, tid);
    // a flat ROI has mean == its value exactly (a rounded sum/N would leave spurious 1e-30 moments)
    const double mean = (vmn == vmx) ? vmn : a1[0] * rN, en = a1[1];
    double a2[4] = {0, 0, 0, 0};
    for (int i = tid; i < HW; i += RADB_NTB)
        if (radb_roi(msk, i, p.label, p.mask_bits)) {
            const double x = (double)img[i], d = x - mean, d2 = d * d;
            a2[0] += d2; a2[1] += d2 * d; a2[2] += d2 * d2; a2[3] += fabs(d);
        }
    cta_sum(a2, slots, tid);
    const double m2 = a2[0] * rN, m3 = a2[1] * rN, m4 = a2[2] * rN, mad = a2[3] * rN;
    // ---- multi-rank radix select: ranks lo/hi of the 10/25/50/75/90 percentiles (numpy 'linear')
    if (tid < 5) {
        const double qq = tid == 0 ? 0.1 : tid == 1 ? 0.25 : tid == 2 ? 0.5 : tid == 3 ? 0.75 : 0.9;
        const double pos = qq * (dN - 1.0), fl = floor(pos);
        int lo = (int)fl;
        lo = lo > N - 1 ? N - 1 : lo;
        const int hi = lo + 1 > N - 1 ? N - 1 : lo + 1;
        fr[10 + tid] = pos - fl;
        rem[2 * tid] = lo;
        rem[2 * tid + 1] = hi;
    }
    if (tid < RADB_FO_RANKS) { prefix[tid] = 0; grp[tid] = 0; }
    if (tid == 0) { ngroups[0] = 1; gprefix[0] = 0; }
    for (int i = tid; i < RADB_FO_RANKS * 256; i += RADB_NTB) hist[i] = 0;
    __syncthreads();
    for (int sh = KY::BITS - 8; sh >= 0; sh -= 8) {
        const int ngr = ngroups[0];
        const bool first = (sh == KY::BITS - 8);
        for (int i = tid; i < HW; i += RADB_NTB)
            if (radb_roi(msk, i, p.label, p.mask_bits)) {
                const unsigned long long k = KY::key(img[i]);
                const unsigned long long kh = first ? 0ULL : (k >> (sh + 8));
                const int b = (int)((k >> sh) & 255ULL);
                for (int g = 0; g < ngr; g++)
                    if (kh == gprefix[g]) atomicAdd(&hist[g * 256 + b], 1);
            }
        __syncthreads();
        if (tid < RADB_FO_RANKS) {
            const int* hg = hist + grp[tid] * 256;
            int cum = 0, bucket = 255;
            for (int b = 0; b < 256; b++) {
                const int c = hg[b];
                if (rem[tid] < cum + c) { bucket = b; break; }
                cum += c;
            }
            prefix[tid] = (prefix[tid] << 8) | (unsigned long long)bucket;
            rem[tid] -= cum;
        }
        __syncthreads();
        if (tid == 0) {  // ranks that still share a prefix share a histogram in the next pass
            int n = 0;
            for (int q = 0; q < RADB_FO_RANKS; q++) {
                int g = -1;
                for (int h = 0; h < n; h++)
                    if (gprefix[h] == prefix[q]) g = h;
                if (g < 0) { g = n; gprefix[n++] = prefix[q]; }
                grp[q] = g;
            }
            ngroups[0] = n;
        }
        for (int i = tid; i < RADB_FO_RANKS * 256; i += RADB_NTB) hist[i] = 0;
        __syncthreads();
    }
    if (tid < RADB_FO_RANKS) fr[tid] = KY::value(prefix[tid]);
    __syncthreads();
    double pc[5];
#pragma unroll
    for (int q = 0; q < 5; q++) pc[q] = fr[2 * q] + (fr[2 * q + 1] - fr[2 * q]) * fr[10 + q];
    const double p10 = pc[0], p25 = pc[1], med = pc[2], p75 = pc[3], p90 = pc[4];
    // ---- robust MAD: values inside [p10, p90]
    double a3[2] = {0, 0};
    for (int i = tid; i < HW; i += RADB_NTB)
        if (radb_roi(msk, i, p.label, p.mask_bits)) {
            const double x = (double)img[i];
            if (x >= p10 && x <= p90) { a3[0] += x; a3[1] += 1.0; }
        }
    cta_sum(a3, slots, tid);
    const double rin = 1.0 / a3[1], in_mean = a3[0] * rin;
    double a4[3] = {0, 0, 0};
    for (int i = tid; i < HW; i += RADB_NTB)
        if (radb_roi(msk, i, p.label, p.mask_bits)) {
            const double x = (double)img[i];
            if (x >= p10 && x <= p90) a4[0] += fabs(x - in_mean);
        }
    for (int i = tid; i < ng; i += RADB_NTB) {
        const double pi = (double)lhist[i] * rN;
        if (lhist[i]) a4[1] -= pi * radb_log2(pi + RADB_EPS);
        a4[2] += pi * pi;
    }
    cta_sum(a4, slots, tid);
    if (tid == 0) {
        o[0] = p10;
        o[1] = p90;
        o[2] = en;
        o[3] = a4[1];
        o[4] = p75 - p25;
        o[5] = (m2 == 0.0) ? 0.0 : m4 / (m2 * m2);
        o[6] = vmx;
        o[7] = mad;
        o[8] = mean;
        o[9] = med;
        o[10] = vmn;
        o[11] = vmx - vmn;
        o[12] = a4[0] * rin;
        o[13] = sqrt(en * rN);
        o[14] = (m2 == 0.0) ? 0.0 : m3 / (m2 * sqrt(m2));
        o[15] = en;  // TotalEnergy: pixel spacing is (1, 1)
        o[16] = a4[2];
        o[17] = m2;
    }
    __syncthreads();
}

// A.3 binImage: gray level of x = number of fp64 edges (low + k*bw, as numpy.arange builds them) <= x
// The edges are formed as NumPy forms them (numpy.arange / numpy.linspace: a rounded product, then a rounded
// sum): no fused multiply-add, or a pixel that sits exactly on an edge could land one level off.
__device__ __forceinline__ double radb_edge(double low, long long k, double bw) { return __dadd_rn(low, __dmul_rn((double)k, bw)); }
__device__ __noinline__ int radb_level(double x, double low, double bw)
{
    long long k = (long long)floor((x - low) / bw);
    while (radb_edge(low, k, bw) > x) k--;
    while (radb_edge(low, k + 1, bw) <= x) k++;
    return (int)(k + 1);
}

template <typename PT> struct fo_dispatch {
    static __device__ __forceinline__ void run(const RadbParams& p, const PT* img, const unsigned char* msk, int HW, int N,
                                               double vmn, double vmx, const int* lhist, int ng, unsigned char* scr,
                                               double* o, int tid)
    {
        fo_generic<PT>(p, img, msk, HW, N, vmn, vmx, lhist, ng, scr, o, tid);
    }
};
template <> struct fo_dispatch<unsigned char> {
    static __device__ __forceinline__ void run(const RadbParams&, const unsigned char*, const unsigned char*, int, int,
                                               double, double, const int*, int, unsigned char*, double*, int) {}
};

// level image element: u8, or u16 when the extractor is sized for more than 255 gray levels (big mode)
template <bool L16> struct LevT { typedef unsigned char T; };
template <> struct LevT<true> { typedef unsigned short T; };

// ------------------------------------------------------------------ build kernel: one CTA per patch
// FAST: the headline configuration as a compile-time specialisation -- uint8 pixels staged by TMA, narrow mode with the
// 4-pixel-word level image (1: with the run list, 2: without -- more gray levels, no room for it), integer binWidth, the four in-plane angles in canonical order, symmetric
// GLCM, alpha = 0, every texture class wanted, no debug output (radb_host.h: radb_fast_config).  The generic instance
// carries all the other paths (one pixel per thread, bbox scans, binCount, generic angle sets ...) as run-time
// branches: 7 000 SASS instructions, of which a CTA executes about half, against an instruction cache of 32 KB shared
// by five CTAs in five different phases (14 % of the warp stalls were "no instruction").  The specialisation drops
// them at compile time.
template <typename PT, bool DBG, bool WIDE, bool L16 = false, int FAST = 0>  // FAST: 0 generic | 1 with run list | 2 without
__device__ void radb_build_cta(const RadbParams& p, long long patch, unsigned char* smem)
{
    typedef typename LevT<L16>::T LT;
    typedef typename UF<WIDE>::W UW;  // union-find word
    const int US = UF<WIDE>::S;
    const UW ULO = (((UW)1) << US) - 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, W = p.W, HW = p.HW, WP = p.WP, XO = FAST ? 4 : p.xo, NA = FAST ? 4 : p.n_angles, NB = 2 * NA;
    const bool use_tma = FAST || p.use_tma, vec4 = FAST || p.vec4;
    const bool dense = FAST || (!WIDE && p.glrlm_dense != 0);  // GLRLM built as u32 [NA][max_ng][16] + global overflow
    const int DL = 16;                                         // (= p.glrlm_dense)
    const int LP = p.lp;  // pitch of the union-find array (pixel (y, x) <-> word y * LP + x)
    const PT* g_img = (const PT*)((const unsigned char*)p.img + (p.img_off ? p.img_off[patch] : patch * p.img_stride));
    const unsigned char* g_msk = radb_mask_ptr(p, patch);
    const long long row = radb_row(p, patch);
    unsigned char* g_rec = p.ws + patch * (long long)p.rec_bytes;              // this patch's record (global)
    unsigned char* g_scr = WIDE ? p.ws_scr + patch * p.scr_bytes : (unsigned char*)0;  // wide-mode scratch (global)
    // narrow: the raw patch is staged in shared memory and the level image / union-find words live
    // there too; wide: pixels are read straight from global memory and those arrays are global.
    const PT* s_img = WIDE ? g_img : (const PT*)(smem + p.o_stage);
    const unsigned char* s_msk = WIDE ? g_msk : smem + p.o_mask;
    UW* lab = WIDE ? (UW*)(g_scr + p.g_lab) : (UW*)(smem + p.o_stage);
    LT* lev = (LT*)(WIDE ? g_scr + p.g_lev : smem + p.o_lev);
    int* hist = (int*)(smem + p.o_hist);
    unsigned char* lut = smem + p.o_lut;
    int* lhist = (int*)(smem + p.o_lhist);
    int* glcm = (WIDE && p.big) ? (int*)(g_rec + (p.o_glcm - p.o_rec)) : (int*)(smem + p.o_glcm);
    int* gldm = (int*)(smem + p.o_gldm);
    int* ngc = (int*)(smem + p.o_ngc);
    int* ngn = (int*)(smem + p.o_ngn);
    int* szm = (int*)(smem + p.o_szm);
    unsigned* ovf = WIDE ? (unsigned*)(g_rec + (p.o_ovf - p.o_rec)) : (unsigned*)(smem + p.o_ovf);
    unsigned char* glrlm_base = WIDE ? g_rec + (p.o_glrlm - p.o_rec) : smem + p.o_glrlm;
    int* misc = (int*)(smem + p.o_misc);
    unsigned short* runs = (unsigned short*)(smem + (p.o_runs >= 0 ? p.o_runs : 0));  // row-run start pixels (narrow)
    const bool keep_runs = FAST == 1 || (!FAST && !WIDE && p.o_runs >= 0);
    double* out = p.out + row * (long long)p.F;

    // ---- phase 0: stage the patch, zero the counters
#ifndef RADB_EMU
    if (!WIDE && use_tma) {
        void* bar = smem + p.o_mbar;
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            const unsigned ib = (unsigned)(HW * sizeof(PT)), mb = (unsigned)(p.mask_bits ? (HW + 7) >> 3 : HW);
            mbar_expect_tx(bar, ib + mb);
            tma_load_1d(smem + p.o_stage, g_img, ib, bar);
            tma_load_1d(smem + p.o_mask, g_msk, mb, bar);
        }
    }
#endif
    {
        uint4* z = (uint4*)(smem + p.o_zero);
        const int nz = ((p.o_runs >= 0 ? p.o_runs : p.smem_total) - p.o_zero) / 16;  // (the run list is written before it is read)
        const uint4 zero = {0u, 0u, 0u, 0u};
        RADB_UNROLL(1)  // (code size: the kernel is instruction-fetch sensitive, profiles/)
        for (int i = tid; i < nz; i += RADB_NTB) z[i] = zero;
        if (WIDE) {  // level image + GLRLM counters in global memory
            uint4* zl = (uint4*)lev;
            const int nl = (int)(p.g_lab / 16);
            for (int i = tid; i < nl; i += RADB_NTB) zl[i] = zero;
            uint4* zr = (uint4*)glrlm_base;
            const int nrr = NA * p.glrlm_stride / 16;
            for (int i = tid; i < nrr; i += RADB_NTB) zr[i] = zero;
            if (p.big) {  // GLCM counters in the global record
                uint4* zg = (uint4*)glcm;
                const int ngc = (NA * p.max_ng * p.max_ng * 4 + 15) / 16;
                for (int i = tid; i < ngc; i += RADB_NTB) zg[i] = zero;
            }
        }
    }
    if (!WIDE) {
#ifndef RADB_EMU
        if (use_tma) {
            mbar_wait(smem + p.o_mbar, 0);
        } else
#endif
        {
            PT* d_img = (PT*)(smem + p.o_stage);
            unsigned char* d_msk = smem + p.o_mask;
            const int mbytes = p.mask_bits ? (HW + 7) >> 3 : HW;
            for (int i = tid; i < HW; i += RADB_NTB) {
                d_img[i] = g_img[i];
                if (i < mbytes) d_msk[i] = g_msk[i];
            }
        }
    }
    __syncthreads();

    // ---- phase 1: ROI histogram (uint8) or value range (other pixel types), bbox, voxel count
    const bool U8 = sizeof(PT) == 1;
    double* wrange = (double*)(smem + p.o_fo);  // non-uint8: per-warp min / max slots
    {
        int np = 0, ymin = H, ymax = -1, xmin = W, xmax = -1;
        double vmn = 1e308, vmx = -1e308;
        if (U8 && vec4) {
            // uint8 patches whose width is a multiple of 4: four pixels per 32-bit shared-memory load
            const int WQ = W >> 2, NQ = HW >> 2;
            const float inv_wq = 1.0f / (float)WQ;
            const unsigned l4 = (unsigned)(p.label & 0xff) * 0x01010101u;
            const unsigned* v4p = (const unsigned*)s_img;
            if (p.mask_bits || (p.label >= 0 && p.label <= 255))
                RADB_UNROLL(1)  // (code size: the kernel is instruction-fetch sensitive, profiles/)
                for (int q = tid; q < NQ; q += RADB_NTB) {
                    const unsigned eq = radb_roi4(s_msk, q, l4, p.mask_bits);
                    if (!eq) continue;
                    const int y = (int)(((float)q + 0.5f) * inv_wq), x0 = (q - y * WQ) << 2;
                    const unsigned v4 = v4p[q];
                    ymin = y < ymin ? y : ymin;
                    ymax = y > ymax ? y : ymax;
                    // branch-free: the pixel count and the x extent come from the bit positions of `eq` (0x80 per
                    // selected byte); unselected pixels increment a scratch word of the record header, so the four
                    // histogram updates are unconditional ATOMS.POPC.INC without BSSY / BRA / BSYNC around them
                    np += __popc(eq);
                    const int kmin = (__ffs((int)eq) - 1) >> 3, kmax = (31 - __clz((int)eq)) >> 3;
                    xmin = x0 + kmin < xmin ? x0 + kmin : xmin;
                    xmax = x0 + kmax > xmax ? x0 + kmax : xmax;
                    int* const trash = misc + 30;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        atomicAdd((eq & (0x80u << (8 * k))) ? &hist[(v4 >> (8 * k)) & 0xffu] : trash, 1);
                }
        } else
        for (int y = warp; y < H; y += RADB_NTB / 32)
            for (int x = lane; x < W; x += 32) {
                int i = y * W + x;
                if (radb_roi(s_msk, i, p.label, p.mask_bits)) {
                    np++;
                    ymin = y < ymin ? y : ymin;
                    ymax = y > ymax ? y : ymax;
                    xmin = x < xmin ? x : xmin;
                    xmax = x > xmax ? x : xmax;
                    if (U8) {
                        atomicAdd(&hist[(int)s_img[i]], 1);
                    } else {
                        const double v = (double)s_img[i];
                        vmn = fmin(vmn, v);
                        vmx = fmax(vmx, v);
                    }
                }
            }
        np = warp_sum_i(np);
        ymin = warp_min_i(ymin);
        ymax = warp_max_i(ymax);
        xmin = warp_min_i(xmin);
        xmax = warp_max_i(xmax);
        if (!U8) {
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                vmn = fmin(vmn, __shfl_xor_sync(FULLMASK, vmn, m));
                vmx = fmax(vmx, __shfl_xor_sync(FULLMASK, vmx, m));
            }
            if (lane == 0) { wrange[warp] = vmn; wrange[RADB_NTB / 32 + warp] = vmx; }
        }
        if (lane == 0 && np) {
            atomicAdd(&misc[0], np);
            // zero-neutral encodings so the zeroed scratch needs no separate init
            atomicMax(&misc[1], H - ymin);
            atomicMax(&misc[2], ymax + 1);
            atomicMax(&misc[3], W - xmin);
            atomicMax(&misc[4], xmax + 1);
        }
    }
    __syncthreads();
    // ROI validity (A.1 step 2, imageoperations.checkMask) and the bin edges (A.3 getBinEdges)
    double low = 0, roi_min = 0, roi_max = 0;
    double bwv = p.bin_width;   // bin width in effect (binCount: derived from the ROI range)
    int lcap = 0x7fffffff;      // binCount: the number of bins caps the level
    {
        const int np = misc[0];
        int st = 0;
        if (np == 0)
            st = 1;
        else {
            int nd = ((misc[2] - 1) > (H - misc[1])) + ((misc[4] - 1) > (W - misc[3]));
            if (nd == 0) st = 2;
            else if (nd < 2) st = 3;
        }
        double vmin = 1e308, vmax = -1e308;
        if (U8) {
            int imin = 256, imax = -1;
            for (int k = 0; k < 8; k++) {
                const int v = k * 32 + lane;  // lane <-> bank
                if (hist[v]) { imin = v < imin ? v : imin; imax = v > imax ? v : imax; }
            }
            vmin = (double)warp_min_i(imin);
            vmax = (double)warp_max_i(imax);
        } else {
            for (int w = 0; w < RADB_NTB / 32; w++) {
                vmin = fmin(vmin, wrange[w]);
                vmax = fmax(vmax, wrange[RADB_NTB / 32 + w]);
            }
        }
        roi_min = vmin;
        roi_max = vmax;
        int ng = 0;
        if (!st) {
            const double bw = p.bin_width;
            if (FAST || (U8 && p.bw_int && p.bin_count <= 0)) {
                // uint8 pixels and an integer binWidth (25, 10: the reference's settings): the fp64 edges low + k*bw
                // are exact integers, so the level is an integer quotient (float reciprocal, exact below 2^16)
                const int ibw = p.bw_int, ivmin = (int)vmin, ivmax = (int)vmax, ilow = ivmin - ivmin % ibw;
                const float rbw = 1.0f / (float)ibw;
                low = (double)ilow;
                RADB_UNROLL(1)  // (code size: the kernel is instruction-fetch sensitive, profiles/)
                for (int v = tid; v < 256; v += RADB_NTB)
                    lut[v] = (v >= ivmin && v <= ivmax) ? (unsigned char)((int)(((float)(v - ilow) + 0.5f) * rbw) + 1) : 0;
                ng = (int)(((float)(ivmax - ilow) + 0.5f) * rbw) + 1;
            } else {
                if (p.bin_count > 0) {
                    // binCount (imageoperations.getBinEdges): numpy.histogram's edges = linspace(min, max, n + 1)
                    // (a flat ROI gets the range min -+ 0.5), the last edge moved to max + 1, so level = #edges <= x
                    // among the first n: the same counting with low = min, width = (max - min) / n, capped at n
                    const bool flat = !(vmax > vmin);
                    low = flat ? vmin - 0.5 : vmin;
                    bwv = flat ? 1.0 / (double)p.bin_count : (vmax - vmin) / (double)p.bin_count;
                    lcap = p.bin_count;
                } else {
                    low = vmin - py_mod(vmin, bw);
                }
                if (U8) {  // value -> level LUT: level = #edges <= x, edges = low + k*binWidth
                    for (int v = tid; v < 256; v += RADB_NTB) {
                        int L = 0;
                        if ((double)v >= vmin && (double)v <= vmax) {
                            L = radb_level((double)v, low, bwv);
                            L = L > lcap ? lcap : L;
                            if (L > 255) L = 255;  // reported through status 4 below
                        }
                        lut[v] = (unsigned char)L;
                    }
                }
                ng = radb_level(vmax, low, bwv);
                ng = ng > lcap ? lcap : ng;
            }
            if (ng > p.max_ng) st = 4;
        }
        if (st) {
            RADB_UNROLL(1)  // (code size: the kernel is instruction-fetch sensitive, profiles/)
            for (int f = tid; f < p.F; f += RADB_NTB) out[f] = nan_f64();
            if (tid == 0) {
                p.status[row] = st;
                if (DBG && p.dbg_ng) p.dbg_ng[patch] = 0;
            }
            return;
        }
        if (tid == 0) { misc[8] = ng; p.status[row] = 0; }
    }
    __syncthreads();
    const int ng = misc[8];

    // Only first-order features enabled (the discretise / histogram stage on its own: the HBM-bound part of the
    // path): uint8 pixels need neither the level image nor any texture matrix -- level histogram from the raw
    // histogram through the LUT, publish the head of the record, done.
    const bool dbg_on = !FAST && DBG && (p.dbg_levels || p.dbg_glcm || p.dbg_glrlm || p.dbg_glszm || p.dbg_gldm || p.dbg_ngn);
    const bool tex = FAST || dbg_on || p.off_glcm >= 0 || p.off_gldm >= 0 || p.off_glrlm >= 0 || p.off_glszm >= 0 || p.off_ngtdm >= 0;
    if (!tex && U8) {
        for (int v = tid; v < 256; v += RADB_NTB)
            if (hist[v]) atomicAdd(&lhist[lut[v] - 1], hist[v]);
        __syncthreads();
        const uint4* src = (const uint4*)(smem + p.o_rec);
        uint4* dst = (uint4*)g_rec;
        const int n16 = ((p.big ? p.o_gldm : p.o_glcm) - p.o_rec) / 16;  // header + raw histogram + level histogram
        for (int i = tid; i < n16; i += RADB_NTB) dst[i] = src[i];
        return;
    }

    // ---- phase 2: discretised level image (padded, 0 outside the ROI) + level histogram
    if (U8 && vec4) {
        const int WQ = W >> 2, NQ = HW >> 2;
        const float inv_wq = 1.0f / (float)WQ;
        const unsigned l4 = (unsigned)(p.label & 0xff) * 0x01010101u;
        const unsigned* v4p = (const unsigned*)s_img;
        if (p.mask_bits || (p.label >= 0 && p.label <= 255))
            RADB_UNROLL(1)  // (code size: the kernel is instruction-fetch sensitive, profiles/)
            for (int q = tid; q < NQ; q += RADB_NTB) {
                const unsigned eq = radb_roi4(s_msk, q, l4, p.mask_bits);
                if (!eq) continue;  // the level image is pre-zeroed
                const int y = (int)(((float)q + 0.5f) * inv_wq), xq = q - y * WQ;
                const unsigned v4 = v4p[q];
                unsigned w = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) w |= (unsigned)lut[(v4 >> (8 * k)) & 0xffu] << (8 * k);  // four unconditional lookups
                w &= (eq >> 7) * 0xffu;  // keep the ROI bytes (eq: 0x80 per selected byte -> 0xff per selected byte)
                ((unsigned*)((unsigned char*)lev + (y + 1) * WP + XO))[xq] = w;  // u8 levels; XO = 4 and WP % 4 == 0: aligned
            }
    } else
    for (int y = warp; y < H; y += RADB_NTB / 32)
        for (int x = lane; x < W; x += 32) {
            int i = y * W + x;
            LT L = 0;
            if (radb_roi(s_msk, i, p.label, p.mask_bits)) {
                if (U8) {
                    L = lut[(int)s_img[i]];
                } else {
                    const int Lv = radb_level((double)s_img[i], low, bwv);
                    L = (LT)(Lv > lcap ? lcap : Lv);
                    atomicAdd(&lhist[L - 1], 1);
                }
            }
            lev[(y + 1) * WP + x + XO] = L;
        }
    if (U8)
        RADB_UNROLL(1)  // (code size: the kernel is instruction-fetch sensitive, profiles/)
        for (int v = tid; v < 256; v += RADB_NTB)
            if (hist[v]) atomicAdd(&lhist[lut[v] - 1], hist[v]);
    if (dense) {
        // dense GLRLM: the record's cells for run lengths > 16 are only ever touched by the (rare) global atomics of the
        // line walks: zero them for the rows that exist (the barrier below orders this before the walks)
        const int per = p.nrp / 8 - DL / 8, tot = NA * ng * per;  // uint4 (8 packed cells) per (angle, level) row
        const float rper = 1.0f / (float)(per > 0 ? per : 1), rng = 1.0f / (float)ng;
        uint4* Z = (uint4*)(g_rec + (p.o_glrlm - p.o_rec));
        const uint4 zero = {0u, 0u, 0u, 0u};
        RADB_UNROLL(1)
        for (int t = tid; t < tot; t += RADB_NTB) {
            const int row = (int)(((float)t + 0.5f) * rper), k = t - row * per;
            const int a = (int)(((float)row + 0.5f) * rng), g = row - a * ng;
            Z[a * (p.glrlm_stride / 16) + g * (p.nrp / 8) + DL / 8 + k] = zero;
        }
    }
    __syncthreads();
    if (!U8 && p.off_fo >= 0)  // first-order features need the raw values: now, before the stage is re-used
        fo_dispatch<PT>::run(p, s_img, s_msk, HW, misc[0], roi_min, roi_max, lhist, ng, smem + p.o_fo, out + p.off_fo, tid);

    // ROI bounding box (every later pixel pass runs over the bbox only, linearised so that all
    // lanes of a warp have work)
    const int by0 = H - misc[1], bx0 = W - misc[3];
    const int bh = misc[2] - by0, bw = misc[4] - bx0;
    const int nbox = bh * bw;
    const float inv_bw = 1.0f / (float)bw;
    // which angle (if any) is the along-row offset (0, +-1): its line walk doubles as the
    // run-labelling pass of the zone finder
    int a_row = FAST ? 1 : -1;
    if (!FAST)
        for (int a = 0; a < NA; a++)
            if (p.ang_y[a] == 0) a_row = a;
    if (a_row < 0) {  // no along-row connectivity: every ROI pixel starts as its own run of length 1
        for (int i = tid; i < H * LP; i += RADB_NTB) lab[i] = (((UW)1) << US) | (UW)i;
    }

    // ---- phase 3a: line walks over the ROI bounding box.  One thread walks one line along one
    // angle, so every lane runs the same trip count: GLRLM runs for all angles; the along-row
    // walk also writes, for every ROI pixel, label = run start | (position inside the run) << US, so the
    // END pixel of a run carries the run length (seed of the zone sizes).  Diagonal lines are wrapped
    // inside the bbox (a wrap forces a run break), so every angle is bw (or bh) lines of equal length.
    // The steps are branch-free: a run ends where the NEXT level differs (outside the bbox every level is
    // 0), and the counter update is one PREDICATED `red.shared` per step instead of a divergent run-end block
    // (which ran at 13 of 32 lanes and was 30 % of the kernel's instructions); lanes that end no run do not touch
    // shared memory (redirecting them to scratch words cost bank conflicts: the shared-memory pipe is the second
    // bottleneck of this kernel).  Narrow mode tracks the byte address of the u16 counter of (level 0, current
    // position) and the increment (1 or 1 << 16: the GLRLM pitch is even, so the half alternates with the
    // position), which leaves one multiply-add and one mask per step for the address.
    // Task layout: the rows (padded to whole warps), then the lines of all other angles back to back -- they
    // run the same code with per-thread data (x step, counter base), so warps are filled across angles.
    {
        const int nrp = p.nrp;
        const int row_tasks = a_row >= 0 ? bh : 0, row_slots = (row_tasks + 31) & ~31;
        int oth[RADB_MAX_ANGLES] = {0, 0, 0, 0}, noth = 0;
        for (int a = 0; a < NA; a++)
            if (a != a_row) oth[noth++] = a;
        const int ntasks = row_slots + noth * bw;
        const unsigned gpitch = 2u * (unsigned)nrp;  // bytes per level of the u16 counters
        for (int t = tid; t < ntasks; t += RADB_NTB) {
            const bool is_row = t < row_slots;
            if (is_row && t >= row_tasks) continue;
            int a = a_row, l = t;
            if (!is_row) {
                const int u = t - row_slots, k = (int)(((float)u + 0.5f) * inv_bw);
                a = oth[0];
                if (k == 1) a = oth[1];
                if (k == 2) a = oth[2];
                if (k == 3) a = oth[3];
                l = u - k * bw;
            }
            unsigned* R = (unsigned*)(glrlm_base + a * p.glrlm_stride);
            // narrow, packed: kc = byte address of the u16 counter (level 0, current position) -- one level row below R
            const radb_saddr kc0 = radb_to_saddr(R) - gpitch;
            radb_saddr kc = kc0, kmax = kc0 - 2;
            unsigned val = 1u;
            int rl = 0, mylen = 0;  // wide: position inside the current run, longest run
            // narrow, dense: counter of (level g, run length k + 1 <= 16) at Dm[g * 16 + k]; longer runs (rare) go to the
            // packed record in global memory (zeroed in phase 2)
            int* const Dm = (int*)glrlm_base + (a * p.max_ng - 1) * DL;
            unsigned* const Rg = (unsigned*)(g_rec + (p.o_glrlm - p.o_rec) + a * p.glrlm_stride);
            int* const dtrash = misc + 30;
            int ki = 0, kimax = -1;
            bool endp = true;       // "the previous pixel ended a run"
            if (is_row) {
                const int base = (by0 + l + 1) * WP + bx0 + XO, lbase = (by0 + l) * LP + bx0;
                int gn = lev[base];  // software-pipelined: the next level is in flight while this one is processed
                int rp = 0;          // position inside the run (the label needs it in both modes)
                RADB_UNROLL(RADB_WALK_UNROLL)
                for (int x = 0; x < bw; x++) {
                    const int g = gn;
                    gn = lev[base + x + 1];  // one past the bbox is the zero border or a pixel outside the ROI: level 0
                    rp = endp ? 1 : rp + 1;
                    if (g) lab[lbase + x] = ((UW)rp << US) | (UW)(lbase + x - rp + 1);
                    if (WIDE) {
                        endp = (gn != g);
                        if (endp && g) { atomicAdd(&R[(g - 1) * nrp + rp - 1], 1u); mylen = rp > mylen ? rp : mylen; }
                    } else if (dense) {
                        ki = rp - 1;
                        endp = (gn != g);
                        const bool fin = endp && g;
                        atomicAdd((fin && ki < DL) ? &Dm[g * DL + ki] : dtrash, 1);  // ATOMS.POPC.INC (same-address lanes merge)
                        if (fin && ki >= DL) add_u16(Rg, (g - 1) * nrp + ki);
                        if (g) kimax = ki > kimax ? ki : kimax;
                    } else {
                        kc = endp ? kc0 : kc + 2;
                        val = endp ? 1u : val ^ 0x10001u;
                        endp = (gn != g);
                        const radb_saddr w = (kc + (unsigned)g * gpitch) & ~(radb_saddr)3;
                        radb_red_add_if(endp && g, w, val);
                        if (g) kmax = kc > kmax ? kc : kmax;
                    }
                }
            } else {
                const int sdx = p.ang_x[a] * p.ang_y[a];  // x step per +1 in y (runs are direction-agnostic)
                const int xwrap = sdx > 0 ? bw : -1, xre = sdx > 0 ? 0 : bw - 1;  // leaving the bbox on this side / re-entry column
                int x = l;
                int pos = (by0 + 1) * WP + bx0 + XO;
                int gn = lev[pos + x];  // software-pipelined like the row walk
                RADB_UNROLL(RADB_WALK_UNROLL)
                for (int y = 0; y < bh; y++) {
                    const int g = gn;
                    pos += WP;
                    x += sdx;
                    const bool brk = (x == xwrap);  // wrapped diagonal: the next pixel is not a neighbour of this one
                    x = brk ? xre : x;              // (vertical lines: sdx = 0 never reaches xwrap = -1)
                    gn = lev[pos + x];  // row below the bbox on the last step: the zero border, in bounds
                    if (WIDE) {
                        rl = endp ? 1 : rl + 1;
                        endp = (gn != g) || brk;
                        if (endp && g) { atomicAdd(&R[(g - 1) * nrp + rl - 1], 1u); mylen = rl > mylen ? rl : mylen; }
                    } else if (dense) {
                        ki = endp ? 0 : ki + 1;
                        endp = (gn != g) || brk;
                        const bool fin = endp && g;
                        atomicAdd((fin && ki < DL) ? &Dm[g * DL + ki] : dtrash, 1);
                        if (fin && ki >= DL) add_u16(Rg, (g - 1) * nrp + ki);
                        if (g) kimax = ki > kimax ? ki : kimax;
                    } else {
                        kc = endp ? kc0 : kc + 2;
                        val = endp ? 1u : val ^ 0x10001u;
                        endp = (gn != g) || brk;
                        const radb_saddr w = (kc + (unsigned)g * gpitch) & ~(radb_saddr)3;
                        radb_red_add_if(endp && g, w, val);
                        if (g) kmax = kc > kmax ? kc : kmax;
                    }
                }
            }
            if (!WIDE) mylen = dense ? kimax + 1 : (int)(kmax + 2 - kc0) >> 1;
            if (mylen) atomicMax(&misc[10 + a], mylen);  // record header: longest run of angle a
        }
    }
    __syncthreads();

    // ---- phase 3b: neighbourhood pass over the bbox -> GLCM, GLDM, NGTDM, zone unions
    // GLCM counters while they are built: row pitch gp, angle stride gas.  Padded (gp = ng + 1): level 0 -- a
    // neighbour or a centre outside the ROI -- has its own row and column, so the increments need no test.
    const bool glcm_pad = FAST || p.glcm_pad, symmetric = FAST || p.symmetric;
    const int gp = glcm_pad ? ng + 1 : ng, gas = gp * gp;
    int* const glcm0 = glcm_pad ? glcm : glcm - (ng + 1);  // cell (level i, level j) of angle a: glcm0[a * gas + i * gp + j]
    {
        int doff[RADB_MAX_ANGLES], loff[RADB_MAX_ANGLES];
        for (int a = 0; a < RADB_MAX_ANGLES; a++) {
            doff[a] = a < NA ? p.ang_y[a] * WP + p.ang_x[a] : 0;
            loff[a] = a < NA ? p.ang_y[a] * LP + p.ang_x[a] : 0;
        }
        const int nd = NB + 1;
        const bool inplane = (NA == 4);  // all 8 neighbours: run-adjacency union rules apply
        // Union requests are queued per warp and executed 32 at a time, so that the
        // data-dependent find loops run with (nearly) all lanes busy.
        const int QCAP = p.uq_cap;
        UW* uq = (UW*)(smem + p.o_uq) + warp * QCAP;
        int qn = 0;
        const unsigned lt_mask = (1u << lane) - 1u;
        auto drain = [&](int count) {  // (inlined at its four call sites: a non-inlined drain function measured 1.5 % slower)
            if (lane < count) {
                const UW pr = uq[qn - count + lane];
                uf_union<UW, UF<WIDE>::S>(lab, (unsigned)(pr >> US), (unsigned)(pr & ULO));
            }
            __syncwarp();
            qn -= count;
        };
        auto push = [&](bool has, UW pair) {
            const unsigned m = __ballot_sync(FULLMASK, has);
            if (has) uq[qn + __popc(m & lt_mask)] = pair;
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) drain(32);
        };
        if (FAST || (!WIDE && !L16 && p.lev4 && glcm_pad && inplane && p.alpha == 0 && a_row == 1)) {
            // Four pixels per thread: the level image is read as aligned 32-bit words (3 rows x 3 words), the eight
            // neighbours of the four pixels of the centre word are byte permutes of those, and the neighbour
            // counts (NGTDM), equal-level counts (GLDM, alpha = 0) and the run-adjacency / run-end tests are
            // byte-parallel over the four pixels.  The four in-plane angles in their canonical order
            // (radb_host.h: make_plan): 0 (1,1) f=SE | 1 (0,1) f=E | 2 (-1,1) f=NE | 3 (1,0) f=S.
            const unsigned* lev32 = (const unsigned*)lev;
            const int WPW = WP >> 2;
            const int qx0 = bx0 >> 2, nqx = ((bx0 + bw - 1) >> 2) - qx0 + 1, nq = bh * nqx;
            const float inv_nqx = 1.0f / (float)nqx;
            int* const gldm_p = gldm - nd;          // row of level c at gldm_p + c * nd (level 0: the pad in front)
            int* const ngc_p = ngc - NB - 1;        // cell (level c, count n >= 1) at ngc_p + c * NB + n
            int* const ngn_p = ngn - NB - 1;
            int* const trash = misc + 30;
            RADB_UNROLL(1)  // (code size: the kernel is instruction-fetch sensitive, profiles/)
            for (int base = 0; base < nq; base += RADB_NTB) {  // uniform trip count (warp collectives below)
                const int idx = base + tid;
                const int yb = (int)(((float)idx + 0.5f) * inv_nqx);
                const int y = by0 + yb, qx = qx0 + (idx - yb * nqx);
                const unsigned* wp = lev32 + (y + 1) * WPW + 1 + qx;
                const unsigned C = idx < nq ? wp[0] : 0u;
                const int li0 = y * LP + 4 * qx;   // union-find index of the word's first pixel
                unsigned F = 0;  // per pixel k (byte k): bit 0 union with N, bit 1 with NW, bit 2 with NE, bit 3 row-run end
                if (C) {
                    const unsigned U0 = wp[-WPW - 1], U1 = wp[-WPW], U2 = wp[-WPW + 1];
                    const unsigned M0 = wp[-1], M2 = wp[1];
                    const unsigned D0 = wp[WPW - 1], D1 = wp[WPW], D2 = wp[WPW + 1];
                    const unsigned NWw = __byte_perm(U0, U1, 0x6543), NEw = __byte_perm(U1, U2, 0x4321);
                    const unsigned Ww = __byte_perm(M0, C, 0x6543), Ew = __byte_perm(C, M2, 0x4321);
                    const unsigned SWw = __byte_perm(D0, D1, 0x6543), SEw = __byte_perm(D1, D2, 0x4321);
                    // 0x80 flags per byte: neighbour level == centre level
                    const unsigned eNW = bytes_eq4(NWw, C), eN = bytes_eq4(U1, C), eNE = bytes_eq4(NEw, C);
                    const unsigned eW = bytes_eq4(Ww, C), eE = bytes_eq4(Ew, C);
                    const unsigned eSW = bytes_eq4(SWw, C), eS = bytes_eq4(D1, C), eSE = bytes_eq4(SEw, C);
                    const unsigned zC = bytes_nz4(C);
                    // per-pixel counts in the four bytes (<= 8 each)
                    const unsigned DEP = (eNW >> 7) + (eN >> 7) + (eNE >> 7) + (eW >> 7) + (eE >> 7) + (eSW >> 7) + (eS >> 7) + (eSE >> 7);
                    const unsigned CNT = (bytes_nz4(NWw) >> 7) + (bytes_nz4(U1) >> 7) + (bytes_nz4(NEw) >> 7) + (bytes_nz4(Ww) >> 7) +
                                         (bytes_nz4(Ew) >> 7) + (bytes_nz4(SWw) >> 7) + (bytes_nz4(D1) >> 7) + (bytes_nz4(SEw) >> 7);
                    // 8-connectivity between row runs, one union per pair of touching runs: with the level above equal,
                    // link to it if this pixel starts its run or the run above starts here; else link a run start to NW
                    // and a run end to NE
                    const unsigned rN = eN & ~(eW & eNW), rNW = ~eN & eNW & ~eW, rNE = ~eN & eNE & ~eE, rE = ~eE & 0x80808080u;
                    F = (((rN >> 7) | (rNW >> 6) | (rNE >> 5) | (rE >> 4)) & (zC >> 7) * 0xfu);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int c = (int)((C >> (8 * k)) & 0xffu);
                        // the 8 neighbours' sum: three bytes of the row above, two of this row, three of the row below
                        unsigned sum;
                        if (k == 0) sum = bytes_dot4(NWw, 0x00010101u, bytes_dot4(Ww, 0x00010001u, bytes_dot4(SWw, 0x00010101u, 0u)));
                        else if (k == 1) sum = bytes_dot4(U1, 0x00010101u, bytes_dot4(C, 0x00010001u, bytes_dot4(D1, 0x00010101u, 0u)));
                        else if (k == 2) sum = bytes_dot4(U1, 0x01010100u, bytes_dot4(C, 0x01000100u, bytes_dot4(D1, 0x01010100u, 0u)));
                        else sum = bytes_dot4(NEw, 0x01010100u, bytes_dot4(Ew, 0x01000100u, bytes_dot4(SEw, 0x01010100u, 0u)));
                        int* const g = glcm0 + c * gp;  // a pixel outside the ROI (c = 0) counts into the garbage row
                        atomicAdd(&g[(SEw >> (8 * k)) & 0xffu], 1);
                        atomicAdd(&g[gas + ((Ew >> (8 * k)) & 0xffu)], 1);
                        atomicAdd(&g[2 * gas + ((NEw >> (8 * k)) & 0xffu)], 1);
                        atomicAdd(&g[3 * gas + ((D1 >> (8 * k)) & 0xffu)], 1);
                        atomicAdd(&gldm_p[c * nd + (int)((DEP >> (8 * k)) & 0xffu)], 1);
                        const int cnt = (int)((CNT >> (8 * k)) & 0xffu);
                        int num = cnt * c - (int)sum;
                        num = num < 0 ? -num : num;
                        const bool has = cnt != 0;  // (c = 0 lands in the pads in front of the two arrays)
                        atomicAdd(has ? &ngc_p[c * NB + cnt] : trash, 1);
                        atomicAdd(has ? &ngn_p[c * NB + cnt] : trash, num);
                    }
                }
                // queue the union requests and append the run ends: one warp scan over (requests | ends << 16)
                const unsigned Fr = F & 0x07070707u, Fe = F & 0x08080808u;
                const int mine = __popc(Fr) | (__popc(Fe) << 16);
                const int excl = warp_excl_scan_i(mine, lane);
                const int tot = __shfl_sync(FULLMASK, excl + mine, 31);
                const int nreq = tot & 0xffff, nend = tot >> 16;
                if (keep_runs && nend) {  // (without the list the zone phases find the run ends by scanning the bbox)
                    int rbase = 0;
                    if (lane == 0) rbase = atomicAdd(&misc[6], nend);
                    rbase = __shfl_sync(FULLMASK, rbase, 0) + (excl >> 16);
                    for (unsigned f = Fe; f; f &= f - 1u) runs[rbase++] = (unsigned short)(li0 + ((__ffs((int)f) - 1) >> 3));
                }
                if (nreq) {
                    if (qn + nreq <= QCAP) {
                        int pos = qn + (excl & 0xffff);
                        for (unsigned f = Fr; f; f &= f - 1u) {
                            const int b = __ffs((int)f) - 1, t = b & 7;
                            const unsigned li = (unsigned)(li0 + (b >> 3));
                            uq[pos++] = ((UW)li << US) | (UW)(li - (unsigned)LP + (unsigned)(t == 0 ? 0 : (t == 1 ? -1 : 1)));
                        }
                        qn += nreq;
                        __syncwarp();
                        while (qn >= 32) drain(32);
                    } else {  // (rare) more requests than the queue holds: pixel by pixel, at most 2 x 32 at a time
#pragma unroll 1
                        for (int k = 0; k < 4; k++) {
                            const unsigned fk = (F >> (8 * k)) & 7u;
                            const unsigned li = (unsigned)(li0 + k);
                            push((fk & 3u) != 0u, ((UW)li << US) | (UW)(li - (unsigned)LP - ((fk & 2u) ? 1u : 0u)));
                            push((fk & 4u) != 0u, ((UW)li << US) | (UW)(li - (unsigned)LP + 1u));
                        }
                    }
                }
            }
            if (qn) { __syncwarp(); drain(qn); }
        } else {
        for (int base = 0; base < nbox; base += RADB_NTB) {  // uniform trip count (warp collectives below)
            const int idx = base + tid;
            const int yb = WIDE ? idx / bw : (int)(((float)idx + 0.5f) * inv_bw);
            const int y = by0 + yb, x = bx0 + (idx - yb * bw);
            const int ctr = (y + 1) * WP + x + XO;
            const int c = idx < nbox ? (int)lev[ctr] : 0;
            const int li = y * LP + x;
            unsigned req[RADB_MAX_ANGLES];  // union partner + 1 (0 = none)
#pragma unroll
            for (int a = 0; a < RADB_MAX_ANGLES; a++) req[a] = 0;
            if (c && inplane) {
                // the four in-plane angles in their canonical order (radb_host.h: make_plan):
                // 0 (1,1) f=SE b=NW | 1 (0,1) f=E b=W | 2 (-1,1) f=NE b=SW | 3 (1,0) f=S b=N
                const int nw = lev[ctr - WP - 1], n_ = lev[ctr - WP], ne = lev[ctr - WP + 1];
                const int w_ = lev[ctr - 1], e_ = lev[ctr + 1];
                const int sw = lev[ctr + WP - 1], s_ = lev[ctr + WP], se = lev[ctr + WP + 1];
                int* g0 = glcm0 + c * gp;
                // Unconditional increments: a neighbour outside the ROI (level 0) is redirected to a scratch word
                // of the record header instead of branching around the atomic -- `if (x) atomicAdd(..)` costs
                // BSSY / BRA / BSYNC per counter in the innermost loop, and the plain form keeps the compiler's
                // ATOMS.POPC.INC (same-address lanes are combined by the hardware).
                int* const trash = ((WIDE && p.big) ? (int*)(g_rec + (p.o_misc - p.o_rec)) : misc) + 30;
                atomicAdd(se ? &g0[se] : trash, 1);
                atomicAdd(e_ ? &g0[gas + e_] : trash, 1);
                atomicAdd(ne ? &g0[2 * gas + ne] : trash, 1);
                atomicAdd(s_ ? &g0[3 * gas + s_] : trash, 1);
                const int al = p.alpha;
                int cnt, sum, dep;
                if (!L16 && al == 0) {
                    // eight u8 neighbours in two words: counts by byte-parallel tests + popc, the sum by IDP.4A
                    // (GLDM with alpha = 0, the pyradiomics default: dependent <=> equal level)
                    const unsigned q0 = (unsigned)nw | ((unsigned)n_ << 8) | ((unsigned)ne << 16) | ((unsigned)w_ << 24);
                    const unsigned q1 = (unsigned)e_ | ((unsigned)sw << 8) | ((unsigned)s_ << 16) | ((unsigned)se << 24);
                    const unsigned c4 = (unsigned)c * 0x01010101u;
                    cnt = __popc(bytes_nz4(q0)) + __popc(bytes_nz4(q1));
                    sum = (int)bytes_sum4(q0, bytes_sum4(q1, 0u));
                    dep = __popc(bytes_eq4(q0, c4)) + __popc(bytes_eq4(q1, c4));
                } else {
                    cnt = (nw != 0) + (n_ != 0) + (ne != 0) + (w_ != 0) + (e_ != 0) + (sw != 0) + (s_ != 0) + (se != 0);
                    sum = nw + n_ + ne + w_ + e_ + sw + s_ + se;  // levels outside the ROI are 0
#define RADB_DEP(v) ((v) != 0 && ((v) - c <= al) && (c - (v) <= al))
                    dep = RADB_DEP(nw) + RADB_DEP(n_) + RADB_DEP(ne) + RADB_DEP(w_) + RADB_DEP(e_) + RADB_DEP(sw) +
                          RADB_DEP(s_) + RADB_DEP(se);
#undef RADB_DEP
                }
                atomicAdd(&gldm[(c - 1) * nd + dep], 1);
                if (cnt) {
                    int num = cnt * c - sum;
                    num = num < 0 ? -num : num;
                    atomicAdd(&ngc[(c - 1) * NB + cnt - 1], 1);
                    atomicAdd(&ngn[(c - 1) * NB + cnt - 1], num);
                }
                // 8-connectivity between row runs: one union per pair of touching runs.
                const bool is_start = w_ != c, is_end = e_ != c;
                if (n_ == c) {
                    if (is_start || nw != c) req[0] = (unsigned)(li - LP) + 1u;
                } else {
                    if (nw == c && is_start) req[0] = (unsigned)(li - LP - 1) + 1u;
                    if (ne == c && is_end) req[1] = (unsigned)(li - LP + 1) + 1u;
                }
            } else if (c) {
                int dep = 0, cnt = 0, sum = 0;
#pragma unroll
                for (int a = 0; a < RADB_MAX_ANGLES; a++) {
                    if (a < NA) {
                        const int f = lev[ctr + doff[a]];
                        const int b = lev[ctr - doff[a]];
                        if (f) {
                            atomicAdd(&glcm0[a * gas + c * gp + f], 1);
                            cnt++;
                            sum += f;
                            int df = f - c;
                            df = df < 0 ? -df : df;
                            dep += (df <= p.alpha);
                        }
                        if (b) {
                            cnt++;
                            sum += b;
                            int db = b - c;
                            db = db < 0 ? -db : db;
                            dep += (db <= p.alpha);
                        }
                        if (a != a_row && b == c) req[a] = (unsigned)(li - loff[a]) + 1u;
                    }
                }
                atomicAdd(&gldm[(c - 1) * nd + dep], 1);
                if (cnt) {
                    int num = cnt * c - sum;
                    num = num < 0 ? -num : num;
                    atomicAdd(&ngc[(c - 1) * NB + cnt - 1], 1);
                    atomicAdd(&ngn[(c - 1) * NB + cnt - 1], num);
                }
            }
            const int nreq = inplane ? 2 : NA;
#pragma unroll
            for (int a = 0; a < RADB_MAX_ANGLES; a++)
                if (a < nreq) push(req[a] != 0, ((UW)(unsigned)li << US) | (UW)(req[a] - 1u));
            if (keep_runs && a_row >= 0) {  // append the row-run ends of this warp's pixels: one atomic per warp
                const bool st = c && lev[ctr + 1] != c;
                const unsigned ms = __ballot_sync(FULLMASK, st);
                if (ms) {
                    int rbase = 0;
                    if (lane == 0) rbase = atomicAdd(&misc[6], __popc(ms));
                    rbase = __shfl_sync(FULLMASK, rbase, 0);
                    if (st) runs[rbase + __popc(ms & lt_mask)] = (unsigned short)li;
                }
            }
        }
        if (qn) drain(qn);
        }
    }
    __syncthreads();

    // ---- phase 4: fold run lengths into their zone root; symmetrise the GLCM
    // A run is visited through its END pixel e, whose size field holds the run length (the line walk wrote
    // position-in-run into every pixel; without an along-row angle every pixel is a run of one).  The start
    // pixel of a run -- the only pixel of a run that can be a root -- holds 1, so a root collects
    // len - 1 from its own run and len from every other run of its zone: size field of a root = zone size.
    const bool by_list = keep_runs && a_row >= 0;  // the run list exists: visit runs, not pixels
    const int nruns = by_list ? misc[6] : 0;
    const float inv_lp = 1.0f / (float)LP;
    // Unions link the larger pixel index under the smaller one, so a tall zone is a long parent chain (the clocks
    // showed chain walks of ~40 hops: 15 % of a CTA's lifetime).  Every run start is owned by exactly one thread
    // here and every chain node is a run start, so the chains are shortened by POINTER JUMPING: a thread keeps
    // replacing its start pixel's parent by the grandparent until that is a root -- all chains halve together,
    // ~log2(depth) steps instead of depth.  (Stores only touch non-roots, whose size field is static; only roots
    // are ever added to.)
    // returns the start pixel of the run if the run is its zone's root run, else 0xffff (no pixel index: H * LP <= 65535)
    auto fold_run = [&](unsigned e) -> unsigned {
        volatile UW* L = (volatile UW*)lab;
        const UW we = L[e];
        if ((unsigned)(we & ULO) == e) return e;  // a one-pixel run that is its zone's root: its size field already counts it
        const unsigned len = (unsigned)(we >> US), st = e - len + 1u;  // e is not a root: its size field (the run length) is static
        UW ws = st == e ? we : L[st];
        unsigned r;
        while (true) {
            const unsigned pa = (unsigned)(ws & ULO);
            if (pa == st) { r = st; break; }
            const unsigned g = (unsigned)(L[pa] & ULO);
            if (g == pa) { r = pa; break; }
            ws = (ws & ~ULO) | (UW)g;
            L[st] = ws;
        }
        const unsigned own = (r == st) ? 1u : 0u;
        atomicAdd(&lab[r], (UW)(len - own) << US);
        return own ? r : 0xffffu;
    };
    // the list entry of a root run becomes its start pixel, every other entry 0xffff: phase 5 then only touches roots
    RADB_UNROLL(1)  // (code size: the kernel is instruction-fetch sensitive, profiles/)
    for (int k = tid; k < nruns; k += RADB_NTB) runs[k] = (unsigned short)fold_run(runs[k]);
    if (!by_list)
    for (int idx = tid; idx < nbox; idx += RADB_NTB) {
        const int yb = WIDE ? idx / bw : (int)(((float)idx + 0.5f) * inv_bw);
        const int y = by0 + yb, x = bx0 + (idx - yb * bw);
        const int ctr = (y + 1) * WP + x + XO;
        const int c = lev[ctr];
        if (c && (a_row < 0 || lev[ctr + 1] != c)) fold_run((unsigned)(y * LP + x));  // run end
    }
    if (glcm_pad) {
        // the padded counters become the record's compact [NA][ng][ng] matrix (symmetrised: P + P^T) in global
        // memory; (a, i, j) by float reciprocals (exact: indices < 2^18)
        const int ng2 = ng * ng, tot = NA * ng2;
        const float r2 = 1.0f / (float)ng2, r1 = 1.0f / (float)ng;
        int* const G = (int*)(g_rec + (p.o_glcm - p.o_rec));
        RADB_UNROLL(1)  // (code size: the kernel is instruction-fetch sensitive, profiles/)
        for (int t = tid; t < tot; t += RADB_NTB) {
            const int a = (int)(((float)t + 0.5f) * r2), cell = t - a * ng2;
            const int i = (int)(((float)cell + 0.5f) * r1), j = cell - i * ng;
            const int* P = glcm0 + a * gas;
            int v = P[(i + 1) * gp + j + 1];
            if (symmetric) v += P[(j + 1) * gp + i + 1];
            G[t] = v;
        }
    } else if (symmetric) {  // big mode: in place in the global record, all angles in one flat loop
        const int ng2 = ng * ng, tot = NA * ng2;
        for (int t = tid; t < tot; t += RADB_NTB) {
            const int a = t / ng2, cell = t - a * ng2;
            const int i = cell / ng, j = cell - i * ng;
            if (i > j) continue;
            int* P = glcm + a * ng2;
            if (i == j) {
                P[cell] *= 2;
            } else {
                const int sum = P[cell] + P[j * ng + i];
                P[cell] = sum;
                P[j * ng + i] = sum;
            }
        }
    }
    __syncthreads();

    // ---- phase 5: zone roots -> GLSZM (dense + overflow)
    if (warp == RADB_NTB / 32 - 1) {  // record header: number of gray levels present in the ROI (MCC: < 2 -> 1)
        int nroi = 0;
        for (int i = lane; i < ng; i += 32) nroi += lhist[i] > 0;
        nroi = warp_sum_i(nroi);
        if (lane == 0) misc[9] = nroi;
    }
    auto emit_zone = [&](int c, int s) {
        if (s <= p.s0) {
            atomicAdd(&szm[(c - 1) * p.s0 + s - 1], 1);
        } else {
            int k = atomicAdd(&misc[5], 1);
            if (k < p.ovf_cap) ovf[k] = ((unsigned)(c - 1) << 24) | (unsigned)s;  // level - 1 (<= 255) | size (< 2^24)
        }
        if (DBG && p.dbg_glszm) atomicAdd(&p.dbg_glszm[(patch * p.max_ng + (c - 1)) * (long long)HW + s - 1], 1);
    };
    // a run end e whose run starts at a root emits that zone (a root that is its run's end is a one-pixel run)
    auto emit_run = [&](unsigned e, int c) {
        const UW we = lab[e];
        if ((unsigned)(we & ULO) == e) { emit_zone(c, (int)(we >> US)); return; }
        const unsigned s = e - (unsigned)(we >> US) + 1u;
        if (s == e) return;
        const UW ws = lab[s];
        if ((unsigned)(ws & ULO) == s) emit_zone(c, (int)(ws >> US));
    };
    RADB_UNROLL(1)  // (code size: the kernel is instruction-fetch sensitive, profiles/)
    for (int k = tid; k < nruns; k += RADB_NTB) {
        const int r = runs[k];
        if (r == 0xffff) continue;
        const int y = (int)(((float)r + 0.5f) * inv_lp), x = r - y * LP;  // exact for r < 65536
        emit_zone((int)lev[(y + 1) * WP + x + XO], (int)(lab[r] >> US));
    }
    if (!by_list)
    for (int idx = tid; idx < nbox; idx += RADB_NTB) {
        const int yb = WIDE ? idx / bw : (int)(((float)idx + 0.5f) * inv_bw);
        const int y = by0 + yb, x = bx0 + (idx - yb * bw);
        const int ctr = (y + 1) * WP + x + XO;
        const int c = lev[ctr];
        if (!c || (a_row >= 0 && lev[ctr + 1] == c)) continue;
        emit_run((unsigned)(y * LP + x), c);
    }
    __syncthreads();

    // ---- optional debug dump of the integer matrices (parity tests)
    if (DBG && p.dbg_ng && tid == 0) p.dbg_ng[patch] = ng;
    if (DBG && p.dbg_levels)
        for (int i = tid; i < HW; i += RADB_NTB)
            p.dbg_levels[patch * HW + i] = lev[(i / W + 1) * WP + (i % W) + XO];
    if (DBG && p.dbg_glcm) {
        const int* G = (const int*)(g_rec + (p.o_glcm - p.o_rec));  // compact, written in phase 4 (visible after the barrier)
        for (int a = 0; a < NA; a++)
            for (int t = tid; t < ng * ng; t += RADB_NTB)
                p.dbg_glcm[((patch * NA + a) * p.max_ng + t / ng) * p.max_ng + t % ng] = G[a * ng * ng + t];
    }
    if (DBG && p.dbg_glrlm && !dense)
        for (int a = 0; a < NA; a++)
            for (int t = tid; t < ng * p.nr; t += RADB_NTB)
                p.dbg_glrlm[((patch * NA + a) * p.max_ng + t / p.nr) * p.nr + t % p.nr] =
                    get_run((const unsigned*)(glrlm_base + a * p.glrlm_stride), (t / p.nr) * p.nrp + t % p.nr, WIDE);
    if (DBG && p.dbg_gldm)
        for (int t = tid; t < ng * (NB + 1); t += RADB_NTB)
            p.dbg_gldm[patch * p.max_ng * (NB + 1) + t] = gldm[t];

    // ---- phase 6: publish the record (header + every integer matrix) for the reduction kernels
    {
        const uint4* src = (const uint4*)(smem + p.o_rec);
        uint4* dst = (uint4*)g_rec;
        const int n16 = p.rec_copy_bytes / 16;  // wide: GLRLM and the overflow list are already in place
        // the GLCM region was written (compact) in phase 4; big mode keeps it out of the copied part altogether
        const int g0 = glcm_pad ? (p.o_glcm - p.o_rec) / 16 : n16, g1 = glcm_pad ? (p.o_gldm - p.o_rec) / 16 : n16;
        RADB_UNROLL(1)  // (code size: the kernel is instruction-fetch sensitive, profiles/)
        for (int i = tid; i < n16; i += RADB_NTB)
            if (i < g0 || i >= g1) dst[i] = src[i];
    }
    if (dense) {  // the dense counters (run lengths <= 16) packed into the record: 8 u32 -> one uint4 of u16
        const int tot = NA * ng * (DL / 8);
        const float r2 = 1.0f / (float)(ng * (DL / 8));
        RADB_UNROLL(1)
        for (int t = tid; t < tot; t += RADB_NTB) {
            const int a = (int)(((float)t + 0.5f) * r2), r = t - a * ng * (DL / 8), g = r / (DL / 8), hh = r % (DL / 8);
            const int4* sp = (const int4*)((const int*)glrlm_base + (a * p.max_ng + g) * DL + hh * 8);
            const int4 c0 = sp[0], c1 = sp[1];
            uint4 o;
            o.x = (unsigned)c0.x | ((unsigned)c0.y << 16);
            o.y = (unsigned)c0.z | ((unsigned)c0.w << 16);
            o.z = (unsigned)c1.x | ((unsigned)c1.y << 16);
            o.w = (unsigned)c1.z | ((unsigned)c1.w << 16);
            *(uint4*)(g_rec + (p.o_glrlm - p.o_rec) + a * p.glrlm_stride + g * p.nrp * 2 + hh * 16) = o;
        }
        if (DBG && p.dbg_glrlm) {  // parity tests: the packed record, once it is complete
            __syncthreads();
            for (int a = 0; a < NA; a++)
                for (int t = tid; t < ng * p.nr; t += RADB_NTB)
                    p.dbg_glrlm[((patch * NA + a) * p.max_ng + t / p.nr) * p.nr + t % p.nr] =
                        get_u16((const unsigned*)(g_rec + (p.o_glrlm - p.o_rec) + a * p.glrlm_stride), (t / p.nr) * p.nrp + t % p.nr);
        }
    }
}

// ------------------------------------------------------------------ angle kernel: GLRLM + GLCM (+MCC)
// One CTA per patch, warp a <-> angle a.  The matrices are read from the record in the global
// workspace (L2-resident: the build kernel has just written them); scratch lives in shared memory.
__device__ void radb_angle_cta(const RadbParams& p, long long patch, unsigned char* smem)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NA = p.n_angles;
    const long long row = radb_row(p, patch);
    if (p.status[row] != 0) return;  // NaN row already written by the build kernel
    const unsigned char* rec = p.ws + patch * (long long)p.rec_bytes;
    const int* misc = (const int*)(rec + (p.o_misc - p.o_rec));
    const int ng = misc[8], nroi = misc[9];
    double* fsc = (double*)(smem + p.a_fsc);
    int* valid = (int*)(smem + p.a_valid);
    double* out = p.out + row * (long long)p.F;
    RadbTabs tb;
    tb.inv2 = p.g_inv2;
    tb.ninv = p.ninv;
    tb.tlog = p.g_tlog;
    tb.red = (double*)(smem + warp * p.a_warp_bytes + p.a_red);
    for (int a = warp; a < NA; a += RADB_NT / 32) {
        unsigned char* ws = smem + warp * p.a_warp_bytes;
        // zero this warp's integer scratch (px, py, padd, psub, pr are contiguous)
        for (int i = lane; i < (p.a_idx - p.a_px) / 4; i += 32) ((int*)(ws + p.a_px))[i] = 0;
        __syncwarp();
        const unsigned* R = (const unsigned*)(rec + (p.o_glrlm - p.o_rec) + a * p.glrlm_stride);
        int ok = glrlm_task(tb, R, p.wide, ng, p.nrp, misc[10 + a], (int*)(ws + p.a_pr), fsc + a * RADB_FSC_STRIDE + RADB_GLCM_NF, lane);
        if (lane == 0) valid[4 + a] = ok;
        const int* P = (const int*)(rec + (p.o_glcm - p.o_rec)) + a * ng * ng;
        double* mccws = p.big ? (double*)(p.ws_scr + patch * p.scr_bytes + p.g_mcc) + (long long)a * p.mcc_stride
                              : (double*)(ws + p.a_mcc);
        // use_lanczos: the MCC column is written by radb_mcc_combine_kernel (the Lanczos kernel runs concurrently
        // with this one); a placeholder stands in here
        const double mcc_placeholder = 0.0;
        ok = glcm_task(p, tb, P, ng, (int*)(ws + p.a_px), (int*)(ws + p.a_py), (int*)(ws + p.a_padd),
                       (int*)(ws + p.a_psub), mccws, ws + p.a_idx, fsc + a * RADB_FSC_STRIDE, lane,
                       p.use_lanczos ? &mcc_placeholder : (const double*)0);
        if (lane == 0) valid[a] = ok;
    }
    __syncthreads();
    // nanmean over the non-empty angles
    if (p.off_glcm >= 0 && tid < RADB_GLCM_NF) {
        double s = 0;
        int k = 0;
        for (int a = 0; a < NA; a++)
            if (valid[a]) { s += fsc[a * RADB_FSC_STRIDE + tid]; k++; }
        double v = k ? s / (double)k : nan_f64();
        if (tid == 19 && nroi < 2) v = 1.0;
        out[p.off_glcm + tid] = v;
    }
    if (p.off_glrlm >= 0 && tid >= 32 && tid < 32 + RADB_GLRLM_NF) {
        const int f = tid - 32;
        double s = 0;
        int k = 0;
        for (int a = 0; a < NA; a++)
            if (valid[4 + a]) { s += fsc[a * RADB_FSC_STRIDE + RADB_GLCM_NF + f]; k++; }
        out[p.off_glrlm + f] = k ? s / (double)k : nan_f64();
    }
}

// ------------------------------------------------------------------ misc kernel: GLSZM, GLDM, NGTDM, first-order
// One CTA per patch, one warp per feature class.
__device__ void radb_misc_cta(const RadbParams& p, long long patch, unsigned char* smem)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NB = 2 * p.n_angles;
    const long long row = radb_row(p, patch);
    if (p.status[row] != 0) return;
    const unsigned char* rec = p.ws + patch * (long long)p.rec_bytes;
    const int* misc = (const int*)(rec + (p.o_misc - p.o_rec));
    const int ng = misc[8];
    double* out = p.out + row * (long long)p.F;
    RadbTabs tb;
    tb.inv2 = p.g_inv2;
    tb.ninv = p.ninv;
    tb.tlog = p.g_tlog;
    tb.red = (double*)(smem + p.m_red) + warp * RADB_RED_DOUBLES;
    if (p.only_big_ovf && (misc[5] <= RADB_LANE_MAX_OVF || p.off_glszm < 0)) return;  // done by radb_misc_lane_kernel
    for (int t = warp; t < (p.only_big_ovf ? 1 : 4); t += RADB_NT / 32) {
        if (t == 0) {
            if (p.off_glszm < 0) continue;
            int* pg = (int*)(smem + p.m_pg);
            for (int i = lane; i < ng; i += 32) pg[i] = 0;
            __syncwarp();
            const int novf = misc[5] < p.ovf_cap ? misc[5] : p.ovf_cap;
            glszm_task(p, tb, (const int*)(rec + (p.o_szm - p.o_rec)), (const unsigned*)(rec + (p.o_ovf - p.o_rec)),
                       (unsigned*)(smem + p.m_ovf2), novf, ng, pg, out + p.off_glszm, lane);
        } else if (t == 1) {
            if (p.off_gldm < 0) continue;
            gldm_task(tb, (const int*)(rec + (p.o_gldm - p.o_rec)), ng, NB + 1, out + p.off_gldm, lane);
        } else if (t == 2) {
            if (p.off_ngtdm < 0 && !p.dbg_ngn) continue;
            double* pi = (double*)(smem + p.m_ngp);
            double dummy[5];
            ngtdm_task((const int*)(rec + (p.o_ngc - p.o_rec)), (const int*)(rec + (p.o_ngn - p.o_rec)), ng, NB, pi,
                       pi + ng, p.off_ngtdm >= 0 ? out + p.off_ngtdm : dummy, lane,
                       p.dbg_ngn ? p.dbg_ngn + patch * p.max_ng : (int*)0,
                       p.dbg_ngs ? p.dbg_ngs + patch * p.max_ng : (double*)0);
        } else {
            if (p.off_fo < 0 || p.pix_bytes != 1) continue;  // non-uint8: done by the build kernel
            fo_task_u8(p, tb, (const int*)(rec + (p.o_hist - p.o_rec)), (const int*)(rec + (p.o_lhist - p.o_rec)), ng,
                       (double*)(smem + p.m_qv), out + p.off_fo, lane);
        }
    }
}

// Nearest-neighbour mask resize (RadiomicExtractor.py:34-35: cv2.resize(mask, (W, H), INTER_NEAREST) when the mask's size
// differs from the image's).  cv2 4.x (resizeNN): source index = min(floor(dst_index * (1 / (dst_size / src_size))),
// src_size - 1) per axis, in double precision; `ify` / `ifx` are those reciprocal scales, computed by the caller.
// One thread per destination pixel, q = linear index over [n][dH][dW].
__device__ __forceinline__ void radb_resize_mask_thread(const unsigned char* src, int sH, int sW, unsigned char* dst,
                                                        int dH, int dW, long long n, double ify, double ifx, long long q)
{
    const long long per = (long long)dH * dW;
    if (q >= n * per) return;
    const long long b = q / per;
    const int r = (int)(q - b * per), y = r / dW, x = r - y * dW;
    int sy = (int)floor((double)y * ify), sx = (int)floor((double)x * ifx);
    sy = sy < sH - 1 ? sy : sH - 1;
    sx = sx < sW - 1 ? sx : sW - 1;
    dst[q] = src[(b * sH + sy) * (long long)sW + sx];
}
// ------------------------------------------------------------------ channel front-end
// RadiomicExtractor.py:29-30,41-47: cv2.imread gives interleaved BGR uint8; the reference then runs
// execute() on cvtColor(BGR2GRAY), R = im[:,:,2], G = im[:,:,1], B = im[:,:,0].  This kernel reads the
// interleaved pixels once and writes the four planes [image][gray, R, G, B][H][W].  Gray is OpenCV's
// 8-bit fixed-point BT.601: (R*9798 + G*19235 + B*3735 + 2^14) >> 15 (bit-exact against
// opencv 4.13, tests/test_emu_kernel.py).  One thread converts 4 pixels (12 bytes in, 4x4 bytes out).
__device__ void radb_bgr_planes_thread(const unsigned char* bgr, unsigned char* planes, long long n_images,
                                       long long HW, long long q)
{
    const long long quads = (HW + 3) / 4;
    if (q >= n_images * quads) return;
    const long long img = q / quads, p0 = (q - img * quads) * 4;
    const unsigned char* src = bgr + (img * HW + p0) * 3;
    unsigned char* dst = planes + img * 4 * HW + p0;
    const int cnt = (int)(HW - p0 < 4 ? HW - p0 : 4);
    for (int k = 0; k < cnt; k++) {
        const int b = src[3 * k], g = src[3 * k + 1], r = src[3 * k + 2];
        dst[k] = (unsigned char)((r * 9798 + g * 19235 + b * 3735 + 16384) >> 15);
        dst[HW + k] = (unsigned char)r;
        dst[2 * HW + k] = (unsigned char)g;
        dst[3 * HW + k] = (unsigned char)b;
    }
}

// ------------------------------------------------------------------ derived image types (point-wise)
// pyradiomics imageoperations.get{Square,SquareRoot,Logarithm,Exponential}Image (params.yml:141-144):
// y = g(x; M) in float64 with M = max |x| over the whole image (for the logarithm the second
// normalisation max|log(|x|+1)| = log(M+1) because the transform is monotone in |x|).
enum { RADB_IT_SQUARE = 1, RADB_IT_SQUAREROOT = 2, RADB_IT_LOGARITHM = 3, RADB_IT_EXPONENTIAL = 4 };
__device__ __forceinline__ double radb_derive_px(int type, double x, double M)
{
    switch (type) {
        case RADB_IT_SQUARE: {
            const double c = 1.0 / sqrt(M);
            return (c * x) * (c * x);
        }
        case RADB_IT_SQUAREROOT:
            return x > 0 ? sqrt(x * M) : (x < 0 ? -sqrt(-x * M) : x);
        case RADB_IT_LOGARITHM: {
            const double y = x > 0 ? log(x + 1.0) : (x < 0 ? -log(-(x - 1.0)) : x);
            return y * (M / log(M + 1.0));
        }
        case RADB_IT_EXPONENTIAL:
            return exp((log(M) / M) * x);
    }
    return x;
}

// ------------------------------------------------------------------ shape kernel: shape2D (9 features)
// pyradiomics shape2D.py + cshape.c:calculate_coefficients2D (SURVEY.md section 8 f rank 1), mask only:
// marching squares over the zero-padded mask.  Everything reduces to integers: the perimeter is
// (#axis-aligned segments) + sqrt(1/2) * (#diagonal segments), the mesh surface is a sum of
// eighths per 2x2 cell (equal to the |shoelace| sum of the contour), the maximum diameter is the
// largest distance between contour vertices (edge midpoints; only the extreme vertices of every
// row can be hull vertices, so 2 candidates per half-row are kept), the axes come from the exact
// second moments of the ROI pixel coordinates.
__device__ void radb_shape_cta(const RadbParams& p, long long patch, unsigned char* smem)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const int H = p.H, W = p.W;
    const long long row = radb_row(p, patch);
    if (p.status[row] != 0 || p.off_shape < 0) return;
    const unsigned char* m = radb_mask_ptr(p, patch);
    double* out = p.out + row * (long long)p.F + p.off_shape;
    const int nrow = 2 * H + 1;                     // doubled y coordinates 0..2H of contour vertices (shifted by +1)
    int* xmin = (int*)smem;                         // [nrow]
    int* xmax = xmin + nrow;                        // [nrow]
    unsigned long long* cta = (unsigned long long*)(smem + ((2 * nrow * 4 + 15) & ~15));  // [9] integer totals
    int* wmax = (int*)(cta + 10);                   // [RADB_NT / 32][2]
    for (int i = tid; i < nrow; i += RADB_NT) { xmin[i] = 0x7fffffff; xmax[i] = -1; }
    if (tid < 9) cta[tid] = 0;
    __syncthreads();
    const int label = p.label;
    long long n = 0, sy = 0, sx = 0, syy = 0, sxx = 0, sxy = 0, eighths = 0;
    int nstraight = 0, ndiag = 0;
    // one thread per 2x2 cell of the padded mask: cell (iy, ix) has corners (iy..iy+1, ix..ix+1), iy, ix >= -1
    const int cw = W + 1, ncell = (H + 1) * cw;
    for (int t = tid; t < ncell; t += RADB_NT) {
        const int iy = t / cw - 1, ix = t - (iy + 1) * cw - 1;
        const bool in0 = iy >= 0, in1 = iy + 1 < H, jn0 = ix >= 0, jn1 = ix + 1 < W;
        const int c0 = (in0 && jn0) ? (int)radb_roi(m, iy * W + ix, label, p.mask_bits) : 0;
        const int c1 = (in0 && jn1) ? (int)radb_roi(m, iy * W + ix + 1, label, p.mask_bits) : 0;
        const int c2 = (in1 && jn1) ? (int)radb_roi(m, (iy + 1) * W + ix + 1, label, p.mask_bits) : 0;
        const int c3 = (in1 && jn0) ? (int)radb_roi(m, (iy + 1) * W + ix, label, p.mask_bits) : 0;
        const int k = c0 + c1 + c2 + c3;
        if (c2) {  // pixel (iy+1, ix+1) is visited exactly once as corner 2
            const long long y = iy + 1, x = ix + 1;
            n++; sy += y; sx += x; syy += y * y; sxx += x * x; sxy += x * y;
        }
        if (k == 0) continue;
        const bool diagonal = (k == 2) && (c0 == c2);
        eighths += k == 4 ? 8 : k == 3 ? 7 : k == 1 ? 1 : diagonal ? 2 : 4;
        if (k == 4) continue;
        if (k == 2 && !diagonal) nstraight++; else ndiag += diagonal ? 2 : 1;
        // contour vertices owned by this cell: midpoints of its top edge (c0|c1) and left edge (c0|c3);
        // coordinates doubled and shifted by +1 so they are non-negative
        if (c0 != c1) {
            const int y2 = 2 * iy + 1, x2 = 2 * ix + 2;
            atomicMin(&xmin[y2], x2);
            atomicMax(&xmax[y2], x2);
        }
        if (c0 != c3) {
            const int y2 = 2 * iy + 2, x2 = 2 * ix + 1;
            atomicMin(&xmin[y2], x2);
            atomicMax(&xmax[y2], x2);
        }
    }
    // CTA reduction of the integer sums
    {
        long long v[9] = {n, sy, sx, syy, sxx, sxy, eighths, (long long)nstraight, (long long)ndiag};
#pragma unroll 1
        for (int i = 0; i < 9; i++) {
            const long long t = warp_sum_ll(v[i]);
            if (lane == 0 && t) atomicAdd(&cta[i], (unsigned long long)t);
        }
    }
    __syncthreads();
    // maximum squared distance between candidate vertices (doubled coordinates -> exact integers)
    long long best = 0;
    for (int a = tid; a < 2 * nrow; a += RADB_NT) {
        const int ya = a >> 1, xa = (a & 1) ? xmax[ya] : xmin[ya];
        if (xmax[ya] < 0) continue;
        for (int yb = ya; yb < nrow; yb++) {
            if (xmax[yb] < 0) continue;
            const long long dy = yb - ya;
            long long d0 = xmin[yb] - xa, d1 = xmax[yb] - xa;
            d0 = d0 * d0 + dy * dy;
            d1 = d1 * d1 + dy * dy;
            best = d0 > best ? d0 : best;
            best = d1 > best ? d1 : best;
        }
    }
    {
        int hi = (int)(best >> 31), lo = (int)(best & 0x7fffffff);
        // 64-bit max across the CTA through two 31-bit halves (values < 2^62)
        int mhi = warp_max_i(hi);
        lo = (hi == mhi) ? lo : -1;
        int mlo = warp_max_i(lo);
        if (lane == 0) { wmax[(tid >> 5) * 2] = mhi; wmax[(tid >> 5) * 2 + 1] = mlo; }
    }
    __syncthreads();
    if (tid == 0) {
        long long bestall = 0;
        for (int w = 0; w < RADB_NT / 32; w++) {
            long long b = ((long long)wmax[w * 2] << 31) | (long long)wmax[w * 2 + 1];
            bestall = b > bestall ? b : bestall;
        }
        const long long* T = (const long long*)cta;  // n, sy, sx, syy, sxx, sxy, eighths, #straight, #diagonal
        const double N = (double)T[0];
        const double rN2 = 1.0 / (N * N);
        // covariance of the coordinates / sqrt(N): population second moments.  N * sum(y^2) reaches 2^72 for a full
        // 4096 x 4096 mask (int64 overflows beyond ~1700 x 1700): the exact numerators are formed in 128 bits
        typedef __int128 i128;
        const double a = (double)((i128)T[0] * T[3] - (i128)T[1] * T[1]) * rN2;  // var(y)
        const double c = (double)((i128)T[0] * T[4] - (i128)T[2] * T[2]) * rN2;  // var(x)
        const double b = (double)((i128)T[0] * T[5] - (i128)T[1] * T[2]) * rN2;  // cov(x, y)
        const double hd = 0.5 * (a - c), rad = sqrt(hd * hd + b * b), mid = 0.5 * (a + c);
        double e1 = mid + rad, e0 = mid - rad;
        if (e0 < 0 && e0 > -1e-10) e0 = 0;
        if (e1 < 0 && e1 > -1e-10) e1 = 0;
        const double surface = (double)T[6] / 8.0;
        const double perimeter = (double)T[7] + (double)T[8] * 0.70710678118654752440;
        out[0] = (e0 < 0 || e1 < 0) ? nan_f64() : sqrt(e0 / e1);
        out[1] = e1 < 0 ? nan_f64() : sqrt(e1) * 4.0;
        out[2] = sqrt((double)bestall) * 0.5;
        out[3] = surface;
        out[4] = e0 < 0 ? nan_f64() : sqrt(e0) * 4.0;
        out[5] = perimeter;
        out[6] = perimeter / surface;
        out[7] = N;
        out[8] = 2.0 * sqrt(3.14159265358979323846 * surface) / perimeter;
    }
}

#ifndef RADB_EMU
template <typename PT, bool DBG, bool WIDE, bool L16 = false, int FAST = 0>
__global__ void __launch_bounds__(RADB_NTB, RADB_NTB_MINB) radb_build_kernel(const RadbParams p)
{
    extern __shared__ __align__(16) unsigned char radb_smem[];
    radb_build_cta<PT, DBG, WIDE, L16, FAST>(p, (long long)blockIdx.x, radb_smem);
}
#ifndef RADB_ANGLE_MINB
#define RADB_ANGLE_MINB 6
#endif
__global__ void __launch_bounds__(RADB_NT, RADB_ANGLE_MINB) radb_angle_kernel(const RadbParams p)
{
    extern __shared__ __align__(16) unsigned char radb_smem[];
    radb_angle_cta(p, (long long)blockIdx.x, radb_smem);
}
__global__ void __launch_bounds__(RADB_NTL) radb_angle_lane_kernel(const RadbParams p)
{
    extern __shared__ __align__(16) unsigned char radb_smem[];
    radb_angle_lane_cta(p, (long long)blockIdx.x, radb_smem);
}
__global__ void __launch_bounds__(RADB_NTM) radb_mcc_g8_kernel(const RadbParams p)
{
    extern __shared__ __align__(16) unsigned char radb_smem[];
    radb_mcc_g8_cta(p, (long long)blockIdx.x, radb_smem);
}
__global__ void __launch_bounds__(RADB_NTZ) radb_mcc_lanczos_kernel(const RadbParams p)
{
    extern __shared__ __align__(16) unsigned char radb_smem[];
    radb_mcc_lanczos_cta(p, (long long)blockIdx.x, radb_smem);
}
__global__ void radb_mcc_combine_kernel(const RadbParams p)
{
    radb_mcc_combine_thread(p, (long long)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(RADB_NT) radb_misc_lane_kernel(const RadbParams p)
{
    extern __shared__ __align__(16) unsigned char radb_smem[];
    radb_misc_lane_cta(p, (long long)blockIdx.x, radb_smem);
}
__global__ void __launch_bounds__(RADB_NT, 8) radb_misc_kernel(const RadbParams p)
{
    extern __shared__ __align__(16) unsigned char radb_smem[];
    radb_misc_cta(p, (long long)blockIdx.x, radb_smem);
}
// per-image max of uint8 pixels (|x| = x), then the point-wise transform
__global__ void radb_image_max_kernel(const unsigned char* img, long long n_images, long long HW, int* mx)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long words = (HW + 3) / 4;
    int m = 0;
    long long image = 0;
    if (t < n_images * words) {
        image = t / words;
        const long long p0 = (t - image * words) * 4;
        const unsigned char* s = img + image * HW + p0;
        const int cnt = (int)(HW - p0 < 4 ? HW - p0 : 4);
        for (int k = 0; k < cnt; k++) m = s[k] > m ? s[k] : m;
    }
    // lanes of a warp may straddle two images only at image boundaries: plain atomics are enough
    if (t < n_images * words && m) atomicMax(&mx[image], m);
}
__global__ void radb_derive_kernel(const unsigned char* img, long long n_images, long long HW, const int* mx, int type,
                                   double* out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_images * HW) return;
    const long long image = t / HW;
    out[t] = radb_derive_px(type, (double)img[t], (double)mx[image]);
}
// packed mask (bit i of the stream <-> pixel i, radb_pack_mask_host) -> uint8 mask: `label` where the bit is set,
// another value elsewhere.  One thread expands 16 bits into one 128-bit store.
__global__ void radb_unpack_mask_kernel(const unsigned char* packed, long long n_bytes, int label, unsigned char* mask)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long p0 = t * 16;
    if (p0 >= n_bytes) return;
    const unsigned on = (unsigned)(label & 0xff), off = on ? 0u : 255u;
    unsigned bits = packed[p0 >> 3];
    if (p0 + 8 < n_bytes) bits |= (unsigned)packed[(p0 >> 3) + 1] << 8;
    if (p0 + 16 <= n_bytes && (((unsigned long long)(mask + p0)) & 15ull) == 0) {
        unsigned w[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            unsigned v = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) v |= (((bits >> (4 * q + k)) & 1u) ? on : off) << (8 * k);
            w[q] = v;
        }
        *(uint4*)(mask + p0) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        for (int k = 0; k < 16 && p0 + k < n_bytes; k++) mask[p0 + k] = (unsigned char)(((bits >> k) & 1u) ? on : off);
    }
}
__global__ void radb_resize_mask_kernel(const unsigned char* src, int sH, int sW, unsigned char* dst, int dH, int dW,
                                        long long n, double ify, double ifx)
{
    radb_resize_mask_thread(src, sH, sW, dst, dH, dW, n, ify, ifx, (long long)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void radb_bgr_planes_kernel(const unsigned char* bgr, unsigned char* planes, long long n_images, long long HW)
{
    radb_bgr_planes_thread(bgr, planes, n_images, HW, (long long)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(RADB_NT, 8) radb_shape_kernel(const RadbParams p)
{
    extern __shared__ __align__(16) unsigned char radb_smem[];
    radb_shape_cta(p, (long long)blockIdx.x, radb_smem);
}
#endif
