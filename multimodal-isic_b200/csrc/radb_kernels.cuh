// radb -- B200-native radiomic feature kernels (sm_100a).  One CTA (128 threads) per patch:
//   stage   : TMA bulk copy (cp.async.bulk + mbarrier) of the raw patch + mask into shared memory
//   discret.: ROI histogram / bbox / min-max -> value->level LUT -> padded level image (smem)
//   build   : one neighbourhood pass feeds GLCM, GLDM, NGTDM, GLRLM (run walk) and the GLSZM
//             union-find; every matrix lives in shared memory as privatised integer counters
//   reduce  : warp-specialised fp64 feature reductions (warp w <-> angle w), shuffle reductions,
//             MCC through Householder tridiagonalisation + Sturm multisection in shared memory
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

// ------------------------------------------------------------------ union-find on u16 labels
__device__ __forceinline__ unsigned uf_find(volatile unsigned short* L, unsigned x)
{
    unsigned p;
    while ((p = L[x]) != x) x = p;
    return x;
}
__device__ __forceinline__ void uf_union(unsigned short* L, unsigned a, unsigned b)
{
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) { unsigned t = a; a = b; b = t; }
        unsigned short old = atomicCAS(&L[a], (unsigned short)a, (unsigned short)b);
        if (old == (unsigned short)a) return;
        a = old;
    }
}

// packed u16 counters updated with one 32-bit shared atomic
__device__ __forceinline__ void add_u16(unsigned* base, int cell)
{
    atomicAdd(&base[cell >> 1], (cell & 1) ? 0x10000u : 1u);
}
__device__ __forceinline__ int get_u16(const unsigned* base, int cell)
{
    return (int)((base[cell >> 1] >> ((cell & 1) * 16)) & 0xffffu);
}

// Python-style modulo (sign of the divisor), as numpy's % in imageoperations.getBinEdges
__device__ __forceinline__ double py_mod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}

// ------------------------------------------------------------------ first-order (u8 raw histogram)
// One warp.  A.5 / pyradiomics firstorder.py: everything except Entropy/Uniformity comes from the
// raw ROI values; for uint8 pixels those are exactly the 256-bin histogram.
__device__ void fo_task_u8(const RadbParams& p, const int* hist, const int* lhist, int ng, double* qv,
                           double* o, int lane)
{
    int h[8];
    long long s1 = 0, s2 = 0;
    int n = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int v = lane * 8 + k;
        h[k] = hist[v];
        n += h[k];
        s1 += (long long)h[k] * v;
        s2 += (long long)h[k] * v * v;
    }
    const int before = warp_excl_scan_i(n, lane);
    const int N = warp_sum_i(n);
    const double dN = (double)N;
    const double S1 = (double)warp_sum_ll(s1);
    const double mean = S1 / dN;
    const double shift = p.shift;
    // min / max
    int vmin = 256, vmax = -1;
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (h[k]) {
            int v = lane * 8 + k;
            vmin = v < vmin ? v : vmin;
            vmax = v > vmax ? v : vmax;
        }
    vmin = warp_min_i(vmin);
    vmax = warp_max_i(vmax);
    // order statistics for the 10/25/50/75/90 percentiles (numpy 'linear' interpolation)
    const double qs[5] = {10.0, 25.0, 50.0, 75.0, 90.0};
    int rk[10];
    double fr[5];
#pragma unroll
    for (int q = 0; q < 5; q++) {
        double pos = (qs[q] / 100.0) * (dN - 1.0);
        double fl = floor(pos);
        int lo = (int)fl;
        if (lo > N - 1) lo = N - 1;
        int hi = lo + 1 > N - 1 ? N - 1 : lo + 1;
        rk[2 * q] = lo;
        rk[2 * q + 1] = hi;
        fr[q] = pos - fl;
    }
    {
        int cum = before;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (h[k]) {
#pragma unroll
                for (int r = 0; r < 10; r++)
                    if (rk[r] >= cum && rk[r] < cum + h[k]) qv[r] = (double)(lane * 8 + k);
            }
            cum += h[k];
        }
    }
    __syncwarp();
    double pc[5];
#pragma unroll
    for (int q = 0; q < 5; q++) {
        double a = qv[2 * q], b = qv[2 * q + 1];
        pc[q] = a + (b - a) * fr[q];
    }
    const double p10 = pc[0], p25 = pc[1], med = pc[2], p75 = pc[3], p90 = pc[4];
    // central moments, MAD, energy, robust MAD
    double m2 = 0, m3 = 0, m4 = 0, mad = 0, en = 0, in_s1 = 0;
    int in_n = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (!h[k]) continue;
        double v = (double)(lane * 8 + k), hk = (double)h[k];
        double d = v - mean, d2 = d * d;
        m2 += hk * d2;
        m3 += hk * d2 * d;
        m4 += hk * d2 * d2;
        mad += hk * fabs(d);
        en += hk * (v + shift) * (v + shift);
        if (v >= p10 && v <= p90) {
            in_n += h[k];
            in_s1 += hk * v;
        }
    }
    m2 = warp_sum(m2) / dN;
    m3 = warp_sum(m3) / dN;
    m4 = warp_sum(m4) / dN;
    mad = warp_sum(mad) / dN;
    en = warp_sum(en);
    const int inN = warp_sum_i(in_n);
    const double in_mean = warp_sum(in_s1) / (double)inN;
    double rmad = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        double v = (double)(lane * 8 + k);
        if (h[k] && v >= p10 && v <= p90) rmad += (double)h[k] * fabs(v - in_mean);
    }
    rmad = warp_sum(rmad) / (double)inN;
    // discretised histogram: Entropy / Uniformity
    double ent = 0, uni = 0;
    for (int i = lane; i < ng; i += 32) {
        double pi = (double)lhist[i] / dN;
        if (lhist[i]) ent -= pi * log2(pi + RADB_EPS);
        uni += pi * pi;
    }
    ent = warp_sum(ent);
    uni = warp_sum(uni);
    if (lane == 0) {
        o[0] = p10;
        o[1] = p90;
        o[2] = en;
        o[3] = ent;
        o[4] = p75 - p25;
        o[5] = (m2 == 0.0) ? 0.0 : m4 / (m2 * m2);
        o[6] = (double)vmax;
        o[7] = mad;
        o[8] = mean;
        o[9] = med;
        o[10] = (double)vmin;
        o[11] = (double)(vmax - vmin);
        o[12] = rmad;
        o[13] = sqrt(en / dN);
        o[14] = (m2 == 0.0) ? 0.0 : m3 / pow(m2, 1.5);
        o[15] = en;  // TotalEnergy: pixel spacing is (1, 1) for GetImageFromArray images
        o[16] = uni;
        o[17] = m2;
    }
    (void)S1;
}

// ------------------------------------------------------------------ MCC
// packed lower-triangular symmetric matrix: element (r, c), r >= c
__device__ __forceinline__ int tri(int r, int c) { return r * (r + 1) / 2 + c; }
__device__ __forceinline__ double sym_get(const double* M, int r, int c)
{
    return r >= c ? M[tri(r, c)] : M[tri(c, r)];
}

// number of eigenvalues of the symmetric tridiagonal (d, e) that are < x (Sturm count via the
// determinant recurrence, rescaled to stay in range)
__device__ __forceinline__ int sturm_count(const double* d, const double* e2, int m, double x)
{
    int cnt = 0;
    double q = d[0] - x;
    if (q < 0.0) cnt++;
    for (int i = 1; i < m; i++) {
        if (q == 0.0) q = 1e-300;
        q = (d[i] - x) - e2[i - 1] / q;
        if (q < 0.0) cnt++;
    }
    return cnt;
}

// k-th smallest eigenvalue (k = 0..m-1) by warp multisection: 32 shifts per round
__device__ double tridiag_kth(const double* d, const double* e2, int m, int k, double lo, double hi,
                              int lane)
{
    for (int it = 0; it < 14; it++) {
        double w = (hi - lo) / 33.0;
        double x = lo + w * (double)(lane + 1);
        int c = sturm_count(d, e2, m, x);
        // lanes with count <= k are left of (or at) the eigenvalue; they form a prefix
        unsigned left = __ballot_sync(FULLMASK, c <= k);
        int nl = __popc(left);
        double nlo = lo + w * (double)nl;
        double nhi = (nl == 32) ? hi : lo + w * (double)(nl + 1);
        lo = nlo;
        hi = nhi;
        if (hi - lo <= 4e-16 * (fabs(lo) + fabs(hi)) + 1e-300) break;
    }
    return 0.5 * (lo + hi);
}

// One warp per angle.  A.6: MCC = sqrt(second largest eigenvalue of Q),
// Q[i][j] = sum_k P[i][k] P[j][k] / (px[i] py[k]); Q is similar to S = A A^T with
// A = Dx^-1/2 P Dy^-1/2.  For a symmetric GLCM A is symmetric, so the eigenvalues of S are the
// squares of those of A and sqrt(lambda_2(S)) = second largest |lambda(A)|.
__device__ double mcc_task(const int* P, const int* px, const int* py, int n, int symmetric, double* ws,
                           unsigned char* idx, int lane)
{
    // compact the levels that occur in this angle's matrix
    int m = 0;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        int present = (i < n) && (px[i] > 0);
        unsigned b = __ballot_sync(FULLMASK, present);
        if (present) idx[m + __popc(b & ((1u << lane) - 1u))] = (unsigned char)i;
        m += __popc(b);
    }
    __syncwarp();
    if (m < 2) return 0.0;
    double* M = ws;
    double* v = ws + m * (m + 1) / 2;
    double* w = v + m;
    double* d = w + m;
    double* e2 = d + m;
    const int ncell = m * (m + 1) / 2;
    if (symmetric) {
        for (int t = lane; t < ncell; t += 32) {
            int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
            while (tri(r + 1, 0) <= t) r++;
            while (tri(r, 0) > t) r--;
            int c = t - tri(r, 0);
            int ir = idx[r], ic = idx[c];
            M[t] = (double)P[ir * n + ic] / sqrt((double)px[ir] * (double)px[ic]);
        }
    } else {
        for (int t = lane; t < ncell; t += 32) {
            int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
            while (tri(r + 1, 0) <= t) r++;
            while (tri(r, 0) > t) r--;
            int c = t - tri(r, 0);
            int ir = idx[r], ic = idx[c];
            double s = 0;
            for (int k = 0; k < n; k++)
                if (py[k] > 0) s += (double)P[ir * n + k] * (double)P[ic * n + k] / (double)py[k];
            M[t] = s / sqrt((double)px[ir] * (double)px[ic]);
        }
    }
    __syncwarp();
    // Householder tridiagonalisation, column k eliminates rows k+2..m-1
    for (int k = 0; k < m - 2; k++) {
        double part = 0;
        for (int r = k + 2 + lane; r < m; r += 32) { double x = M[tri(r, k)]; part += x * x; }
        double tail = warp_sum(part);
        double x0 = M[tri(k + 1, k)];
        if (tail == 0.0) {
            if (lane == 0) { d[k] = M[tri(k, k)]; e2[k] = x0 * x0; }
            __syncwarp();
            continue;
        }
        double nrm = sqrt(tail + x0 * x0);
        double alpha = x0 > 0 ? -nrm : nrm;
        // v = x - alpha e1 (indices k+1..m-1), beta = 2 / v^T v
        double vtv = tail + (x0 - alpha) * (x0 - alpha);
        double beta = 2.0 / vtv;
        for (int r = k + 1 + lane; r < m; r += 32) v[r] = (r == k + 1) ? x0 - alpha : M[tri(r, k)];
        __syncwarp();
        // w = beta * M22 v
        double kpart = 0;
        for (int r = k + 1 + lane; r < m; r += 32) {
            double s = 0;
            for (int c = k + 1; c < m; c++) s += sym_get(M, r, c) * v[c];
            s *= beta;
            w[r] = s;
            kpart += s * v[r];
        }
        double K = 0.5 * beta * warp_sum(kpart);
        __syncwarp();
        for (int r = k + 1 + lane; r < m; r += 32) w[r] -= K * v[r];
        __syncwarp();
        // M22 -= v w^T + w v^T (lower triangle)
        for (int r = k + 1 + lane; r < m; r += 32) {
            double vr = v[r], wr = w[r];
            for (int c = k + 1; c <= r; c++) M[tri(r, c)] -= vr * w[c] + wr * v[c];
        }
        if (lane == 0) { d[k] = M[tri(k, k)]; e2[k] = alpha * alpha; }
        __syncwarp();
    }
    if (lane == 0) {
        d[m - 2] = M[tri(m - 2, m - 2)];
        double x = M[tri(m - 1, m - 2)];
        e2[m - 2] = x * x;
        d[m - 1] = M[tri(m - 1, m - 1)];
    }
    __syncwarp();
    // Gershgorin bounds
    double glo = 1e300, ghi = -1e300;
    for (int i = lane; i < m; i += 32) {
        double r = (i > 0 ? sqrt(e2[i - 1]) : 0.0) + (i < m - 1 ? sqrt(e2[i]) : 0.0);
        glo = fmin(glo, d[i] - r);
        ghi = fmax(ghi, d[i] + r);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        glo = fmin(glo, __shfl_xor_sync(FULLMASK, glo, s));
        ghi = fmax(ghi, __shfl_xor_sync(FULLMASK, ghi, s));
    }
    double span = ghi - glo;
    glo -= 1e-12 * (span + 1.0);
    ghi += 1e-12 * (span + 1.0);
    double l2 = tridiag_kth(d, e2, m, m - 2, glo, ghi, lane);
    if (symmetric) {
        double l1 = tridiag_kth(d, e2, m, 0, glo, ghi, lane);
        return fmax(fabs(l2), fabs(l1));
    }
    return sqrt(fmax(l2, 0.0));
}

// ------------------------------------------------------------------ GLCM features (one warp, one angle)
// A.6 / pyradiomics glcm.py.  P holds the final integer counts of this angle (symmetrised when
// symmetricalGLCM).  Returns 0 when the angle is empty (upstream deletes it from the nanmean).
__device__ int glcm_task(const RadbParams& p, const int* P, int n, int* px, int* py, int* padd, int* psub,
                         double* ws, unsigned char* idx, double* o, int lane)
{
    long long sN = 0, sI = 0, sJ = 0, sIJ = 0, sD2 = 0, sC2 = 0;
    int maxc = 0;
    for (int i = 0; i < n; i++) {
        int rs = 0;
        for (int j = lane; j < n; j += 32) {
            int c = P[i * n + j];
            if (c) {
                rs += c;
                sIJ += (long long)c * (i + 1) * (j + 1);
                sD2 += (long long)c * (i - j) * (i - j);
                sC2 += (long long)c * c;
                sJ += (long long)c * (j + 1);
                maxc = c > maxc ? c : maxc;
                py[j] += c;  // lane (j mod 32) owns column j
                atomicAdd(&padd[i + j], c);
                atomicAdd(&psub[i > j ? i - j : j - i], c);
            }
        }
        rs = warp_sum_i(rs);
        if (lane == 0) px[i] = rs;
        sN += (lane == 0) ? rs : 0;
        sI += (lane == 0) ? (long long)rs * (i + 1) : 0;
    }
    sN = warp_sum_ll(sN);
    __syncwarp();
    if (sN == 0) return 0;
    const double N = (double)sN;
    const double ux = (double)warp_sum_ll(sI) / N;
    const double uy = (double)warp_sum_ll(sJ) / N;
    const double autoc = (double)warp_sum_ll(sIJ) / N;
    const double contrast = (double)warp_sum_ll(sD2) / N;
    const double energy = (double)warp_sum_ll(sC2) / (N * N);
    const double maxp = (double)warp_max_i(maxc) / N;
    const double log2N = log2(N);
    // pass B: cluster moments, correlation terms, joint entropy
    double ct = 0, cs = 0, cp = 0, ssq = 0, ssqy = 0, corm = 0, hxy = 0, h1corr = 0;
    for (int i = 0; i < n; i++) {
        const double di = (double)(i + 1) - ux;
        const double pxi = (double)px[i];
        for (int j = lane; j < n; j += 32) {
            int c = P[i * n + j];
            if (!c) continue;
            double pij = (double)c / N;
            double dj = (double)(j + 1) - uy;
            double s = (double)(i + 1) + (double)(j + 1) - ux - uy;
            double s2 = s * s;
            ct += pij * s2;
            cs += pij * s2 * s;
            cp += pij * s2 * s2;
            ssq += pij * di * di;
            ssqy += pij * dj * dj;
            corm += pij * di * dj;
            hxy -= pij * log2(pij + RADB_EPS);
            h1corr += (double)c * N / (pxi * (double)py[j]);  // p / (px*py)
        }
    }
    ct = warp_sum(ct);
    cs = warp_sum(cs);
    cp = warp_sum(cp);
    ssq = warp_sum(ssq);
    ssqy = warp_sum(ssqy);
    corm = warp_sum(corm);
    hxy = warp_sum(hxy);
    h1corr = warp_sum(h1corr);
    // marginals
    double hx = 0, hy = 0, hx0 = 0, hy0 = 0;
    int nx = 0, ny = 0;
    for (int i = lane; i < n; i += 32) {
        if (px[i]) {
            double q = (double)px[i] / N;
            hx -= q * log2(q + RADB_EPS);
            hx0 -= q * (log2((double)px[i]) - log2N);
            nx++;
        }
        if (py[i]) {
            double q = (double)py[i] / N;
            hy -= q * log2(q + RADB_EPS);
            hy0 -= q * (log2((double)py[i]) - log2N);
            ny++;
        }
    }
    hx = warp_sum(hx);
    hy = warp_sum(hy);
    hx0 = warp_sum(hx0);
    hy0 = warp_sum(hy0);
    nx = warp_sum_i(nx);
    ny = warp_sum_i(ny);
    // log2(px*py + eps) = log2 px + log2 py + eps/(px*py*ln2) + O(eps^2): HXY1/HXY2 in closed form
    const double hxy1 = hx0 + hy0 - RADB_EPS_LN2 * h1corr;
    const double hxy2 = hx0 + hy0 - RADB_EPS_LN2 * (double)nx * (double)ny;
    // |i-j| marginal
    double da = 0, de = 0, idv = 0, idm = 0, idmn = 0, idn = 0, iv = 0;
    const double dn = (double)n;
    for (int k = lane; k < n; k += 32) {
        if (!psub[k]) continue;
        double q = (double)psub[k] / N, dk = (double)k;
        da += dk * q;
        de -= q * log2(q + RADB_EPS);
        idv += q / (1.0 + dk);
        idm += q / (1.0 + dk * dk);
        idmn += q / (1.0 + (dk * dk) / (dn * dn));
        idn += q / (1.0 + dk / dn);
        if (k > 0) iv += q / (dk * dk);
    }
    da = warp_sum(da);
    de = warp_sum(de);
    idv = warp_sum(idv);
    idm = warp_sum(idm);
    idmn = warp_sum(idmn);
    idn = warp_sum(idn);
    iv = warp_sum(iv);
    double dvar = 0;
    for (int k = lane; k < n; k += 32)
        if (psub[k]) dvar += (double)psub[k] / N * ((double)k - da) * ((double)k - da);
    dvar = warp_sum(dvar);
    // i+j marginal (index k <-> i+j = k+2)
    double sa = 0, se = 0;
    for (int k = lane; k < 2 * n - 1; k += 32) {
        if (!padd[k]) continue;
        double q = (double)padd[k] / N;
        sa += (double)(k + 2) * q;
        se -= q * log2(q + RADB_EPS);
    }
    sa = warp_sum(sa);
    se = warp_sum(se);
    const double mcc = mcc_task(P, px, py, n, p.symmetric, ws, idx, lane);
    if (lane == 0) {
        const double sigx = sqrt(ssq), sigy = sqrt(ssqy);
        const double div = fmax(hx, hy);
        double im2 = 1.0 - exp(-2.0 * (hxy2 - hxy));
        im2 = im2 < 0.0 ? 0.0 : im2;
        o[0] = autoc;
        o[1] = cp;
        o[2] = cs;
        o[3] = ct;
        o[4] = contrast;
        o[5] = (sigx * sigy == 0.0) ? 1.0 : corm / (sigx * sigy + RADB_EPS);
        o[6] = da;
        o[7] = de;
        o[8] = dvar;
        o[9] = idv;
        o[10] = idm;
        o[11] = idmn;
        o[12] = idn;
        o[13] = (div != 0.0) ? (hxy - hxy1) / div : 0.0;
        o[14] = sqrt(im2);
        o[15] = iv;
        o[16] = ux;
        o[17] = energy;
        o[18] = hxy;
        o[19] = mcc;
        o[20] = maxp;
        o[21] = sa;
        o[22] = se;
        o[23] = ssq;
    }
    return 1;
}

// ------------------------------------------------------------------ GLRLM features (one warp, one angle)
// A.7 / pyradiomics glrlm.py.  R = packed u16 counters [n][nr]; pr = int scratch [nr] (zeroed).
__device__ int glrlm_task(const unsigned* R, int cell0, int n, int nr, int* pr, double* o, int lane)
{
    long long sN = 0, sGI = 0, sGI2 = 0, sG2 = 0;
    double lgl = 0, e1 = 0, srl = 0, srh = 0, lrl = 0, lrh = 0;
    int nnz = 0;
    for (int i = 0; i < n; i++) {
        int rs = 0;
        const double i2 = (double)(i + 1) * (double)(i + 1);
        for (int j = lane; j < nr; j += 32) {
            int c = get_u16(R, cell0 + i * nr + j);
            if (!c) continue;
            rs += c;
            pr[j] += c;
            double dc = (double)c, j2 = (double)(j + 1) * (double)(j + 1);
            if (c > 1) e1 += dc * log2(dc);
            nnz++;
            srl += dc / (i2 * j2);
            srh += dc * i2 / j2;
            lrl += dc * j2 / i2;
            lrh += dc * i2 * j2;
        }
        rs = warp_sum_i(rs);
        if (lane == 0 && rs) {
            sN += rs;
            sGI += (long long)rs * (i + 1);
            sGI2 += (long long)rs * (i + 1) * (i + 1);
            sG2 += (long long)rs * rs;
            lgl += (double)rs / i2;
        }
    }
    sN = warp_sum_ll(sN);
    __syncwarp();
    if (sN == 0) return 0;
    sGI = warp_sum_ll(sGI);
    sGI2 = warp_sum_ll(sGI2);
    sG2 = warp_sum_ll(sG2);
    lgl = warp_sum(lgl);
    e1 = warp_sum(e1);
    srl = warp_sum(srl);
    srh = warp_sum(srh);
    lrl = warp_sum(lrl);
    lrh = warp_sum(lrh);
    nnz = warp_sum_i(nnz);
    long long sRJ = 0, sRJ2 = 0, sR2 = 0;
    double sre = 0;
    for (int j = lane; j < nr; j += 32) {
        int c = pr[j];
        if (!c) continue;
        sRJ += (long long)c * (j + 1);
        sRJ2 += (long long)c * (j + 1) * (j + 1);
        sR2 += (long long)c * c;
        sre += (double)c / ((double)(j + 1) * (double)(j + 1));
    }
    sRJ = warp_sum_ll(sRJ);
    sRJ2 = warp_sum_ll(sRJ2);
    sR2 = warp_sum_ll(sR2);
    sre = warp_sum(sre);
    if (lane == 0) {
        const double N = (double)sN;
        o[0] = (double)sG2 / N;
        o[1] = (double)sG2 / (N * N);
        o[2] = (double)(sN * sGI2 - sGI * sGI) / (N * N);
        o[3] = (double)sGI2 / N;
        o[4] = (double)sRJ2 / N;
        o[5] = lrh / N;
        o[6] = lrl / N;
        o[7] = lgl / N;
        o[8] = log2(N) - e1 / N - RADB_EPS_LN2 * (double)nnz;
        o[9] = (double)sR2 / N;
        o[10] = (double)sR2 / (N * N);
        o[11] = N / (double)sRJ;
        o[12] = (double)(sN * sRJ2 - sRJ * sRJ) / (N * N);
        o[13] = sre / N;
        o[14] = srh / N;
        o[15] = srl / N;
    }
    return 1;
}

// ------------------------------------------------------------------ zone-like sums (GLSZM / GLDM)
struct ZoneSums {
    long long N, GI, GI2, G2, J1, J2, PJ2;
    double lgl, e1, small, sl, sh, ll, lh;
    int nnz;
};
__device__ __forceinline__ void zs_init(ZoneSums& z)
{
    z.N = z.GI = z.GI2 = z.G2 = z.J1 = z.J2 = z.PJ2 = 0;
    z.lgl = z.e1 = z.small = z.sl = z.sh = z.ll = z.lh = 0;
    z.nnz = 0;
}
// lane-local accumulation of one cell (level i (1-based), size j, count c)
__device__ __forceinline__ void zs_cell(ZoneSums& z, int i, int j, int c)
{
    double dc = (double)c, i2 = (double)i * (double)i, j2 = (double)j * (double)j;
    if (c > 1) z.e1 += dc * log2(dc);
    z.nnz++;
    z.small += dc / j2;
    z.sl += dc / (i2 * j2);
    z.sh += dc * i2 / j2;
    z.ll += dc * j2 / i2;
    z.lh += dc * i2 * j2;
    z.J1 += (long long)c * j;
    z.J2 += (long long)c * j * j;
}
__device__ __forceinline__ void zs_reduce(ZoneSums& z)
{
    z.N = warp_sum_ll(z.N);
    z.GI = warp_sum_ll(z.GI);
    z.GI2 = warp_sum_ll(z.GI2);
    z.G2 = warp_sum_ll(z.G2);
    z.J1 = warp_sum_ll(z.J1);
    z.J2 = warp_sum_ll(z.J2);
    z.PJ2 = warp_sum_ll(z.PJ2);
    z.lgl = warp_sum(z.lgl);
    z.e1 = warp_sum(z.e1);
    z.small = warp_sum(z.small);
    z.sl = warp_sum(z.sl);
    z.sh = warp_sum(z.sh);
    z.ll = warp_sum(z.ll);
    z.lh = warp_sum(z.lh);
    z.nnz = warp_sum_i(z.nnz);
}

// A.8 / pyradiomics glszm.py.  Dense counters Z[n][s0] (sizes 1..s0) + overflow list of
// (level << 16 | size) zones with size > s0.  pg = int scratch [n] (zeroed).
__device__ void glszm_task(const RadbParams& p, const int* Z, const unsigned* ovf, int novf, int n, int* pg,
                           double* o, int lane)
{
    const int s0 = p.s0;
    ZoneSums z;
    zs_init(z);
    // dense part: lane owns columns j = lane (+32..)
    for (int j = lane; j < s0; j += 32) {
        int cs = 0;
        for (int i = 0; i < n; i++) {
            int c = Z[i * s0 + j];
            if (!c) continue;
            cs += c;
            atomicAdd(&pg[i], c);
            zs_cell(z, i + 1, j + 1, c);
        }
        z.PJ2 += (long long)cs * cs;
    }
    // overflow zones: multiplicities by pairwise comparison (list is short: <= HW/(s0+1))
    for (int e = lane; e < novf; e += 32) {
        unsigned key = ovf[e];
        int sz = (int)(key & 0xffffu), lv = (int)(key >> 16);
        int same_key = 0, same_size = 0, first_key = 1, first_size = 1;
        for (int f = 0; f < novf; f++) {
            unsigned k2 = ovf[f];
            if (k2 == key) { same_key++; if (f < e) first_key = 0; }
            if ((int)(k2 & 0xffffu) == sz) { same_size++; if (f < e) first_size = 0; }
        }
        atomicAdd(&pg[lv - 1], 1);
        if (first_key) zs_cell(z, lv, sz, same_key);
        if (first_size) z.PJ2 += (long long)same_size * same_size;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        int g = pg[i];
        if (!g) continue;
        z.N += g;
        z.GI += (long long)g * (i + 1);
        z.GI2 += (long long)g * (i + 1) * (i + 1);
        z.G2 += (long long)g * g;
        z.lgl += (double)g / ((double)(i + 1) * (double)(i + 1));
    }
    zs_reduce(z);
    if (lane == 0) {
        const double N = z.N ? (double)z.N : 1.0;
        const double Np = z.J1 ? (double)z.J1 : 1.0;
        o[0] = (double)z.G2 / N;
        o[1] = (double)z.G2 / (N * N);
        o[2] = (double)(z.N * z.GI2 - z.GI * z.GI) / (N * N);
        o[3] = (double)z.GI2 / N;
        o[4] = (double)z.J2 / N;
        o[5] = z.lh / N;
        o[6] = z.ll / N;
        o[7] = z.lgl / N;
        o[8] = (double)z.PJ2 / N;
        o[9] = (double)z.PJ2 / (N * N);
        o[10] = z.small / N;
        o[11] = z.sh / N;
        o[12] = z.sl / N;
        o[13] = z.N ? log2(N) - z.e1 / N - RADB_EPS_LN2 * (double)z.nnz : 0.0;
        o[14] = N / Np;
        o[15] = (double)(z.N * z.J2 - z.J1 * z.J1) / (N * N);
    }
}

// A.9 / pyradiomics gldm.py.  D[n][nd] counters, column = dependence count (size j = col + 1).
__device__ void gldm_task(const int* D, int n, int nd, double* o, int lane)
{
    ZoneSums z;
    zs_init(z);
    int colsum[9];
#pragma unroll
    for (int j = 0; j < 9; j++) colsum[j] = 0;
    for (int i = lane; i < n; i += 32) {
        int g = 0;
#pragma unroll
        for (int j = 0; j < 9; j++) {
            if (j >= nd) break;
            int c = D[i * nd + j];
            if (!c) continue;
            g += c;
            colsum[j] += c;
            zs_cell(z, i + 1, j + 1, c);
        }
        if (g) {
            z.N += g;
            z.GI += (long long)g * (i + 1);
            z.GI2 += (long long)g * (i + 1) * (i + 1);
            z.G2 += (long long)g * g;
            z.lgl += (double)g / ((double)(i + 1) * (double)(i + 1));
        }
    }
#pragma unroll
    for (int j = 0; j < 9; j++) {
        int cs = warp_sum_i(colsum[j]);
        if (lane == 0) z.PJ2 += (long long)cs * cs;
    }
    zs_reduce(z);
    if (lane == 0) {
        const double N = z.N ? (double)z.N : 1.0;
        o[0] = z.N ? log2(N) - z.e1 / N - RADB_EPS_LN2 * (double)z.nnz : 0.0;
        o[1] = (double)z.PJ2 / N;
        o[2] = (double)z.PJ2 / (N * N);
        o[3] = (double)(z.N * z.J2 - z.J1 * z.J1) / (N * N);
        o[4] = (double)z.G2 / N;
        o[5] = (double)(z.N * z.GI2 - z.GI * z.GI) / (N * N);
        o[6] = (double)z.GI2 / N;
        o[7] = (double)z.J2 / N;
        o[8] = z.lh / N;
        o[9] = z.ll / N;
        o[10] = z.lgl / N;
        o[11] = z.small / N;
        o[12] = z.sh / N;
        o[13] = z.sl / N;
    }
}

// A.9 / pyradiomics ngtdm.py.  C[n][nb] = #voxels of level i with (col+1) valid neighbours,
// S[n][nb] = sum over those voxels of |(col+1)*i - sum(neighbour levels)| (integers, so the
// float sum s_i = sum_col S/(col+1) does not depend on the order voxels were visited).
__device__ void ngtdm_task(const int* C, const int* S, int n, int nb, double* pi, double* si, double* o,
                           int lane, int* dbg_n, double* dbg_s)
{
    long long nvp_l = 0;
    for (int i = lane; i < n; i += 32) {
        int ni = 0;
        double s = 0;
        for (int c = 0; c < nb; c++) {
            ni += C[i * nb + c];
            s += (double)S[i * nb + c] / (double)(c + 1);
        }
        pi[i] = (double)ni;
        si[i] = s;
        nvp_l += ni;
        if (dbg_n) { dbg_n[i] = ni; dbg_s[i] = s; }
    }
    const double Nvp = (double)warp_sum_ll(nvp_l);
    __syncwarp();
    if (Nvp == 0.0) {
        if (lane < 5) o[lane] = nan_f64();
        return;
    }
    for (int i = lane; i < n; i += 32) pi[i] = pi[i] / Nvp;
    __syncwarp();
    double sum_ps = 0, sum_s = 0, absd = 0, cplx = 0, contr = 0, stren = 0;
    int ngp = 0;
    for (int i = lane; i < n; i += 32) {
        double p_i = pi[i];
        if (p_i == 0.0) continue;
        ngp++;
        double s_i = si[i], di = (double)(i + 1);
        sum_ps += p_i * s_i;
        sum_s += s_i;
        for (int j = 0; j < n; j++) {
            double p_j = pi[j];
            if (p_j == 0.0) continue;
            double dj = (double)(j + 1), dd = di - dj;
            absd += fabs(di * p_i - dj * p_j);
            cplx += fabs(dd) * (p_i * s_i + p_j * si[j]) / (p_i + p_j);
            contr += p_i * p_j * dd * dd;
            stren += (p_i + p_j) * dd * dd;
        }
    }
    sum_ps = warp_sum(sum_ps);
    sum_s = warp_sum(sum_s);
    absd = warp_sum(absd);
    cplx = warp_sum(cplx);
    contr = warp_sum(contr);
    stren = warp_sum(stren);
    ngp = warp_sum_i(ngp);
    if (lane == 0) {
        double div = (double)ngp * (double)(ngp - 1);
        o[0] = (absd != 0.0) ? sum_ps / absd : 0.0;
        o[1] = (sum_ps != 0.0) ? 1.0 / sum_ps : 1e6;
        o[2] = cplx / Nvp;
        o[3] = (div != 0.0) ? contr * sum_s / Nvp / div : 0.0;
        o[4] = (sum_s != 0.0) ? stren / sum_s : 0.0;
    }
}

// ------------------------------------------------------------------ the per-patch CTA
template <typename PT>
__device__ void radb_cta(const RadbParams& p, long long patch, unsigned char* smem)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, W = p.W, HW = p.HW, WP = p.WP, NA = p.n_angles, NB = 2 * p.n_angles;
    PT* s_img = (PT*)(smem + p.o_stage);
    unsigned char* s_msk = smem + p.o_mask;
    unsigned short* lab = (unsigned short*)(smem + p.o_stage);
    unsigned char* lev = smem + p.o_lev;
    unsigned* zsize = (unsigned*)(smem + p.o_zsize);
    int* hist = (int*)(smem + p.o_hist);
    unsigned char* lut = smem + p.o_lut;
    int* lhist = (int*)(smem + p.o_lhist);
    int* glcm = (int*)(smem + p.o_glcm);
    unsigned* glrlm = (unsigned*)(smem + p.o_glrlm);
    int* gldm = (int*)(smem + p.o_gldm);
    int* ngc = (int*)(smem + p.o_ngc);
    int* ngn = (int*)(smem + p.o_ngn);
    int* szm = (int*)(smem + p.o_szm);
    unsigned* ovf = (unsigned*)(smem + p.o_ovf);
    double* fsc = (double*)(smem + p.o_fsc);
    int* misc = (int*)(smem + p.o_misc);
    const PT* g_img = (const PT*)((const unsigned char*)p.img + patch * p.img_stride);
    const unsigned char* g_msk = p.mask + patch * p.mask_stride;
    double* out = p.out + patch * (long long)p.F;

    // ---- phase 0: stage the patch, zero the counters
#ifndef RADB_EMU
    if (p.use_tma) {
        void* bar = smem + p.o_mbar;
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            const unsigned ib = (unsigned)(HW * sizeof(PT)), mb = (unsigned)HW;
            mbar_expect_tx(bar, ib + mb);
            tma_load_1d(s_img, g_img, ib, bar);
            tma_load_1d(s_msk, g_msk, mb, bar);
        }
    }
#endif
    {
        uint4* z = (uint4*)(smem + p.o_zero);
        const int nz = (p.smem_total - p.o_zero) / 16;
        const uint4 zero = {0u, 0u, 0u, 0u};
        for (int i = tid; i < nz; i += RADB_NT) z[i] = zero;
    }
#ifndef RADB_EMU
    if (p.use_tma) {
        mbar_wait(smem + p.o_mbar, 0);
    } else
#endif
    {
        for (int i = tid; i < HW; i += RADB_NT) {
            s_img[i] = g_img[i];
            s_msk[i] = g_msk[i];
        }
    }
    __syncthreads();

    // ---- phase 1: ROI histogram, bbox, voxel count
    {
        int np = 0, ymin = H, ymax = -1, xmin = W, xmax = -1;
        for (int y = warp; y < H; y += RADB_NT / 32)
            for (int x = lane; x < W; x += 32) {
                int i = y * W + x;
                if ((int)s_msk[i] == p.label) {
                    np++;
                    ymin = y < ymin ? y : ymin;
                    ymax = y > ymax ? y : ymax;
                    xmin = x < xmin ? x : xmin;
                    xmax = x > xmax ? x : xmax;
                    atomicAdd(&hist[(int)s_img[i]], 1);
                }
            }
        np = warp_sum_i(np);
        ymin = warp_min_i(ymin);
        ymax = warp_max_i(ymax);
        xmin = warp_min_i(xmin);
        xmax = warp_max_i(xmax);
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
    // ROI validity (A.1 step 2, imageoperations.checkMask)
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
        int vmin = 256, vmax = -1;
        for (int k = 0; k < 8; k++) {
            int v = lane * 8 + k;
            if (hist[v]) { vmin = v < vmin ? v : vmin; vmax = v > vmax ? v : vmax; }
        }
        vmin = warp_min_i(vmin);
        vmax = warp_max_i(vmax);
        // A.3 getBinEdges / binImage: level = #edges <= x, edges = low + k*binWidth
        int ng = 0;
        if (!st) {
            const double bw = p.bin_width;
            const double low = (double)vmin - py_mod((double)vmin, bw);
            ng = 0;
            for (int v = tid; v < 256; v += RADB_NT) {
                int L = 0;
                if (v >= vmin && v <= vmax) {
                    double x = (double)v;
                    long long k = (long long)floor((x - low) / bw);
                    while (low + (double)k * bw > x) k--;
                    while (low + (double)(k + 1) * bw <= x) k++;
                    L = (int)(k + 1);
                    if (L > 255) L = 255;  // reported through status 4 below
                }
                lut[v] = (unsigned char)L;
            }
            {
                double x = (double)vmax;
                long long k = (long long)floor((x - low) / bw);
                while (low + (double)k * bw > x) k--;
                while (low + (double)(k + 1) * bw <= x) k++;
                ng = (int)(k + 1);
            }
            if (ng > p.max_ng) st = 4;
        }
        if (st) {
            for (int f = tid; f < p.F; f += RADB_NT) out[f] = nan_f64();
            if (tid == 0) {
                p.status[patch] = st;
                if (p.dbg_ng) p.dbg_ng[patch] = 0;
            }
            return;
        }
        if (tid == 0) { misc[8] = ng; p.status[patch] = 0; }
    }
    __syncthreads();
    const int ng = misc[8];

    // ---- phase 2: discretised level image (padded, 0 outside the ROI) + level histogram
    for (int y = warp; y < H; y += RADB_NT / 32)
        for (int x = lane; x < W; x += 32) {
            int i = y * W + x;
            unsigned char L = ((int)s_msk[i] == p.label) ? lut[(int)s_img[i]] : (unsigned char)0;
            lev[(y + 1) * WP + x + 1] = L;
        }
    for (int v = tid; v < 256; v += RADB_NT)
        if (hist[v]) atomicAdd(&lhist[lut[v] - 1], hist[v]);
    __syncthreads();
    // the stage buffer is dead now: it becomes the union-find label array
    for (int i = tid; i < HW; i += RADB_NT) lab[i] = (unsigned short)i;
    __syncthreads();

    // ---- phase 3: neighbourhood pass -> GLCM, GLDM, NGTDM, GLRLM, zone unions
    {
        int doff[RADB_MAX_ANGLES], loff[RADB_MAX_ANGLES];
        for (int a = 0; a < RADB_MAX_ANGLES; a++) {
            doff[a] = a < NA ? p.ang_y[a] * WP + p.ang_x[a] : 0;
            loff[a] = a < NA ? p.ang_y[a] * W + p.ang_x[a] : 0;
        }
        const int nr = p.nr, nd = NB + 1;
        for (int y = warp; y < H; y += RADB_NT / 32)
            for (int x = lane; x < W; x += 32) {
                const int ctr = (y + 1) * WP + x + 1;
                const int c = lev[ctr];
                if (!c) continue;
                const int li = y * W + x;
                int dep = 0, cnt = 0, sum = 0;
#pragma unroll
                for (int a = 0; a < RADB_MAX_ANGLES; a++) {
                    if (a >= NA) break;
                    const int f = lev[ctr + doff[a]];
                    const int b = lev[ctr - doff[a]];
                    if (f) {
                        atomicAdd(&glcm[(a * ng + c - 1) * ng + f - 1], 1);
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
                    if (b == c) {
                        uf_union(lab, (unsigned)li, (unsigned)(li - loff[a]));
                    } else {
                        int len = 1, q = ctr + doff[a];
                        while (lev[q] == c) { len++; q += doff[a]; }
                        add_u16(glrlm, (a * ng + c - 1) * nr + len - 1);
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
    }
    __syncthreads();

    // ---- phase 4: flatten labels, zone sizes; symmetrise the GLCM
    for (int y = warp; y < H; y += RADB_NT / 32)
        for (int x = lane; x < W; x += 32) {
            if (!lev[(y + 1) * WP + x + 1]) continue;
            const int li = y * W + x;
            unsigned r = uf_find(lab, (unsigned)li);
            add_u16(zsize, (int)r);
        }
    if (p.symmetric) {
        for (int a = 0; a < NA; a++) {
            int* P = glcm + a * ng * ng;
            for (int t = tid; t < ng * ng; t += RADB_NT) {
                int i = t / ng, j = t - i * ng;
                if (i > j) continue;
                if (i == j) {
                    P[t] *= 2;
                } else {
                    int s = P[i * ng + j] + P[j * ng + i];
                    P[i * ng + j] = s;
                    P[j * ng + i] = s;
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 5: zones -> GLSZM (dense + overflow)
    for (int y = warp; y < H; y += RADB_NT / 32)
        for (int x = lane; x < W; x += 32) {
            const int c = lev[(y + 1) * WP + x + 1];
            if (!c) continue;
            const int li = y * W + x;
            if (lab[li] != (unsigned short)li) continue;
            const int s = get_u16(zsize, li);
            if (s <= p.s0) {
                atomicAdd(&szm[(c - 1) * p.s0 + s - 1], 1);
            } else {
                int k = atomicAdd(&misc[5], 1);
                if (k < p.ovf_cap) ovf[k] = ((unsigned)c << 16) | (unsigned)s;
            }
            if (p.dbg_glszm) atomicAdd(&p.dbg_glszm[(patch * p.max_ng + (c - 1)) * (long long)HW + s - 1], 1);
        }
    __syncthreads();

    // ---- optional debug dump of the integer matrices (parity tests)
    if (p.dbg_ng && tid == 0) p.dbg_ng[patch] = ng;
    if (p.dbg_levels)
        for (int i = tid; i < HW; i += RADB_NT)
            p.dbg_levels[patch * HW + i] = lev[(i / W + 1) * WP + (i % W) + 1];
    if (p.dbg_glcm)
        for (int a = 0; a < NA; a++)
            for (int t = tid; t < ng * ng; t += RADB_NT)
                p.dbg_glcm[((patch * NA + a) * p.max_ng + t / ng) * p.max_ng + t % ng] = glcm[a * ng * ng + t];
    if (p.dbg_glrlm)
        for (int a = 0; a < NA; a++)
            for (int t = tid; t < ng * p.nr; t += RADB_NT)
                p.dbg_glrlm[((patch * NA + a) * p.max_ng + t / p.nr) * p.nr + t % p.nr] =
                    get_u16(glrlm, a * ng * p.nr + t);
    if (p.dbg_gldm)
        for (int t = tid; t < ng * (NB + 1); t += RADB_NT)
            p.dbg_gldm[patch * p.max_ng * (NB + 1) + t] = gldm[t];

    // ---- phase 6: warp-specialised feature reductions
    //   tasks 0..NA-1      : GLCM features of angle t        -> fsc[t][0..23]
    //   tasks NA..2NA-1    : GLRLM features of angle t-NA    -> fsc[t-NA][24..39]
    //   then GLSZM, GLDM, NGTDM, first-order                 -> fsc[NA*40 + ...]
    double* single = fsc + NA * RADB_FSC_STRIDE;  // [0..15] glszm, [16..29] gldm, [30..34] ngtdm, [35..52] fo
    int* valid = misc + 16;                       // [a] glcm angle valid, [4+a] glrlm angle valid
    const int ntask = 2 * NA + 4;
    for (int t = warp; t < ntask; t += RADB_NT / 32) {
        if (t < NA) {
            const int a = t;
            int ok = glcm_task(p, glcm + a * ng * ng, ng, (int*)(smem + p.o_px) + a * ng,
                               (int*)(smem + p.o_py) + a * ng, (int*)(smem + p.o_padd) + a * 2 * ng,
                               (int*)(smem + p.o_psub) + a * ng, (double*)(smem + p.o_mcc) + a * p.mcc_stride,
                               smem + p.o_idx + a * ng, fsc + a * RADB_FSC_STRIDE, lane);
            if (lane == 0) valid[a] = ok;
        } else if (t < 2 * NA) {
            const int a = t - NA;
            int ok = glrlm_task(glrlm, a * ng * p.nr, ng, p.nr, (int*)(smem + p.o_pr) + a * p.nr,
                                fsc + a * RADB_FSC_STRIDE + RADB_GLCM_NF, lane);
            if (lane == 0) valid[4 + a] = ok;
        } else if (t == 2 * NA) {
            int novf = misc[5] < p.ovf_cap ? misc[5] : p.ovf_cap;
            glszm_task(p, szm, ovf, novf, ng, (int*)(smem + p.o_pg), single,
                       lane);
        } else if (t == 2 * NA + 1) {
            gldm_task(gldm, ng, NB + 1, single + 16, lane);
        } else if (t == 2 * NA + 2) {
            double* pi = (double*)(smem + p.o_ngp);
            ngtdm_task(ngc, ngn, ng, NB, pi, pi + ng, single + 30, lane,
                       p.dbg_ngn ? p.dbg_ngn + patch * p.max_ng : (int*)0,
                       p.dbg_ngs ? p.dbg_ngs + patch * p.max_ng : (double*)0);
        } else {
            fo_task_u8(p, hist, lhist, ng, (double*)(smem + p.o_qv), single + 35, lane);
        }
    }
    __syncthreads();

    // ---- phase 7: nanmean over angles, write the feature row
    {
        int nroi = 0;  // number of gray levels present in the ROI (MCC: < 2 -> 1)
        for (int i = 0; i < ng; i++) nroi += lhist[i] > 0;
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
        if (tid >= 64) {
            const int f = tid - 64;
            if (f < 16) { if (p.off_glszm >= 0) out[p.off_glszm + f] = single[f]; }
            else if (f < 30) { if (p.off_gldm >= 0) out[p.off_gldm + f - 16] = single[f]; }
            else if (f < 35) { if (p.off_ngtdm >= 0) out[p.off_ngtdm + f - 30] = single[f]; }
            else if (f < 53) { if (p.off_fo >= 0) out[p.off_fo + f - 35] = single[f]; }
        }
    }
}

#ifndef RADB_EMU
template <typename PT>
__global__ void __launch_bounds__(RADB_NT) radb_extract_kernel(const RadbParams p)
{
    extern __shared__ __align__(16) unsigned char radb_smem[];
    radb_cta<PT>(p, (long long)blockIdx.x, radb_smem);
}
#endif
