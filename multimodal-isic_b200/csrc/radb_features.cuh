// Per-matrix fp64 feature reductions (one warp per task).  Included by radb_kernels.cuh.
// Formulas: pyradiomics 3.1.0 firstorder.py / glcm.py / glrlm.py / glszm.py / gldm.py / ngtdm.py as
// restated in SURVEY.md Appendix A.5-A.9 (reached from /root/reference/RadiomicExtractor.py:38-48).
//
// Numerics shared by all tasks:
//  * counts are integers, so moments of counts are accumulated exactly in int64 and variances are
//    formed as (N*S2 - S1^2)/N^2;
//  * every entropy  -sum p*log2(p + eps)  with p = c/N is evaluated as
//        log2 N - (sum c*log2 c)/N - nnz*eps/ln2          (first-order in eps, error O(eps^2)),
//    with log2 c read from a per-CTA table for c < 128 (the GLCM sums c*(log2 c - log2 N) so that
//    HXY2 - HXY is exactly 0 for a one-level ROI, as it is upstream);
//  * 1/k^2 comes from a per-CTA table (no fp64 divisions inside the cell loops).
#pragma once

// Out-of-line fp64 helpers: log2 / the table fallbacks expand to 25-70 instructions each; inlining
// them at every call site is what made the kernels several times larger than the instruction cache.
__device__ __noinline__ double radb_log2(double x) { return log2(x); }
__device__ __noinline__ double radb_inv_sq(int k) { return 1.0 / ((double)k * (double)k); }
__device__ __noinline__ double radb_div(double a, double b) { return a / b; }
__device__ __noinline__ double radb_sqrt(double a) { return sqrt(a); }

#define RADB_TLOG_N 2048  // entries of the log2(count) table

// Read-only global loads (ld.global.nc): records and tables are written by earlier kernels / the host, so the
// compiler may hoist and batch these loads across the shared-memory stores in between.
#ifdef RADB_EMU
#define RADB_LDG(ptr) (*(ptr))
#else
#define RADB_LDG(ptr) __ldg(ptr)
#endif

struct RadbTabs {
    const double* inv2;  // inv2[k-1] = 1/k^2, k = 1..ninv
    int ninv;
    const double* tlog;  // tlog[c] = log2(c), c = 1..RADB_TLOG_N-1 (tlog[0] = 0)
    double* red;         // this warp's reduction scratch (RADB_RED_DOUBLES doubles, shared memory)
};
__device__ __forceinline__ double tab_inv2(const RadbTabs& t, int k)
{
    return k <= t.ninv ? RADB_LDG(&t.inv2[k - 1]) : radb_inv_sq(k);
}
__device__ __forceinline__ double tab_log2(const RadbTabs& t, int c)
{
    return c < RADB_TLOG_N ? RADB_LDG(&t.tlog[c]) : radb_log2((double)c);
}
__device__ __forceinline__ double tab_clog(const RadbTabs& t, int c) { return (double)c * tab_log2(t, c); }

// ------------------------------------------------------------------ first-order (u8 raw histogram)
// One warp.  A.5: everything except Entropy/Uniformity comes from the raw ROI values; for uint8
// pixels those are exactly the 256-bin histogram (lane l owns bins 8l..8l+7).
__device__ void fo_task_u8(const RadbParams& p, const RadbTabs& tb, const int* hist, const int* lhist, int ng, double* qv,
                           double* o, int lane)
{
    const int* h = hist + lane * 8;  // this lane's 8 bins (re-read in each pass: keeps the code small)
    long long s1 = 0;
    int n = 0, vmin = 256, vmax = -1;
#pragma unroll 1
    for (int k = 0; k < 8; k++) {
        const int v = lane * 8 + k, hk = h[k];
        n += hk;
        s1 += (long long)hk * v;
        vmin = (hk && v < vmin) ? v : vmin;
        vmax = (hk && v > vmax) ? v : vmax;
    }
    const int before = warp_excl_scan_i(n, lane);
    const int N = warp_sum_i(n);
    const double dN = (double)N, rN = radb_div(1.0, dN);
    const double mean = (double)warp_sum_ll(s1) * rN;
    const double shift = p.shift;
    vmin = warp_min_i(vmin);
    vmax = warp_max_i(vmax);
    // order statistics for the 10/25/50/75/90 percentiles (numpy 'linear' interpolation): rank r
    // lives in the lane whose cumulative range [before, before+n) contains it
    double fr[5];
#pragma unroll 1
    for (int q = 0; q < 5; q++) {
        const double qq = q == 0 ? 0.1 : q == 1 ? 0.25 : q == 2 ? 0.5 : q == 3 ? 0.75 : 0.9;
        const double pos = qq * (dN - 1.0);
        const double fl = floor(pos);
        if (lane == 0) qv[10 + q] = pos - fl;
        int lo = (int)fl;
        lo = lo > N - 1 ? N - 1 : lo;
        const int hi = lo + 1 > N - 1 ? N - 1 : lo + 1;
#pragma unroll 1
        for (int e = 0; e < 2; e++) {
            const int r = e ? hi : lo;
            if (r >= before && r < before + n) {
                int cum = before, bin = 0;
#pragma unroll 1
                for (int k = 0; k < 8; k++) {
                    bin = (r >= cum) ? k : bin;
                    cum += h[k];
                }
                qv[2 * q + e] = (double)(lane * 8 + bin);
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 5; q++) fr[q] = qv[2 * q] + (qv[2 * q + 1] - qv[2 * q]) * qv[10 + q];
    const double p10 = fr[0], p25 = fr[1], med = fr[2], p75 = fr[3], p90 = fr[4];
    // central moments, MAD, energy, robust MAD
    double m2 = 0, m3 = 0, m4 = 0, mad = 0, en = 0, in_s1 = 0;
    int in_n = 0;
#pragma unroll 1
    for (int k = 0; k < 8; k++) {
        const double v = (double)(lane * 8 + k), hk = (double)h[k];
        const double d = v - mean, d2 = d * d;
        m2 += hk * d2;
        m3 += hk * d2 * d;
        m4 += hk * d2 * d2;
        mad += hk * fabs(d);
        en += hk * (v + shift) * (v + shift);
        const bool in = (v >= p10) && (v <= p90);
        in_n += in ? h[k] : 0;
        in_s1 += in ? hk * v : 0.0;
    }
    {
        double r[6] = {m2, m3, m4, mad, en, in_s1};
        warp_sum_n(r, tb.red, lane);
        m2 = r[0] * rN; m3 = r[1] * rN; m4 = r[2] * rN; mad = r[3] * rN; en = r[4]; in_s1 = r[5];
    }
    const int inN = warp_sum_i(in_n);
    const double rin = radb_div(1.0, (double)inN);
    const double in_mean = in_s1 * rin;
    double rmad = 0;
#pragma unroll 1
    for (int k = 0; k < 8; k++) {
        const double v = (double)(lane * 8 + k);
        const bool in = (v >= p10) && (v <= p90);
        rmad += in ? (double)h[k] * fabs(v - in_mean) : 0.0;
    }
    // discretised histogram: Entropy / Uniformity (A.5: p = level histogram / N)
    double ent = 0, uni = 0;
#pragma unroll 1
    for (int i = lane; i < ng; i += 32) {
        const double pi = (double)lhist[i] * rN;
        if (lhist[i]) ent -= pi * radb_log2(pi + RADB_EPS);
        uni += pi * pi;
    }
    {
        double r[3] = {rmad, ent, uni};
        warp_sum_n(r, tb.red, lane);
        rmad = r[0] * rin; ent = r[1]; uni = r[2];
    }
    if (lane == 0) {
        o[0] = p10;
        o[1] = p90;
        o[2] = en;
        o[3] = ent;
        o[4] = p75 - p25;
        o[5] = (m2 == 0.0) ? 0.0 : radb_div(m4, m2 * m2);
        o[6] = (double)vmax;
        o[7] = mad;
        o[8] = mean;
        o[9] = med;
        o[10] = (double)vmin;
        o[11] = (double)(vmax - vmin);
        o[12] = rmad;
        o[13] = radb_sqrt(en * rN);
        o[14] = (m2 == 0.0) ? 0.0 : radb_div(m3, m2 * radb_sqrt(m2));
        o[15] = en;  // TotalEnergy: pixel spacing is (1, 1) for GetImageFromArray images
        o[16] = uni;
        o[17] = m2;
    }
}

// ------------------------------------------------------------------ MCC
// packed lower-triangular symmetric matrix: element (r, c), r >= c
__device__ __forceinline__ int tri(int r, int c) { return r * (r + 1) / 2 + c; }
__device__ __forceinline__ double sym_get(const double* M, int r, int c)
{
    return r >= c ? M[tri(r, c)] : M[tri(c, r)];
}

__device__ __forceinline__ int sturm_count(const double* d, const double* e2, int m, double x)
{
    // p_i = det(T_i - x I): p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2}; #sign changes = #eigenvalues < x.
    // Division-free; a zero inherits the sign of its predecessor; rescaled to stay in range.
    double p0 = 1.0, p1 = d[0] - x;
    bool neg = p1 < 0.0;
    int cnt = neg ? 1 : 0;
    int i = 1;
#pragma unroll 1
    for (; i + 1 < m; i += 2) {  // two steps per iteration, one rescale (|p| changes by < 1e17 per step)
        const double pa = (d[i] - x) * p1 - e2[i - 1] * p0;
        const bool na = (pa < 0.0) || (pa == 0.0 && neg);
        const double pb = (d[i + 1] - x) * pa - e2[i] * p1;
        const bool nb = (pb < 0.0) || (pb == 0.0 && na);
        cnt += (na != neg) + (nb != na);
        neg = nb;
        const double ap = fabs(pb);
        const double sc = ap > 1e100 ? 1e-100 : (ap < 1e-100 ? 1e100 : 1.0);
        p0 = pa * sc;
        p1 = pb * sc;
    }
    if (i < m) {
        const double pn = (d[i] - x) * p1 - e2[i - 1] * p0;
        const bool nneg = (pn < 0.0) || (pn == 0.0 && neg);
        cnt += (nneg != neg) ? 1 : 0;
    }
    return cnt;
}

// k-th smallest eigenvalue (k = 0..m-1) by warp multisection: 32 shifts per round
__device__ double tridiag_kth(const double* d, const double* e2, int m, int k, double lo, double hi,
                              int lane)
{
    for (int it = 0; it < 7; it++) {
        double w = (hi - lo) * (1.0 / 33.0);
        double x = lo + w * (double)(lane + 1);
        int c = sturm_count(d, e2, m, x);
        // lanes with count <= k are left of (or at) the eigenvalue; they form a prefix
        unsigned left = __ballot_sync(FULLMASK, c <= k);
        int nl = __popc(left);
        double nlo = lo + w * (double)nl;
        double nhi = (nl == 32) ? hi : lo + w * (double)(nl + 1);
        lo = nlo;
        hi = nhi;
        if (hi - lo <= 1e-10) break;  // eigenvalues live in [-1, 1] (33^7 = 4e10 shrink); the feature needs rtol 1e-6
    }
    return 0.5 * (lo + hi);
}

// One warp per angle.  A.6: MCC = radb_sqrt(second largest eigenvalue of Q),
// Q[i][j] = sum_k P[i][k] P[j][k] / (px[i] py[k]); Q is similar to S = A A^T with
// A = Dx^-1/2 P Dy^-1/2.  For a symmetric GLCM A is symmetric, so the eigenvalues of S are the
// squares of those of A and radb_sqrt(lambda_2(S)) = second largest |lambda(A)|.
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
    // d doubles as scratch for 1/radb_sqrt(px) while the matrix is built (d is first written below)
    for (int r = lane; r < m; r += 32) d[r] = radb_div(1.0, radb_sqrt((double)px[idx[r]]));
    __syncwarp();
    const int bgw = m <= 8 ? 8 : (m <= 16 ? 16 : 32), bng = 32 / bgw;
    for (int r0 = 0; r0 < m; r0 += bng) {
        const int r = r0 + lane / bgw;
        if (r >= m) continue;
        const int ir = idx[r];
        const double rr = d[r];
        for (int c = lane & (bgw - 1); c <= r; c += bgw) {
            const int ic = idx[c];
            double val;
            if (symmetric) {
                val = (double)P[ir * n + ic];
            } else {
                val = 0;
                for (int k = 0; k < n; k++)
                    if (py[k] > 0) val += radb_div((double)P[ir * n + k] * (double)P[ic * n + k], (double)py[k]);
            }
            M[tri(r, c)] = val * rr * d[c];
        }
    }
    __syncwarp();
    // Householder tridiagonalisation, column k eliminates rows k+2..m-1.  Small matrices (m <= 16) would
    // leave most lanes idle with one lane per row: L = 4 / 2 / 1 consecutive lanes share a row and split
    // its columns (c = first + sub, first + sub + L, ...), combined with xor shuffles.
    const int L = m <= 8 ? 4 : (m <= 16 ? 2 : 1);
    const int rpi = 32 / L, roff = lane / L, sub = lane & (L - 1);
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
        double nrm = radb_sqrt(tail + x0 * x0);
        double alpha = x0 > 0 ? -nrm : nrm;
        // v = x - alpha e1 (indices k+1..m-1), beta = 2 / v^T v
        double vtv = tail + (x0 - alpha) * (x0 - alpha);
        double beta = radb_div(2.0, vtv);
        for (int r = k + 1 + lane; r < m; r += 32) v[r] = (r == k + 1) ? x0 - alpha : M[tri(r, k)];
        __syncwarp();
        // w = beta * M22 v
        double kpart = 0;
        for (int r0 = k + 1; r0 < m; r0 += rpi) {
            const int r = r0 + roff;
            double s = 0;
            if (r < m) {
                const int rb = tri(r, 0);
                for (int c = k + 1 + sub; c < m; c += L) s += (c <= r ? M[rb + c] : M[tri(c, r)]) * v[c];
            }
            for (int mm = 1; mm < L; mm <<= 1) s += __shfl_xor_sync(FULLMASK, s, mm);
            if (r < m && sub == 0) {
                s *= beta;
                w[r] = s;
                kpart += s * v[r];
            }
        }
        const double K = 0.5 * beta * warp_sum(kpart);
        __syncwarp();
        // M22 -= v w'^T + w' v^T with w' = w - K v (lower triangle)
        for (int r0 = k + 1; r0 < m; r0 += rpi) {
            const int r = r0 + roff;
            if (r >= m) continue;
            const double vr = v[r], wr = w[r] - K * vr;
            const int rb = tri(r, 0);
            for (int c = k + 1 + sub; c <= r; c += L) {
                const double vc = v[c];
                M[rb + c] -= vr * (w[c] - K * vc) + wr * vc;
            }
        }
        __syncwarp();
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
        double r = (i > 0 ? radb_sqrt(e2[i - 1]) : 0.0) + (i < m - 1 ? radb_sqrt(e2[i]) : 0.0);
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
        // second largest |lambda(A)|: lambda_min matters only if it lies below -|lambda_2|
        const double t = fabs(l2);
        if (sturm_count(d, e2, m, -t * (1.0 + 1e-9) - 1e-12) == 0) return t;
        const double l1 = tridiag_kth(d, e2, m, 0, glo, ghi, lane);
        return fmax(t, fabs(l1));
    }
    return radb_sqrt(fmax(l2, 0.0));
}

// ------------------------------------------------------------------ MCC, four angles per warp
// With one warp per matrix most of mcc_task is per-step overhead (two warp reductions, a square root, a
// division and three warp barriers per Householder step, 7 rounds of Sturm counts) paid for by 32 lanes of
// which m - k - 1 <= 25 have a row to work on.  Here ONE warp solves the eigenproblems of all four angles at
// once (radb_mcc_g8_kernel): the aligned 8-lane group g owns angle g, rows are dealt to its 8 lanes, and every
// collective runs on the group's own mask, so the per-step overhead is shared by four matrices (ncu: 2.4x
// fewer warp instructions at Ng 26).  Same mathematics as mcc_task.
__device__ __forceinline__ double grp8_sum(double v, unsigned gm)
{
#pragma unroll
    for (int m = 4; m >= 1; m >>= 1) v += __shfl_xor_sync(gm, v, m);
    return v;
}
__device__ double tridiag_kth_g8(const double* d, const double* e2, int m, int k, double lo, double hi, int gl, unsigned gm,
                                 int gshift)
{
    for (int it = 0; it < 12; it++) {  // 9-section: 9^11 = 3e10 shrink
        const double w = (hi - lo) * (1.0 / 9.0);
        const double x = lo + w * (double)(gl + 1);
        const int c = sturm_count(d, e2, m, x);
        const unsigned left = (__ballot_sync(gm, c <= k) >> gshift) & 0xffu;  // a prefix of the group
        const int nl = __popc(left);
        const double nlo = lo + w * (double)nl;
        const double nhi = (nl == 8) ? hi : lo + w * (double)(nl + 1);
        lo = nlo;
        hi = nhi;
        if (hi - lo <= 1e-10) break;
    }
    return 0.5 * (lo + hi);
}
__device__ double mcc_task_g8(const int* P, const int* px, const int* py, int n, int symmetric, double* ws,
                              unsigned char* idx, int gl, unsigned gm, int gshift)
{
    int m = 0;
    for (int base = 0; base < n; base += 8) {
        const int i = base + gl;
        const int present = (i < n) && (px[i] > 0);
        const unsigned b = (__ballot_sync(gm, present) >> gshift) & 0xffu;
        if (present) idx[m + __popc(b & ((1u << gl) - 1u))] = (unsigned char)i;
        m += __popc(b);
    }
    __syncwarp(gm);
    if (m < 2) return 0.0;
    // d shares its array with v and e2 with w: step k writes d[k], e2[k] and uses v[r], w[r] for r > k only
    // (tri(m) + 2m doubles per matrix instead of tri(m) + 4m: one more resident CTA per SM at Ng 26)
    double* M = ws;
    double* v = ws + m * (m + 1) / 2;
    double* w = v + m;
    double* d = v;
    double* e2 = w;
    for (int r = gl; r < m; r += 8) d[r] = radb_div(1.0, radb_sqrt((double)px[idx[r]]));
    __syncwarp(gm);
    for (int r = gl; r < m; r += 8) {
        const int ir = idx[r];
        const double rr = d[r];
        for (int c = 0; c <= r; c++) {
            const int ic = idx[c];
            double val;
            if (symmetric) {
                val = (double)P[ir * n + ic];
            } else {
                val = 0;
                for (int k = 0; k < n; k++)
                    if (py[k] > 0) val += radb_div((double)P[ir * n + k] * (double)P[ic * n + k], (double)py[k]);
            }
            M[tri(r, c)] = val * rr * d[c];
        }
    }
    __syncwarp(gm);
    for (int k = 0; k < m - 2; k++) {
        double part = 0;
        for (int r = k + 2 + gl; r < m; r += 8) { const double x = M[tri(r, k)]; part += x * x; }
        const double tail = grp8_sum(part, gm);
        const double x0 = M[tri(k + 1, k)];
        if (tail == 0.0) {
            if (gl == 0) { d[k] = M[tri(k, k)]; e2[k] = x0 * x0; }
            __syncwarp(gm);
            continue;
        }
        const double nrm = radb_sqrt(tail + x0 * x0);
        const double alpha = x0 > 0 ? -nrm : nrm;
        const double vtv = tail + (x0 - alpha) * (x0 - alpha);
        const double beta = radb_div(2.0, vtv);
        for (int r = k + 1 + gl; r < m; r += 8) v[r] = (r == k + 1) ? x0 - alpha : M[tri(r, k)];
        __syncwarp(gm);
        double kpart = 0;
        for (int r = k + 1 + gl; r < m; r += 8) {
            const int rb = tri(r, 0);
            double s = 0;
            for (int c = k + 1; c <= r; c++) s += M[rb + c] * v[c];
            for (int c = r + 1; c < m; c++) s += M[tri(c, r)] * v[c];
            s *= beta;
            w[r] = s;
            kpart += s * v[r];
        }
        const double K = 0.5 * beta * grp8_sum(kpart, gm);
        __syncwarp(gm);
        for (int r = k + 1 + gl; r < m; r += 8) {
            const double vr = v[r], wr = w[r] - K * vr;
            const int rb = tri(r, 0);
            for (int c = k + 1; c <= r; c++) {
                const double vc = v[c];
                M[rb + c] -= vr * (w[c] - K * vc) + wr * vc;
            }
        }
        __syncwarp(gm);
        if (gl == 0) { d[k] = M[tri(k, k)]; e2[k] = alpha * alpha; }
        __syncwarp(gm);
    }
    if (gl == 0) {
        d[m - 2] = M[tri(m - 2, m - 2)];
        const double x = M[tri(m - 1, m - 2)];
        e2[m - 2] = x * x;
        d[m - 1] = M[tri(m - 1, m - 1)];
    }
    __syncwarp(gm);
    double glo = 1e300, ghi = -1e300;
    for (int i = gl; i < m; i += 8) {
        const double r = (i > 0 ? radb_sqrt(e2[i - 1]) : 0.0) + (i < m - 1 ? radb_sqrt(e2[i]) : 0.0);
        glo = fmin(glo, d[i] - r);
        ghi = fmax(ghi, d[i] + r);
    }
#pragma unroll
    for (int s = 4; s >= 1; s >>= 1) {
        glo = fmin(glo, __shfl_xor_sync(gm, glo, s));
        ghi = fmax(ghi, __shfl_xor_sync(gm, ghi, s));
    }
    const double span = ghi - glo;
    glo -= 1e-12 * (span + 1.0);
    ghi += 1e-12 * (span + 1.0);
    const double l2 = tridiag_kth_g8(d, e2, m, m - 2, glo, ghi, gl, gm, gshift);
    if (symmetric) {
        const double t = fabs(l2);
        if (sturm_count(d, e2, m, -t * (1.0 + 1e-9) - 1e-12) == 0) return t;
        const double l1 = tridiag_kth_g8(d, e2, m, 0, glo, ghi, gl, gm, gshift);
        return fmax(t, fabs(l1));
    }
    return radb_sqrt(fmax(l2, 0.0));
}

// ------------------------------------------------------------------ GLCM features (one warp, one angle)
// A.6.  P holds the final integer counts of this angle (symmetrised when symmetricalGLCM).
// Returns 0 when the angle is empty (upstream deletes it from the nanmean).
// `mcc_pre`: the MCC of this angle when radb_mcc_lanczos_kernel has already computed it (null: computed here).
__device__ int glcm_task(const RadbParams& p, const RadbTabs& tb, const int* P, int n, int* px, int* py, int* padd,
                         int* psub, double* ws, unsigned char* idx, double* o, int lane, const double* mcc_pre = (const double*)0)
{
    const bool sym = p.symmetric != 0;
    if (sym) py = px;  // symmetric matrix: identical marginals
    long long sN = 0, sI = 0, sJ = 0, sIJ = 0, sD2 = 0, sC2 = 0;
    double sclog = 0;
    int maxc = 0, nnz = 0;
    // Small matrices (n <= 16: binWidth 25 gives n <= 11) would leave most lanes idle with one row per
    // warp iteration: the warp is split into 32/gw groups of gw lanes and every group takes its own row.
    const int gw = n <= 8 ? 8 : (n <= 16 ? 16 : 32);
    const int ngrp = 32 / gw, gi = lane / gw, gl = lane & (gw - 1);
    for (int i0 = 0; i0 < n; i0 += ngrp) {
        const int i = i0 + gi;
        int rs = 0;
        if (i < n)
            for (int j = gl; j < n; j += gw) {
                const int c = P[i * n + j];
                if (c) {
                    rs += c;
                    sIJ += (long long)c * (i + 1) * (j + 1);
                    sD2 += (long long)c * (i - j) * (i - j);
                    sC2 += (long long)c * c;
                    sJ += (long long)c * (j + 1);
                    nnz++;
                    maxc = c > maxc ? c : maxc;
                    if (!sym) atomicAdd(&py[j], c);
                    atomicAdd(&padd[i + j], c);
                    atomicAdd(&psub[i > j ? i - j : j - i], c);
                }
            }
        rs = group_sum_i(rs, gw, lane);
        if (gl == 0 && i < n) {
            px[i] = rs;
            sN += rs;
            sI += (long long)rs * (i + 1);
        }
    }
    sN = warp_sum_ll(sN);
    __syncwarp();
    if (sN == 0) return 0;
    const double N = (double)sN, rN = radb_div(1.0, N);
    double r0[5] = {(double)sI, (double)sJ, (double)sIJ, (double)sD2, (double)sC2};  // exact integers < 2^53
    warp_sum_n(r0, tb.red, lane);
    const double ux = radb_div(r0[0], N);  // true divisions: exactly 1 for a one-level matrix (sigma = 0, Correlation = 1)
    const double uy = radb_div(r0[1], N);
    const double autoc = r0[2] * rN;
    const double contrast = r0[3] * rN;
    const double energy = r0[4] * rN * rN;
    const double maxp = (double)warp_max_i(maxc) * rN;
    // every log2 of an integer count goes through tab_log2 (same source for c, px and N), so that
    // c*(log2 c - log2 N) and the marginal terms cancel exactly for a one-level ROI
    const double log2N = tab_log2(tb, (int)sN);
    nnz = warp_sum_i(nnz);
    // marginals: entropies and reciprocal tables (ws is free until mcc_task builds its matrix)
    double* rpx = ws;       // N / px[i]
    double* rpy = ws + n;   // 1 / py[j]
    double hx = 0, hy = 0, hx0 = 0, hy0 = 0;
    int nx = 0, ny = 0;
    for (int i = lane; i < n; i += 32) {
        const int a = px[i], b = py[i];
        rpx[i] = a ? radb_div(N, (double)a) : 0.0;
        rpy[i] = b ? radb_div(1.0, (double)b) : 0.0;
        if (a) {
            hx0 -= (double)a * rN * (tab_log2(tb, a) - log2N);
            nx++;
        }
        if (b) {
            hy0 -= (double)b * rN * (tab_log2(tb, b) - log2N);
            ny++;
        }
    }
    __syncwarp();
    {
        double r1[2] = {hx0, hy0};
        warp_sum_n(r1, tb.red, lane);
        hx0 = r1[0]; hy0 = r1[1];
    }
    nx = warp_sum_i(nx);
    ny = warp_sum_i(ny);
    // HX = -sum px*log2(px + eps) = HX0 - nx*eps/ln2 to first order in eps (as for HXY)
    hx = hx0 - RADB_EPS_LN2 * (double)nx;
    hy = hy0 - RADB_EPS_LN2 * (double)ny;
    // pass B: cluster moments and correlation terms
    double ct = 0, cs = 0, cp = 0, ssq = 0, ssqy = 0, corm = 0, h1corr = 0;
    for (int i0 = 0; i0 < n; i0 += ngrp) {
        const int i = i0 + gi;
        if (i >= n) continue;
        const double di = (double)(i + 1) - ux;
        const double rpxi = rpx[i];
        for (int j = gl; j < n; j += gw) {
            const int c = P[i * n + j];
            if (!c) continue;
            const double dc = (double)c;
            const double pij = dc * rN;
            const double dj = (double)(j + 1) - uy;
            const double s = (double)(i + 1) + (double)(j + 1) - ux - uy;
            const double s2 = s * s;
            ct += pij * s2;
            cs += pij * s2 * s;
            cp += pij * s2 * s2;
            ssq += pij * di * di;
            ssqy += pij * dj * dj;
            corm += pij * di * dj;
            h1corr += dc * rpxi * rpy[j];  // p / (px*py)
            sclog += dc * (tab_log2(tb, c) - log2N);
        }
    }
    {
        double r2[8] = {ct, cs, cp, ssq, ssqy, corm, h1corr, sclog};
        warp_sum_n(r2, tb.red, lane);
        ct = r2[0]; cs = r2[1]; cp = r2[2]; ssq = r2[3]; ssqy = r2[4]; corm = r2[5]; h1corr = r2[6]; sclog = r2[7];
    }
    const double hxy = -sclog * rN - RADB_EPS_LN2 * (double)nnz;
    // log2(px*py + eps) = log2 px + log2 py + eps/(px*py*ln2) + O(eps^2): HXY1/HXY2 in closed form
    const double hxy1 = hx0 + hy0 - RADB_EPS_LN2 * h1corr;
    const double hxy2 = hx0 + hy0 - RADB_EPS_LN2 * (double)nx * (double)ny;
    // |i-j| marginal
    double da = 0, de = 0, idv = 0, idm = 0, idmn = 0, idn = 0, iv = 0;
    const double rdn = radb_div(1.0, (double)n);
    for (int k = lane; k < n; k += 32) {
        if (!psub[k]) continue;
        const double q = (double)psub[k] * rN, dk = (double)k;
        da += dk * q;
        de -= q * (tab_log2(tb, psub[k]) - log2N) + RADB_EPS_LN2;
        idv += radb_div(q, 1.0 + dk);
        idm += radb_div(q, 1.0 + dk * dk);
        idmn += radb_div(q, 1.0 + (dk * dk) * rdn * rdn);
        idn += radb_div(q, 1.0 + dk * rdn);
        if (k > 0) iv += q * tab_inv2(tb, k);
    }
    // i+j marginal (index k <-> i+j = k+2)
    double sa = 0, se = 0;
    for (int k = lane; k < 2 * n - 1; k += 32) {
        if (!padd[k]) continue;
        const double q = (double)padd[k] * rN;
        sa += (double)(k + 2) * q;
        se -= q * (tab_log2(tb, padd[k]) - log2N) + RADB_EPS_LN2;
    }
    {
        double r3[9] = {da, de, idv, idm, idmn, idn, iv, sa, se};
        warp_sum_n(r3, tb.red, lane);
        da = r3[0]; de = r3[1]; idv = r3[2]; idm = r3[3]; idmn = r3[4]; idn = r3[5]; iv = r3[6]; sa = r3[7]; se = r3[8];
    }
    double dvar = 0;
    for (int k = lane; k < n; k += 32)
        if (psub[k]) dvar += (double)psub[k] * rN * ((double)k - da) * ((double)k - da);
    dvar = warp_sum(dvar);
    __syncwarp();
    const double mcc = mcc_pre ? *mcc_pre : mcc_task(P, px, py, n, p.symmetric, ws, idx, lane);
    if (lane == 0) {
        const double sigx = radb_sqrt(ssq), sigy = radb_sqrt(ssqy);
        const double div = fmax(hx, hy);
        double im2 = 1.0 - exp(-2.0 * (hxy2 - hxy));
        im2 = im2 < 0.0 ? 0.0 : im2;
        o[0] = autoc;
        o[1] = cp;
        o[2] = cs;
        o[3] = ct;
        o[4] = contrast;
        o[5] = (sigx * sigy == 0.0) ? 1.0 : radb_div(corm, sigx * sigy + RADB_EPS);
        o[6] = da;
        o[7] = de;
        o[8] = dvar;
        o[9] = idv;
        o[10] = idm;
        o[11] = idmn;
        o[12] = idn;
        o[13] = (div != 0.0) ? radb_div(hxy - hxy1, div) : 0.0;
        o[14] = radb_sqrt(im2);
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
// A.7.  R = run counters [n][nr] (packed u16, or u32 in wide mode); pr = int scratch [nr] (zeroed).
__device__ int glrlm_task(const RadbTabs& tb, const unsigned* R, int wide, int n, int nr, int maxlen, int* pr,
                          double* o, int lane)
{
    long long sN = 0, sGI = 0, sGI2 = 0, sG2 = 0;
    double lgl = 0, e1 = 0, srl = 0, srh = 0, lrl = 0, lrh = 0;
    int nnz = 0;
    // only run lengths 1..maxlen occur (recorded by the build kernel); short maxima let several rows
    // share one warp iteration (groups of gw lanes, one row per group)
    const int gw = maxlen <= 8 ? 8 : (maxlen <= 16 ? 16 : 32);
    const int ngrp = 32 / gw, gi = lane / gw, gl = lane & (gw - 1);
    for (int i0 = 0; i0 < n; i0 += ngrp) {
        const int i = i0 + gi;
        int rs = 0;
        const double i2 = (double)(i + 1) * (double)(i + 1), ri2 = tab_inv2(tb, i < n ? i + 1 : 1);
        for (int j = gl; j < maxlen && i < n; j += gw) {
            const int c = get_run(R, i * nr + j, wide);
            if (!c) continue;
            rs += c;
            if (ngrp == 1) pr[j] += c; else atomicAdd(&pr[j], c);
            const double dc = (double)c, j2 = (double)(j + 1) * (double)(j + 1), rj2 = tab_inv2(tb, j + 1);
            e1 += tab_clog(tb, c);
            nnz++;
            srl += dc * ri2 * rj2;
            srh += dc * i2 * rj2;
            lrl += dc * j2 * ri2;
            lrh += dc * i2 * j2;
        }
        rs = group_sum_i(rs, gw, lane);
        if (gl == 0 && rs) {
            sN += rs;
            sGI += (long long)rs * (i + 1);
            sGI2 += (long long)rs * (i + 1) * (i + 1);
            sG2 += (long long)rs * rs;
            lgl += (double)rs * ri2;
        }
    }
    sN = warp_sum_ll(sN);
    __syncwarp();
    if (sN == 0) return 0;
    {
        double r0[9] = {(double)sGI, (double)sGI2, (double)sG2, lgl, e1, srl, srh, lrl, lrh};  // first three: exact
        warp_sum_n(r0, tb.red, lane);
        sGI = (long long)r0[0]; sGI2 = (long long)r0[1]; sG2 = (long long)r0[2];
        lgl = r0[3]; e1 = r0[4]; srl = r0[5]; srh = r0[6]; lrl = r0[7]; lrh = r0[8];
    }
    nnz = warp_sum_i(nnz);
    long long sRJ = 0, sRJ2 = 0, sR2 = 0;
    double sre = 0;
    for (int j = lane; j < maxlen; j += 32) {
        const int c = pr[j];
        if (!c) continue;
        sRJ += (long long)c * (j + 1);
        sRJ2 += (long long)c * (j + 1) * (j + 1);
        sR2 += (long long)c * c;
        sre += (double)c * tab_inv2(tb, j + 1);
    }
    {
        double r1[4] = {(double)sRJ, (double)sRJ2, (double)sR2, sre};
        warp_sum_n(r1, tb.red, lane);
        sRJ = (long long)r1[0]; sRJ2 = (long long)r1[1]; sR2 = (long long)r1[2]; sre = r1[3];
    }
    if (lane == 0) {
        const double N = (double)sN, rN = radb_div(1.0, N);
        o[0] = (double)sG2 * rN;
        o[1] = (double)sG2 * rN * rN;
        o[2] = (double)(sN * sGI2 - sGI * sGI) * rN * rN;
        o[3] = (double)sGI2 * rN;
        o[4] = (double)sRJ2 * rN;
        o[5] = lrh * rN;
        o[6] = lrl * rN;
        o[7] = lgl * rN;
        o[8] = radb_log2(N) - e1 * rN - RADB_EPS_LN2 * (double)nnz;
        o[9] = (double)sR2 * rN;
        o[10] = (double)sR2 * rN * rN;
        o[11] = radb_div(N, (double)sRJ);
        o[12] = (double)(sN * sRJ2 - sRJ * sRJ) * rN * rN;
        o[13] = sre * rN;
        o[14] = srh * rN;
        o[15] = srl * rN;
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
__device__ __forceinline__ void zs_cell(ZoneSums& z, const RadbTabs& tb, int i, int j, int c)
{
    const double dc = (double)c, i2 = (double)i * (double)i, j2 = (double)j * (double)j;
    const double ri2 = tab_inv2(tb, i), rj2 = tab_inv2(tb, j);
    z.e1 += tab_clog(tb, c);
    z.nnz++;
    z.small += dc * rj2;
    z.sl += dc * ri2 * rj2;
    z.sh += dc * i2 * rj2;
    z.ll += dc * j2 * ri2;
    z.lh += dc * i2 * j2;
    z.J1 += (long long)c * j;
    z.J2 += (long long)c * j * j;
}
__device__ __forceinline__ void zs_level(ZoneSums& z, const RadbTabs& tb, int i, int g)
{
    z.N += g;
    z.GI += (long long)g * i;
    z.GI2 += (long long)g * i * i;
    z.G2 += (long long)g * g;
    z.lgl += (double)g * tab_inv2(tb, i);
}
__device__ __forceinline__ void zs_reduce(ZoneSums& z, const RadbTabs& tb, int lane)
{
    double r[14] = {(double)z.N, (double)z.GI, (double)z.GI2, (double)z.G2, (double)z.J1, (double)z.J2,
                    (double)z.PJ2, z.lgl, z.e1, z.small, z.sl, z.sh, z.ll, z.lh};  // first seven: exact integers
    warp_sum_n(r, tb.red, lane);
    z.N = (long long)r[0]; z.GI = (long long)r[1]; z.GI2 = (long long)r[2]; z.G2 = (long long)r[3];
    z.J1 = (long long)r[4]; z.J2 = (long long)r[5]; z.PJ2 = (long long)r[6];
    z.lgl = r[7]; z.e1 = r[8]; z.small = r[9]; z.sl = r[10]; z.sh = r[11]; z.ll = r[12]; z.lh = r[13];
    z.nnz = warp_sum_i(z.nnz);
}

// A.8.  Dense counters Z[n][s0] (sizes 1..s0) + overflow list of ((level - 1) << 24 | size) zones with
// size > s0.  pg = int scratch [n] (zeroed), sorted = scratch for the rank-sorted overflow list.
__device__ void glszm_task(const RadbParams& p, const RadbTabs& tb, const int* Z, const unsigned* ovf,
                           unsigned* sorted, int novf, int n, int* pg, double* o, int lane)
{
    const int s0 = p.s0;
    ZoneSums z;
    zs_init(z);
    // dense part: cells linearised over the warp; pg / column sums through integer shared atomics
    for (int t = lane; t < n * s0; t += 32) {
        const int c = Z[t];
        if (!c) continue;
        const int i = t / s0, j = t - i * s0;
        atomicAdd(&pg[i], c);
        zs_cell(z, tb, i + 1, j + 1, c);
    }
    for (int j = lane; j < s0; j += 32) {
        int cs = 0;
        for (int i = 0; i < n; i++) cs += Z[i * s0 + j];
        z.PJ2 += (long long)cs * cs;
    }
    // overflow zones (list is short: <= HW/(s0+1)).  The list was appended in atomic order, so
    // first rank-sort it by key: every later sum then runs in an order that depends on the data
    // only (bit-reproducible output), and equal keys become adjacent.
    for (int e = lane; e < novf; e += 32) {
        unsigned key = ovf[e];
        int rank = 0;
        for (int f = 0; f < novf; f++) {
            unsigned k2 = ovf[f];
            rank += (k2 < key) || (k2 == key && f < e);
        }
        sorted[rank] = key;
    }
    __syncwarp();
    for (int e = lane; e < novf; e += 32) {
        unsigned key = sorted[e];
        int sz = (int)(key & 0xffffffu), lv = (int)(key >> 24) + 1;
        atomicAdd(&pg[lv - 1], 1);
        if (e == 0 || sorted[e - 1] != key) {  // first of its (level, size) cell
            int cnt = 1;
            while (e + cnt < novf && sorted[e + cnt] == key) cnt++;
            zs_cell(z, tb, lv, sz, cnt);
        }
        int same_size = 0, first_size = 1;
        for (int f = 0; f < novf; f++)
            if ((int)(sorted[f] & 0xffffffu) == sz) { same_size++; if (f < e) first_size = 0; }
        if (first_size) z.PJ2 += (long long)same_size * same_size;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32)
        if (pg[i]) zs_level(z, tb, i + 1, pg[i]);
    zs_reduce(z, tb, lane);
    if (lane == 0) {
        const double N = z.N ? (double)z.N : 1.0, rN = radb_div(1.0, N);
        const double Np = z.J1 ? (double)z.J1 : 1.0;
        o[0] = (double)z.G2 * rN;
        o[1] = (double)z.G2 * rN * rN;
        o[2] = (double)(z.N * z.GI2 - z.GI * z.GI) * rN * rN;
        o[3] = (double)z.GI2 * rN;
        o[4] = (double)z.J2 * rN;
        o[5] = z.lh * rN;
        o[6] = z.ll * rN;
        o[7] = z.lgl * rN;
        o[8] = (double)z.PJ2 * rN;
        o[9] = (double)z.PJ2 * rN * rN;
        o[10] = z.small * rN;
        o[11] = z.sh * rN;
        o[12] = z.sl * rN;
        o[13] = z.N ? radb_log2(N) - z.e1 * rN - RADB_EPS_LN2 * (double)z.nnz : 0.0;
        o[14] = radb_div(N, Np);
        o[15] = (double)(z.N * z.J2 - z.J1 * z.J1) * rN * rN;
    }
}

// A.9 / gldm.py.  D[n][nd] counters, column = dependence count (size j = col + 1).
__device__ void gldm_task(const RadbTabs& tb, const int* D, int n, int nd, double* o, int lane)
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
            if (j < nd) {
                const int c = D[i * nd + j];
                if (c) {
                    g += c;
                    colsum[j] += c;
                    zs_cell(z, tb, i + 1, j + 1, c);
                }
            }
        }
        if (g) zs_level(z, tb, i + 1, g);
    }
#pragma unroll
    for (int j = 0; j < 9; j++) {
        int cs = warp_sum_i(colsum[j]);
        if (lane == 0) z.PJ2 += (long long)cs * cs;
    }
    zs_reduce(z, tb, lane);
    if (lane == 0) {
        const double N = z.N ? (double)z.N : 1.0, rN = radb_div(1.0, N);
        o[0] = z.N ? radb_log2(N) - z.e1 * rN - RADB_EPS_LN2 * (double)z.nnz : 0.0;
        o[1] = (double)z.PJ2 * rN;
        o[2] = (double)z.PJ2 * rN * rN;
        o[3] = (double)(z.N * z.J2 - z.J1 * z.J1) * rN * rN;
        o[4] = (double)z.G2 * rN;
        o[5] = (double)(z.N * z.GI2 - z.GI * z.GI) * rN * rN;
        o[6] = (double)z.GI2 * rN;
        o[7] = (double)z.J2 * rN;
        o[8] = z.lh * rN;
        o[9] = z.ll * rN;
        o[10] = z.lgl * rN;
        o[11] = z.small * rN;
        o[12] = z.sh * rN;
        o[13] = z.sl * rN;
    }
}

// A.9 / ngtdm.py.  C[n][nb] = #voxels of level i with (col+1) valid neighbours,
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
            s += radb_div((double)S[i * nb + c], (double)(c + 1));
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
    const double rNvp = radb_div(1.0, Nvp);
    for (int i = lane; i < n; i += 32) pi[i] = radb_div(pi[i], Nvp);  // true division, as upstream: see absd below
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
            // upstream tests this sum against exactly 0 (Busyness): the two products must be rounded
            // separately -- an fma would leave a spurious 1e-17 when i*p_i == j*p_j
            absd += fabs(__dmul_rn(di, p_i) - __dmul_rn(dj, p_j));
            cplx += radb_div(fabs(dd) * (p_i * s_i + p_j * si[j]), p_i + p_j);
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
        o[0] = (absd != 0.0) ? radb_div(sum_ps, absd) : 0.0;
        o[1] = (sum_ps != 0.0) ? radb_div(1.0, sum_ps) : 1e6;
        o[2] = cplx * rNvp;
        o[3] = (div != 0.0) ? radb_div(contr * sum_s * rNvp, div) : 0.0;
        o[4] = (sum_s != 0.0) ? radb_div(stren, sum_s) : 0.0;
    }
}
