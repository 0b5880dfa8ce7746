// Thread-level ("lane") feature reductions: one THREAD per (patch, angle) instead of one warp.
// Included by radb_kernels.cuh.
//
// Why: at the gray-level counts the reference's settings produce (binWidth 25 -> Ng <= 11, binWidth 10
// -> Ng <= 26 for uint8 pixels) a GLCM / GLRLM angle is a few hundred cells.  Spread over a warp that is
// 4 cells per lane followed by a dozen warp reductions, and the MCC eigenproblem (Householder + Sturm) is
// a chain of ~10 dependent steps each ending in a warp reduction: ncu showed 37 k warp instructions per
// patch in the warp-per-angle kernel, half of them reduction / synchronisation overhead and most at
// 16-19 active lanes.  A thread that walks its own matrix serially needs no reduction at all; 32
// independent matrices per warp keep every lane busy and the per-patch instruction count drops ~20x.
//
// Storage: every thread owns `l_doubles` fp64 slots of shared memory, slot-major (slot k of thread t at
// smem[k * RADB_LSTRIDE + t]: consecutive lanes -> consecutive banks, conflict-free).  Integer scratch is
// packed two-per-slot inside the thread's own slots (never aliases another thread's data, so no CTA
// barriers are needed).  Layout for a patch with n gray levels (T = n(n+1)/2):
//   M [0, T)        packed lower triangle: GLCM counts as doubles, later A = Dx^-1/2 P Dx^-1/2 (compacted)
//   D [T, T+n)      ints px | psub during the count pass; later 1/sqrt(px); later Householder d / v
//   E [T+n, T+2n)   ints padd during the count pass; later N/px; later Householder e^2 / w
// The GLRLM task runs first and keeps its column sums (ints) at the start of the region.
// Formulas: pyradiomics glcm.py / glrlm.py as restated in SURVEY.md A.6 / A.7 (same closed forms as the
// warp-level tasks in radb_features.cuh); symmetric GLCMs only (symmetricalGLCM: True, params.yml:119) --
// the asymmetric case stays on the warp kernel.
#pragma once

struct LaneMem {
    double* d;  // this thread's slot 0 (fp64 view)
    int* i;     // this thread's slot 0 (int view)
};
#define LMD(m, k) (m).d[(k) * RADB_LSTRIDE]
#define LMI(m, k) (m).i[((k) >> 1) * (2 * RADB_LSTRIDE) + ((k) & 1)]

// ------------------------------------------------------------------ GLRLM, one thread
// R = run counters [n][nr] of one angle (packed u16, or u32 in wide mode); returns 0 for an empty angle.
__device__ int glrlm_lane(const RadbTabs& tb, const unsigned* R, int wide, int n, int nr, int maxlen, LaneMem lm,
                          double* o)
{
    for (int j = 0; j < maxlen; j++) LMI(lm, j) = 0;  // column sums p_r(j)
    long long sN = 0, sGI = 0, sGI2 = 0, sG2 = 0;
    double lgl = 0, e1 = 0, srl = 0, srh = 0, lrl = 0, lrh = 0;
    int nnz = 0;
    for (int i = 0; i < n; i++) {
        int rs = 0;
        long long bj2 = 0;  // sum_j c * j^2 (exact)
        double arj = 0;     // sum_j c / j^2
        const int c0 = i * nr;
        auto cell = [&](int j, int c) {
            rs += c;
            LMI(lm, j) += c;
            e1 += tab_clog(tb, c);
            nnz++;
            arj += (double)c * tab_inv2(tb, j + 1);
            bj2 += (long long)c * (j + 1) * (j + 1);
        };
        if (!wide && !(nr & 1)) {  // packed u16 pairs, rows start on a word: four independent word loads per step
            const unsigned* Rw = R + (c0 >> 1);
            const int nw = (maxlen + 1) >> 1;
            for (int w0 = 0; w0 < nw; w0 += 4) {
                unsigned ww[4];
#pragma unroll
                for (int u = 0; u < 4; u++) ww[u] = (w0 + u < nw) ? RADB_LDG(Rw + w0 + u) : 0u;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (!ww[u]) continue;
                    const int lo = (int)(ww[u] & 0xffffu), hi = (int)(ww[u] >> 16), j = 2 * (w0 + u);
                    if (lo) cell(j, lo);
                    if (hi) cell(j + 1, hi);
                }
            }
        } else {
            for (int j = 0; j < maxlen; j++) {
                const int c = get_run(R, c0 + j, wide);
                if (c) cell(j, c);
            }
        }
        if (rs) {
            const double i2 = (double)(i + 1) * (double)(i + 1), ri2 = tab_inv2(tb, i + 1);
            sN += rs;
            sGI += (long long)rs * (i + 1);
            sGI2 += (long long)rs * (i + 1) * (i + 1);
            sG2 += (long long)rs * rs;
            lgl += (double)rs * ri2;
            srl += ri2 * arj;
            srh += i2 * arj;
            lrl += ri2 * (double)bj2;
            lrh += i2 * (double)bj2;
        }
    }
    if (sN == 0) return 0;
    long long sRJ = 0, sRJ2 = 0, sR2 = 0;
    double sre = 0;
    for (int j = 0; j < maxlen; j++) {
        const int c = LMI(lm, j);
        if (!c) continue;
        sRJ += (long long)c * (j + 1);
        sRJ2 += (long long)c * (j + 1) * (j + 1);
        sR2 += (long long)c * c;
        sre += (double)c * tab_inv2(tb, j + 1);
    }
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
    return 1;
}

// ------------------------------------------------------------------ MCC, one thread
// #eigenvalues of the tridiagonal (d, e2) below x (same recurrence and sign convention as sturm_count, slot-major
// operands) and the last term of the sequence, p_m(x) = det(T - x I), in `pm`.  No rescaling: this kernel only takes
// matrices of at most 14 levels (use_lane == 1), T = Q^T A Q has its spectrum, diagonal and x inside [-1, 1] and
// e^2 <= 1, so |p_i| <= 3^i and a product of m <= 14 distances to eigenvalues cannot underflow either.  (The rescaling
// was a quarter of this kernel's Sturm instructions and sat on the recurrence's dependency chain.)
__device__ __forceinline__ int sturm_lane(LaneMem lm, int oD, int oE, int m, double x, double& pm)
{
    double p0 = 1.0, p1 = LMD(lm, oD) - x;
    bool neg = p1 < 0.0;
    int cnt = neg ? 1 : 0;
#pragma unroll 1
    for (int i = 1; i < m; i++) {
        const double pn = (LMD(lm, oD + i) - x) * p1 - LMD(lm, oE + i - 1) * p0;
        const bool nn = (pn < 0.0) || (pn == 0.0 && neg);
        cnt += (nn != neg) ? 1 : 0;
        neg = nn;
        p0 = p1;
        p1 = pn;
    }
    pm = p1;
    return cnt;
}
// k-th smallest eigenvalue; the spectrum of A = Dx^-1/2 P Dx^-1/2 lies in [-1, 1].  Bisection on the Sturm count until
// the bracket holds lambda_k alone (count(lo) == k, count(hi) == k + 1: p_m changes sign across it), then regula falsi
// (Illinois) on p_m inside the bracket -- every evaluation still moves a bracket end by its COUNT, so a wrong secant
// step costs an iteration, never the eigenvalue; every eighth step is a bisection.  ~12 evaluations instead of ~35.
__device__ double tridiag_kth_lane(LaneMem lm, int oD, int oE, int m, int k)
{
    double lo = -1.000001, hi = 1.000001, flo = 0.0, fhi = 0.0, prev = 2.0;
    int cl = 0, ch = m, side = 0;
    bool klo = false, khi = false;
#pragma unroll 1
    for (int it = 0; it < 44 && hi - lo > 1e-10; it++) {
        double mid = 0.5 * (lo + hi);
        const bool secant = cl == k && ch == k + 1 && klo && khi && ((flo < 0.0) != (fhi < 0.0)) && (it & 7) != 7;
        if (secant) {
            const double w = hi - lo, s = lo - flo * radb_div(w, fhi - flo);
            mid = fmin(fmax(s, lo + 1e-4 * w), hi - 1e-4 * w);
            if (fabs(mid - prev) <= 2e-11) return mid;  // superlinear: the error is far below the last step
            prev = mid;
        }
        double f;
        const int c = sturm_lane(lm, oD, oE, m, mid, f);
        if (c <= k) {  // count(x) <= k  <=>  x <= lambda_k
            lo = mid; cl = c; flo = f; klo = true;
            if (side < 0) fhi *= 0.5;
            side = -1;
        } else {
            hi = mid; ch = c; fhi = f; khi = true;
            if (side > 0) flo *= 0.5;
            side = 1;
        }
    }
    return 0.5 * (lo + hi);
}

// On entry: M = GLCM counts (packed lower triangle, n x n, as doubles), E[i] = N / px[i] (0: level absent).
// Returns the second largest |eigenvalue| of A (= sqrt of the second largest eigenvalue of Q, A.6).
__device__ double mcc_lane(LaneMem lm, int n, double rN)
{
    const int oD = n * (n + 1) / 2, oE = oD + n;
    int m = 0;
    for (int i = 0; i < n; i++) {
        const double r = LMD(lm, oE + i);
        LMD(lm, oD + i) = r > 0.0 ? radb_sqrt(r * rN) : 0.0;  // 1 / sqrt(px[i])
        m += r > 0.0;
    }
    if (m < 2) return 0.0;
    {   // compact the levels that occur and scale, in place (the write index never passes the read index)
        int w = 0;
        for (int i = 0; i < n; i++) {
            const double ri = LMD(lm, oD + i);
            if (ri == 0.0) continue;
            const int rb = i * (i + 1) / 2;
            for (int j = 0; j <= i; j++) {
                const double rj = LMD(lm, oD + j);
                if (rj == 0.0) continue;
                LMD(lm, w) = LMD(lm, rb + j) * ri * rj;
                w++;
            }
        }
    }
    // Householder tridiagonalisation of the m x m packed matrix; d -> D, e^2 -> E, v / w in their tails
#pragma unroll 1
    for (int k = 0; k < m - 2; k++) {
        const int ck = (k + 1) * (k + 2) / 2 + k;  // tri(k+1, k)
        double tail = 0;
        {
            int idx = ck + k + 2;  // tri(k+2, k)
            for (int r = k + 2; r < m; r++) {
                const double x = LMD(lm, idx);
                tail += x * x;
                LMD(lm, oD + r) = x;  // v[r]
                LMD(lm, oE + r) = 0.0;
                idx += r + 1;
            }
        }
        const double x0 = LMD(lm, ck), dk = LMD(lm, k * (k + 1) / 2 + k);
        if (tail == 0.0) {
            LMD(lm, oD + k) = dk;
            LMD(lm, oE + k) = x0 * x0;
            continue;
        }
        const double nrm = radb_sqrt(tail + x0 * x0);
        const double alpha = x0 > 0 ? -nrm : nrm;
        const double v0 = x0 - alpha;
        const double beta = radb_div(2.0, tail + v0 * v0);
        LMD(lm, oD + k + 1) = v0;
        LMD(lm, oE + k + 1) = 0.0;
        // q = M22 v in one sweep over the lower triangle; vMv = v^T q
        double vmv = 0;
        for (int r = k + 1; r < m; r++) {
            const int rb = r * (r + 1) / 2;
            const double vr = LMD(lm, oD + r);
            double acc = 0;
            for (int c = k + 1; c < r; c++) {
                const double mv = LMD(lm, rb + c);
                acc += mv * LMD(lm, oD + c);
                LMD(lm, oE + c) += mv * vr;
            }
            acc += LMD(lm, rb + r) * vr;
            LMD(lm, oE + r) += acc;
        }
        for (int r = k + 1; r < m; r++) vmv += LMD(lm, oE + r) * LMD(lm, oD + r);
        const double K = 0.5 * beta * beta * vmv;
        for (int r = k + 1; r < m; r++) LMD(lm, oE + r) = beta * LMD(lm, oE + r) - K * LMD(lm, oD + r);  // w'
        // M22 -= v w'^T + w' v^T
        for (int r = k + 1; r < m; r++) {
            const int rb = r * (r + 1) / 2;
            const double vr = LMD(lm, oD + r), wr = LMD(lm, oE + r);
            for (int c = k + 1; c <= r; c++) LMD(lm, rb + c) -= vr * LMD(lm, oE + c) + wr * LMD(lm, oD + c);
        }
        LMD(lm, oD + k) = dk;
        LMD(lm, oE + k) = alpha * alpha;
    }
    {
        const int a = (m - 2) * (m - 1) / 2 + (m - 2), b = (m - 1) * m / 2;
        LMD(lm, oD + m - 2) = LMD(lm, a);
        const double x = LMD(lm, b + m - 2);
        LMD(lm, oE + m - 2) = x * x;
        LMD(lm, oD + m - 1) = LMD(lm, b + m - 1);
    }
    const double t = fabs(tridiag_kth_lane(lm, oD, oE, m, m - 2));
    // second largest |lambda(A)|: lambda_min matters only if it lies below -|lambda_2|
    double pm;
    if (sturm_lane(lm, oD, oE, m, -t * (1.0 + 1e-9) - 1e-12, pm) == 0) return t;
    return fmax(t, fabs(tridiag_kth_lane(lm, oD, oE, m, 0)));
}

// ------------------------------------------------------------------ GLCM, one thread (symmetric matrices)
// P = final integer counts of one angle [n][n] (already symmetrised: P[i][j] == P[j][i], diagonal doubled).
// (Staging the matrices cooperatively through shared memory instead of reading them per thread was tried:
// the odd slot stride it needs costs one resident CTA per SM and the kernel got 25 % slower.)
// MCC = false: mid-size matrices whose packed fp64 copy would not leave enough resident threads -- nothing is
// stored (T = 0: the region is just D | E), the second pass re-reads P, and the eigenproblem is solved by
// radb_mcc_g8_kernel (o[19] is filled in by the caller from the record).
template <bool MCC>
__device__ int glcm_lane(const RadbTabs& tb, const int* P, int n, LaneMem lm, double* o)
{
    const int T = MCC ? n * (n + 1) / 2 : 0;
    const int iPX = 2 * T, iSUB = 2 * T + n, iADD = 2 * T + 2 * n;  // int views inside D | E
    for (int k = 0; k < 4 * n; k++) LMI(lm, iPX + k) = 0;
    long long sIJ = 0, sD2 = 0, sC2 = 0;
    int nnz = 0, maxc = 0;
    for (int i = 0; i < n; i++) {
        const int* row = P + i * n;
        const int rb = i * (i + 1) / 2;
        int rs = 0;
        for (int j0 = 0; j0 < i; j0 += 4) {  // cell (i, j) stands for (i, j) and (j, i); four loads in flight
            int cc[4];
#pragma unroll
            for (int u = 0; u < 4; u++) cc[u] = (j0 + u < i) ? RADB_LDG(row + j0 + u) : 0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = j0 + u, c = cc[u];
                if (MCC && j < i) LMD(lm, rb + j) = (double)c;
                if (c) {
                    rs += c;
                    LMI(lm, iPX + j) += c;
                    LMI(lm, iADD + i + j) += 2 * c;
                    LMI(lm, iSUB + i - j) += 2 * c;
                    sIJ += 2LL * c * (i + 1) * (j + 1);
                    sD2 += 2LL * c * (i - j) * (i - j);
                    sC2 += 2LL * c * c;
                    nnz += 2;
                    maxc = c > maxc ? c : maxc;
                }
            }
        }
        const int c = RADB_LDG(row + i);
        if (MCC) LMD(lm, rb + i) = (double)c;
        if (c) {
            rs += c;
            LMI(lm, iADD + 2 * i) += c;
            LMI(lm, iSUB) += c;
            sIJ += (long long)c * (i + 1) * (i + 1);
            sC2 += (long long)c * c;
            nnz++;
            maxc = c > maxc ? c : maxc;
        }
        LMI(lm, iPX + i) += rs;
    }
    long long sN = 0, sI = 0;
    for (int i = 0; i < n; i++) {
        const int a = LMI(lm, iPX + i);
        sN += a;
        sI += (long long)a * (i + 1);
    }
    if (sN == 0) return 0;
    const double N = (double)sN, rN = radb_div(1.0, N);
    const double ux = radb_div((double)sI, N);  // = uy; a true division: exactly 1 for a one-level matrix (sigma = 0)
    const double autoc = (double)sIJ * rN, contrast = (double)sD2 * rN, energy = (double)sC2 * rN * rN;
    const double maxp = (double)maxc * rN;
    const double log2N = tab_log2(tb, (int)sN);
    // marginal entropies (same table for c, px and N: exact cancellation for a one-level ROI)
    double hx0 = 0;
    int nx = 0;
    for (int i = 0; i < n; i++) {
        const int a = LMI(lm, iPX + i);
        if (a) { hx0 -= (double)a * rN * (tab_log2(tb, a) - log2N); nx++; }
    }
    const double hx = hx0 - RADB_EPS_LN2 * (double)nx;
    // |i-j| marginal
    double da = 0, de = 0, idv = 0, idm = 0, idmn = 0, idn = 0, iv = 0;
    const double rdn = radb_div(1.0, (double)n);
    for (int k = 0; k < n; k++) {
        const int c = LMI(lm, iSUB + k);
        if (!c) continue;
        const double q = (double)c * rN, dk = (double)k;
        da += dk * q;
        de -= q * (tab_log2(tb, c) - log2N) + RADB_EPS_LN2;
        idv += radb_div(q, 1.0 + dk);
        idm += radb_div(q, 1.0 + dk * dk);
        idmn += radb_div(q, 1.0 + (dk * dk) * rdn * rdn);
        idn += radb_div(q, 1.0 + dk * rdn);
        if (k > 0) iv += q * tab_inv2(tb, k);
    }
    double dvar = 0;
    for (int k = 0; k < n; k++) {
        const int c = LMI(lm, iSUB + k);
        if (c) dvar += (double)c * rN * ((double)k - da) * ((double)k - da);
    }
    // i+j marginal (index k <-> i+j = k+2)
    double sa = 0, se = 0;
    for (int k = 0; k < 2 * n - 1; k++) {
        const int c = LMI(lm, iADD + k);
        if (!c) continue;
        const double q = (double)c * rN;
        sa += (double)(k + 2) * q;
        se -= q * (tab_log2(tb, c) - log2N) + RADB_EPS_LN2;
    }
    // E[i] = N / px[i] (padd is dead; px lives in D and is consumed here)
    const int oE = T + n;
    for (int i = 0; i < n; i++) {
        const int a = LMI(lm, iPX + i);
        LMD(lm, oE + i) = a ? radb_div(N, (double)a) : 0.0;
    }
    // second pass over the stored counts: cluster moments, correlation, joint entropy
    double ct = 0, cs = 0, cp = 0, ssq = 0, corm = 0, h1 = 0, sclog = 0;
    for (int i = 0; i < n; i++) {
        const int rb = i * (i + 1) / 2;
        const int* row = P + i * n;
        const double di = (double)(i + 1) - ux, rpxi = LMD(lm, oE + i);
        for (int j0 = 0; j0 < i; j0 += 4) {
            double dd[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                dd[u] = (j0 + u < i) ? (MCC ? LMD(lm, rb + j0 + u) : (double)RADB_LDG(row + j0 + u)) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const double dc = dd[u];
                if (dc == 0.0) continue;
                const int j = j0 + u;
                const double pij = dc * rN, dj = (double)(j + 1) - ux, s = di + dj, s2 = s * s;
                const double p2 = pij + pij;
                ct += p2 * s2;
                cs += p2 * s2 * s;
                cp += p2 * s2 * s2;
                ssq += pij * (di * di + dj * dj);
                corm += p2 * di * dj;
                h1 += 2.0 * dc * rpxi * LMD(lm, oE + j);
                sclog += 2.0 * dc * (tab_log2(tb, (int)dc) - log2N);
            }
        }
        const double dc = MCC ? LMD(lm, rb + i) : (double)RADB_LDG(row + i);
        if (dc != 0.0) {
            const double pij = dc * rN, s = di + di, s2 = s * s;
            ct += pij * s2;
            cs += pij * s2 * s;
            cp += pij * s2 * s2;
            ssq += pij * di * di;
            corm += pij * di * di;
            h1 += dc * rpxi * rpxi;
            sclog += dc * (tab_log2(tb, (int)dc) - log2N);
        }
    }
    const double h1corr = h1 * rN;  // sum p / (px * py)
    const double hxy = -sclog * rN - RADB_EPS_LN2 * (double)nnz;
    const double hxy1 = hx0 + hx0 - RADB_EPS_LN2 * h1corr;
    const double hxy2 = hx0 + hx0 - RADB_EPS_LN2 * (double)nx * (double)nx;
    const double mcc = MCC ? mcc_lane(lm, n, rN) : 0.0;
    double im2 = 1.0 - exp(-2.0 * (hxy2 - hxy));
    im2 = im2 < 0.0 ? 0.0 : im2;
    o[0] = autoc;
    o[1] = cp;
    o[2] = cs;
    o[3] = ct;
    o[4] = contrast;
    {
        const double sg = radb_sqrt(ssq);
        o[5] = (sg * sg == 0.0) ? 1.0 : radb_div(corm, sg * sg + RADB_EPS);
    }
    o[6] = da;
    o[7] = de;
    o[8] = dvar;
    o[9] = idv;
    o[10] = idm;
    o[11] = idmn;
    o[12] = idn;
    o[13] = (hx != 0.0) ? radb_div(hxy - hxy1, hx) : 0.0;
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
    return 1;
}

// ------------------------------------------------------------------ the kernel body
// Thread g of the grid <-> (patch g / NAP, angle g % NAP), NAP = n_angles rounded up to a power of two,
// so the angles of a patch are adjacent lanes of one warp and the nanmean over the non-empty angles
// (A.6: upstream deletes empty angles before numpy.nanmean) is two xor shuffles.
template <int NF>
__device__ __forceinline__ void lane_mean(double (&f)[NF], int valid, int nap)
{
    int k = valid;
    for (int m = 1; m < nap; m <<= 1) k += __shfl_xor_sync(FULLMASK, k, m);
    const double rk = k == 4 ? 0.25 : k == 3 ? (1.0 / 3.0) : k == 2 ? 0.5 : 1.0;
#pragma unroll
    for (int i = 0; i < NF; i++) {
        double s = valid ? f[i] : 0.0;
        for (int m = 1; m < nap; m <<= 1) s += __shfl_xor_sync(FULLMASK, s, m);
        f[i] = k ? s * rk : nan_f64();
    }
}
template <int NF>
__device__ __forceinline__ void lane_store(const double (&f)[NF], int nap, int a, double* out)
{
#pragma unroll
    for (int i = 0; i < NF; i++)
        if ((i & (nap - 1)) == a) out[i] = f[i];  // the lanes of a patch share the stores
}

__device__ void radb_angle_lane_cta(const RadbParams& p, long long cta, unsigned char* smem)
{
    const int t = threadIdx.x;
    const int NA = p.n_angles, NAP = p.l_nap;
    const long long g = cta * RADB_NTL + t;
    const long long patch = g / NAP;
    const int a = (int)(g - patch * NAP);
    const long long row = patch < p.B ? radb_row(p, patch) : 0;
    const bool patch_ok = patch < p.B && p.status[row] == 0;  // status != 0: NaN row written by the build kernel
    const bool live = patch_ok && a < NA;
    LaneMem lm;
    lm.d = (double*)smem + t;
    lm.i = (int*)smem + 2 * t;
    RadbTabs tb;
    tb.inv2 = p.g_inv2;
    tb.ninv = p.ninv;
    tb.tlog = p.g_tlog;
    tb.red = (double*)0;
    const unsigned char* rec = p.ws + (patch_ok ? patch : 0) * (long long)p.rec_bytes;
    const int* misc = (const int*)(rec + (p.o_misc - p.o_rec));
    int ng = 0, nroi = 0;
    if (patch_ok) { ng = misc[8]; nroi = misc[9]; }
    double* out = p.out + row * (long long)p.F;
    if (p.off_glrlm >= 0) {
        double f[RADB_GLRLM_NF];
#pragma unroll
        for (int i = 0; i < RADB_GLRLM_NF; i++) f[i] = 0.0;
        int ok = 0;
        if (live)
            ok = glrlm_lane(tb, (const unsigned*)(rec + (p.o_glrlm - p.o_rec) + a * p.glrlm_stride), p.wide, ng, p.nrp,
                            misc[10 + a], lm, f);
        lane_mean<RADB_GLRLM_NF>(f, ok, NAP);
        if (patch_ok) lane_store<RADB_GLRLM_NF>(f, NAP, a, out + p.off_glrlm);
    }
    if (p.off_glcm >= 0) {
        double f[RADB_GLCM_NF];
#pragma unroll
        for (int i = 0; i < RADB_GLCM_NF; i++) f[i] = 0.0;
        int ok = 0;
        if (live) {
            const int* P = (const int*)(rec + (p.o_glcm - p.o_rec)) + a * ng * ng;
            if (p.use_lane == 1) {
                ok = glcm_lane<true>(tb, P, ng, lm, f);
            } else {
                ok = glcm_lane<false>(tb, P, ng, lm, f);
                f[19] = ((const double*)(misc + RADB_REC_MCC_INT))[a];  // written by radb_mcc_g8_kernel
            }
        }
        lane_mean<RADB_GLCM_NF>(f, ok, NAP);
        if (nroi < 2) f[19] = 1.0;  // MCC of a one-level ROI (glcm.py: "flat region" special case)
        if (patch_ok) lane_store<RADB_GLCM_NF>(f, NAP, a, out + p.off_glcm);
    }
}

// ==================================================================== misc classes, one thread per (patch, class)
// CTA = 128 threads = 128 patches, thread <-> patch: a warp reduces first-order, GLDM, NGTDM and GLSZM one after
// the other for its 32 patches, so it runs one code path at a time with all lanes busy.  Per-thread fp64
// scratch (2 * max_ng + 16 slots) is slot-major per warp (slot k of lane l at base[k * 32 + l]).
// Same closed forms as the warp-level tasks in radb_features.cuh.
#define MLD(b, k) (b)[(k) * 32]
#define MLU(b, k) (((unsigned*)&(b)[((k) >> 1) * 32])[(k) & 1])  // u32 view: two per fp64 slot of the same thread
#define RADB_LANE_MAX_OVF 48  // GLSZM patches with more overflow zones than this go to the warp kernel

// A.8 GLSZM: dense counters Z[n][s0] + overflow list ((level - 1) << 24 | size) of the zones larger than s0
__device__ void glszm_lane(const RadbParams& p, const RadbTabs& tb, const int* Z, const unsigned* ovf, int novf, int n,
                           double* scr, double* o)
{
    const int s0 = p.s0;
    ZoneSums z;
    zs_init(z);
    int col[16];  // column sums when s0 == 16 (narrow mode); the wide layout (s0 = 64) takes a second pass
#pragma unroll
    for (int j = 0; j < 16; j++) col[j] = 0;
    for (int i = 0; i < n; i++) {
        int g = 0;
        if (s0 == 16) {
            const int4* Zr = (const int4*)(Z + i * 16);  // 64-byte rows of a 16-byte aligned matrix
            int4 q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) q[u] = RADB_LDG(Zr + u);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int c4[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int c = c4[k];
                    if (!c) continue;
                    g += c;
                    col[4 * u + k] += c;
                    zs_cell(z, tb, i + 1, 4 * u + k + 1, c);
                }
            }
        } else {
            for (int j = 0; j < s0; j++) {
                const int c = RADB_LDG(Z + i * s0 + j);
                if (!c) continue;
                g += c;
                zs_cell(z, tb, i + 1, j + 1, c);
            }
        }
        for (int e = 0; e < novf; e++) g += ((int)(RADB_LDG(ovf + e) >> 24) == i);
        if (g) zs_level(z, tb, i + 1, g);
    }
    if (s0 == 16) {
#pragma unroll
        for (int j = 0; j < 16; j++) z.PJ2 += (long long)col[j] * col[j];
    } else {
        for (int j = 0; j < s0; j++) {
            int cs = 0;
            for (int i = 0; i < n; i++) cs += RADB_LDG(Z + i * s0 + j);
            z.PJ2 += (long long)cs * cs;
        }
    }
    // overflow zones: the list was appended in atomic order, so it is first insertion-sorted by (size, level)
    // into this thread's scratch -- every sum below then runs in an order that depends on the data only
    // (bit-reproducible rows), equal cells are adjacent and so are equal sizes
    for (int e = 0; e < novf; e++) {
        const unsigned k0 = RADB_LDG(ovf + e), key = ((k0 & 0xffffffu) << 8) | (k0 >> 24);
        int j = e;
        while (j > 0 && MLU(scr, j - 1) > key) { MLU(scr, j) = MLU(scr, j - 1); j--; }
        MLU(scr, j) = key;
    }
    for (int e = 0; e < novf;) {
        const unsigned key = MLU(scr, e);
        const int sz = (int)(key >> 8);
        int same_size = 0;
        while (e < novf && (int)(MLU(scr, e) >> 8) == sz) {  // all zones of this size
            const unsigned k1 = MLU(scr, e);
            int cnt = 0;
            while (e < novf && MLU(scr, e) == k1) { cnt++; e++; }
            zs_cell(z, tb, (int)(k1 & 0xffu) + 1, sz, cnt);
            same_size += cnt;
        }
        z.PJ2 += (long long)same_size * same_size;
    }
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

// A.9 GLDM: D[n][nd] counters, column = dependence count
__device__ void gldm_lane(const RadbTabs& tb, const int* D, int n, int nd, double* o)
{
    ZoneSums z;
    zs_init(z);
    int col[9];  // nd <= 2 * RADB_MAX_ANGLES + 1
#pragma unroll
    for (int j = 0; j < 9; j++) col[j] = 0;
    for (int i = 0; i < n; i++) {
        int g = 0, cc[9];
#pragma unroll
        for (int j = 0; j < 9; j++) cc[j] = j < nd ? RADB_LDG(D + i * nd + j) : 0;
#pragma unroll
        for (int j = 0; j < 9; j++) {
            const int c = cc[j];
            if (!c) continue;
            g += c;
            col[j] += c;
            zs_cell(z, tb, i + 1, j + 1, c);
        }
        if (g) zs_level(z, tb, i + 1, g);
    }
#pragma unroll
    for (int j = 0; j < 9; j++) z.PJ2 += (long long)col[j] * col[j];
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

// A.9 NGTDM: C[n][nb] voxel counts per neighbour count, S[n][nb] integer numerators (see ngtdm_task)
__device__ void ngtdm_lane(const int* C, const int* S, int n, int nb, double* scr, double* o, int* dbg_n, double* dbg_s)
{
    long long nvp = 0;
    for (int i = 0; i < n; i++) {
        int ni = 0;
        double s = 0;
        int cc[8], ss[8];  // nb <= 2 * RADB_MAX_ANGLES
#pragma unroll
        for (int c = 0; c < 8; c++) {
            cc[c] = c < nb ? RADB_LDG(C + i * nb + c) : 0;
            ss[c] = c < nb ? RADB_LDG(S + i * nb + c) : 0;
        }
#pragma unroll
        for (int c = 0; c < 8; c++) {
            if (!cc[c]) continue;
            ni += cc[c];
            s += radb_div((double)ss[c], (double)(c + 1));
        }
        MLD(scr, i) = (double)ni;
        MLD(scr, n + i) = s;
        nvp += ni;
        if (dbg_n) { dbg_n[i] = ni; dbg_s[i] = s; }
    }
    if (nvp == 0) {
        for (int f = 0; f < 5; f++) o[f] = nan_f64();
        return;
    }
    const double Nvp = (double)nvp, rNvp = radb_div(1.0, Nvp);
    for (int i = 0; i < n; i++) {
        const double c = MLD(scr, i);
        if (c != 0.0) MLD(scr, i) = radb_div(c, Nvp);  // true division, as upstream (Busyness tests a sum against 0)
    }
    double sum_ps = 0, sum_s = 0, absd = 0, cplx = 0, contr = 0, stren = 0;
    int ngp = 0;
    for (int i = 0; i < n; i++) {
        const double p_i = MLD(scr, i);
        if (p_i == 0.0) continue;
        ngp++;
        const double s_i = MLD(scr, n + i), di = (double)(i + 1), ip = __dmul_rn(di, p_i);
        sum_ps += p_i * s_i;
        sum_s += s_i;
        for (int j = 0; j < i; j++) {  // the (i, j) and (j, i) terms are equal
            const double p_j = MLD(scr, j);
            if (p_j == 0.0) continue;
            const double dj = (double)(j + 1), dd = di - dj, d2 = dd * dd;
            absd += 2.0 * fabs(ip - __dmul_rn(dj, p_j));
            cplx += 2.0 * radb_div(fabs(dd) * (p_i * s_i + p_j * MLD(scr, n + j)), p_i + p_j);
            contr += 2.0 * p_i * p_j * d2;
            stren += 2.0 * (p_i + p_j) * d2;
        }
    }
    const double div = (double)ngp * (double)(ngp - 1);
    o[0] = (absd != 0.0) ? radb_div(sum_ps, absd) : 0.0;
    o[1] = (sum_ps != 0.0) ? radb_div(1.0, sum_ps) : 1e6;
    o[2] = cplx * rNvp;
    o[3] = (div != 0.0) ? radb_div(contr * sum_s * rNvp, div) : 0.0;
    o[4] = (sum_s != 0.0) ? radb_div(stren, sum_s) : 0.0;
}

// A.5 first-order for uint8 pixels from the 256-bin raw histogram + the level histogram
__device__ void fo_lane_u8(const RadbParams& p, const int* hist, const int* lhist, int ng, int N, double* scr, double* o)
{
    const double dN = (double)N, rN = radb_div(1.0, dN), shift = p.shift;
    // ranks behind the 10/25/50/75/90 percentiles (numpy 'linear'): lo/hi order statistics, ascending
    for (int q = 0; q < 5; q++) {
        const double qq = q == 0 ? 0.1 : q == 1 ? 0.25 : q == 2 ? 0.5 : q == 3 ? 0.75 : 0.9;
        const double pos = qq * (dN - 1.0), fl = floor(pos);
        int lo = (int)fl;
        lo = lo > N - 1 ? N - 1 : lo;
        const int hi = lo + 1 > N - 1 ? N - 1 : lo + 1;
        MLD(scr, 10 + q) = pos - fl;
        MLD(scr, 2 * q) = (double)lo;      // rank now, order statistic after the pass
        MLD(scr, 2 * q + 1) = (double)hi;
    }
    long long s1 = 0;
    int vmin = 256, vmax = -1, cum = 0;
    // the lo ranks (slots 0, 2, ..) and the hi ranks (slots 1, 3, ..) are each non-decreasing: one cursor per sequence
    int tl = 0, th = 1;
    int nextl = (int)MLD(scr, 0), nexth = (int)MLD(scr, 1);
    for (int v0 = 0; v0 < 256; v0 += 4) {
      const int4 h4 = RADB_LDG((const int4*)(hist + v0));  // 16-byte aligned histogram, four bins per load
      const int hq[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int v = v0 + u, hk = hq[u];
        if (!hk) continue;
        s1 += (long long)hk * v;
        vmin = v < vmin ? v : vmin;
        vmax = v;
        cum += hk;
        while (tl < 10 && nextl < cum) {
            MLD(scr, tl) = (double)v;
            tl += 2;
            nextl = tl < 10 ? (int)MLD(scr, tl) : 0;
        }
        while (th < 10 && nexth < cum) {
            MLD(scr, th) = (double)v;
            th += 2;
            nexth = th < 10 ? (int)MLD(scr, th) : 0;
        }
      }
    }
    const double mean = (double)s1 * rN;
    double pc[5];
#pragma unroll
    for (int q = 0; q < 5; q++) pc[q] = MLD(scr, 2 * q) + (MLD(scr, 2 * q + 1) - MLD(scr, 2 * q)) * MLD(scr, 10 + q);
    const double p10 = pc[0], p25 = pc[1], med = pc[2], p75 = pc[3], p90 = pc[4];
    double m2 = 0, m3 = 0, m4 = 0, mad = 0, en = 0, in_s1 = 0;
    int in_n = 0;
    for (int v0 = vmin & ~3; v0 <= vmax; v0 += 4) {  // bins outside [vmin, vmax] are empty
      const int4 h4 = RADB_LDG((const int4*)(hist + v0));
      const int hq[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int v = v0 + u, hi_ = hq[u];
        if (!hi_) continue;
        const double dv = (double)v, hk = (double)hi_, d = dv - mean, d2 = d * d;
        m2 += hk * d2;
        m3 += hk * d2 * d;
        m4 += hk * d2 * d2;
        mad += hk * fabs(d);
        en += hk * (dv + shift) * (dv + shift);
        if (dv >= p10 && dv <= p90) { in_n += hi_; in_s1 += hk * dv; }
      }
    }
    m2 *= rN; m3 *= rN; m4 *= rN; mad *= rN;
    const double rin = radb_div(1.0, (double)in_n), in_mean = in_s1 * rin;
    double rmad = 0;
    for (int v0 = vmin & ~3; v0 <= vmax; v0 += 4) {
      const int4 h4 = RADB_LDG((const int4*)(hist + v0));
      const int hq[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const double dv = (double)(v0 + u);
        if (hq[u] && dv >= p10 && dv <= p90) rmad += (double)hq[u] * fabs(dv - in_mean);
      }
    }
    double ent = 0, uni = 0;
    for (int i = 0; i < ng; i++) {
        const int c = RADB_LDG(lhist + i);
        if (!c) continue;
        const double pi = (double)c * rN;
        ent -= pi * radb_log2(pi + RADB_EPS);
        uni += pi * pi;
    }
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
    o[12] = rmad * rin;
    o[13] = radb_sqrt(en * rN);
    o[14] = (m2 == 0.0) ? 0.0 : radb_div(m3, m2 * radb_sqrt(m2));
    o[15] = en;  // TotalEnergy: pixel spacing is (1, 1)
    o[16] = uni;
    o[17] = m2;
}

__device__ void radb_misc_lane_cta(const RadbParams& p, long long cta, unsigned char* smem)
{
    // thread <-> patch; a warp runs the four classes one after the other for its 32 patches (one code path per
    // warp at any time, and no warp waits for a slower class of another warp)
    const int tid = threadIdx.x;
    const long long patch = cta * RADB_NT + tid;
    if (patch >= p.B) return;  // no collectives below: early exit is safe
    const long long row = radb_row(p, patch);
    if (p.status[row] != 0) return;
    const unsigned char* rec = p.ws + patch * (long long)p.rec_bytes;
    const int* misc = (const int*)(rec + (p.o_misc - p.o_rec));
    const int ng = misc[8], NB = 2 * p.n_angles;
    double* out = p.out + row * (long long)p.F;
    double* scr = (double*)smem + (tid >> 5) * (p.ml_doubles * 32) + (tid & 31);
    RadbTabs tb;
    tb.inv2 = p.g_inv2;
    tb.ninv = p.ninv;
    tb.tlog = p.g_tlog;
    tb.red = (double*)0;
    if (p.off_fo >= 0 && p.pix_bytes == 1)  // non-uint8: done by the build kernel
        fo_lane_u8(p, (const int*)(rec + (p.o_hist - p.o_rec)), (const int*)(rec + (p.o_lhist - p.o_rec)), ng, misc[0], scr,
                   out + p.off_fo);
    if (p.off_gldm >= 0) gldm_lane(tb, (const int*)(rec + (p.o_gldm - p.o_rec)), ng, NB + 1, out + p.off_gldm);
    if (p.off_ngtdm >= 0 || p.dbg_ngn) {
        double dummy[5];
        ngtdm_lane((const int*)(rec + (p.o_ngc - p.o_rec)), (const int*)(rec + (p.o_ngn - p.o_rec)), ng, NB, scr,
                   p.off_ngtdm >= 0 ? out + p.off_ngtdm : dummy, p.dbg_ngn ? p.dbg_ngn + patch * p.max_ng : (int*)0,
                   p.dbg_ngs ? p.dbg_ngs + patch * p.max_ng : (double*)0);
    }
    const int novf = misc[5];
    if (p.off_glszm >= 0 && novf <= RADB_LANE_MAX_OVF)
        glszm_lane(p, tb, (const int*)(rec + (p.o_szm - p.o_rec)), (const unsigned*)(rec + (p.o_ovf - p.o_rec)), novf, ng, scr,
                   out + p.off_glszm);
}

// ==================================================================== MCC kernel: 8 lanes per (patch, angle)
// The aligned 8-lane group g of a warp owns one (patch, angle) task: row sums of its matrix (p_x), then mcc_task_g8.  The
// four values go to the record header (RADB_REC_MCC_INT) for the thread-per-angle kernel's nanmean.
__device__ void radb_mcc_g8_cta(const RadbParams& p, long long cta, unsigned char* smem)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (p.off_glcm < 0) return;
    const int g = lane >> 3, gl = lane & 7, NA = p.n_angles;
    // 1, 2 or 4 angles: the four groups of a warp take consecutive (patch, angle) tasks, so a one-angle extractor (the
    // literal force2D reading of /root/reference/params.yml:100) solves four patches per warp instead of leaving 24
    // lanes idle; 3 angles: one patch per warp, group <-> angle
    const bool packed = NA == 1 || NA == 2 || NA == 4;
    const long long wi = cta * (RADB_NTM / 32) + warp, t = wi * 4 + g;
    const long long patch = packed ? t / NA : wi;
    const int ang = packed ? (int)(t - patch * NA) : g;
    if (patch >= p.B || ang >= NA) return;  // every collective below runs on the group's own mask
    if (p.status[radb_row(p, patch)] != 0) return;
    const unsigned gm = 0xffu << (8 * g);
    unsigned char* rec = p.ws + patch * (long long)p.rec_bytes;
    int* misc = (int*)(rec + (p.o_misc - p.o_rec));
    const int n = misc[8];
    const int* P = (const int*)(rec + (p.o_glcm - p.o_rec)) + ang * n * n;
    unsigned char* ws = smem + (warp * 4 + g) * p.g8_group_bytes;
    int* px = (int*)(ws + p.g8_px);
    for (int i = gl; i < n; i += 8) {
        int rs = 0;
        for (int j = 0; j < n; j++) rs += P[i * n + j];
        px[i] = rs;
    }
    __syncwarp(gm);
    const double mcc = mcc_task_g8(P, px, px, n, 1, (double*)(ws + p.g8_mcc), ws + p.g8_idx, gl, gm, 8 * g);
    if (gl == 0) ((double*)(misc + RADB_REC_MCC_INT))[ang] = mcc;
}
