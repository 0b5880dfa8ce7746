// MCC for many gray levels (more than 40: the binWidth sweep of BASELINE.json configs[4], 64 .. 256 levels).
// Included by radb_kernels.cuh.
//
// MCC = second largest |eigenvalue| of A = Dx^-1/2 P Dx^-1/2 (symmetric GLCM).  The dense
// Householder tridiagonalisation of the warp-per-angle kernel is O(m^3) per angle on ONE warp: 150 us per patch at
// 256 levels (round 1: 6 k patches/s).  Two facts make it cheap instead:
//   * the top eigenpair is known (lambda_1 = 1, v_1 = sqrt(px / N)), so the wanted value is the spectral radius of
//     the deflated operator B = A - v_1 v_1^T -- an EXTREME eigenvalue, which Lanczos finds in a few dozen
//     matrix-vector products;
//   * a 64x64 patch has at most 4032 voxel pairs per angle, so at 256 levels the matrix is >= 88 % zeros: the
//     products run over a CSR copy of the non-zeros built once per (patch, angle).
// One CTA of RADB_NTZ threads per (patch, angle): CSR build (warp per row, ballot compaction -> deterministic entry
// order), Lanczos without reorthogonalisation (ghost copies of converged Ritz values do not move the extreme ones),
// deflation re-applied in every step, extreme Ritz values of T_k by warp multisection (Sturm counts).  Stops when
// the spectral radius agrees to 1e-12 at two consecutive check points, at k = m (exact), or on breakdown.
// The per-angle results go to the record header (RADB_REC_MCC_INT) like those of radb_mcc_g8_kernel.
#pragma once

#define RADB_NTZ 128            // threads per CTA of the Lanczos kernel
#define RADB_LZ_KMAX 448        // Lanczos steps at most (alpha / beta^2 arrays)

// fixed-order CTA sums of K doubles per thread (4 warps): butterflies, one slot per warp, everyone adds the slots
template <int K>
__device__ __forceinline__ void lz_cta_sum(double (&v)[K], double* slots /*[K][4]*/, int tid)
{
    const int lane = tid & 31, warp = tid >> 5, NW = RADB_NTZ / 32;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const double t = warp_sum(v[k]);
        if (lane == 0) slots[k * NW + warp] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) {
        double t = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) t += slots[k * NW + w];
        v[k] = t;
    }
    __syncthreads();
}

// Largest and smallest eigenvalue of the symmetric tridiagonal (d, e2) of order m at once: warps 0-1 multisect for
// the largest (64 shifts per round), warps 2-3 for the smallest; 8 rounds shrink [-1, 1] to 2 / 65^8 < 1e-14.
// `sh` = 4 ints of shared scratch.  Every thread of the CTA must call it; all get both values.
__device__ void lz_extremes(const double* d, const double* e2, int m, int* sh, int tid, double& emax, double& emin)
{
    const int lane = tid & 31, warp = tid >> 5, half = warp >> 1, t64 = tid & 63;
    const int k = half == 0 ? m - 1 : 0;  // index (ascending) of the wanted eigenvalue
    double lo = -1.0000001, hi = 1.0000001;
#pragma unroll 1
    for (int it = 0; it < 8; it++) {
        const double w = (hi - lo) * (1.0 / 65.0);
        const double x = lo + w * (double)(t64 + 1);
        const int c = sturm_count(d, e2, m, x);
        const unsigned left = __ballot_sync(FULLMASK, c <= k);  // shifts left of (or at) the eigenvalue: a prefix of the 64
        if (lane == 0) sh[warp] = __popc(left);
        __syncthreads();
        const int nl = sh[2 * half] + sh[2 * half + 1];
        __syncthreads();
        const double nlo = lo + w * (double)nl;
        const double nhi = (nl == 64) ? hi : lo + w * (double)(nl + 1);
        lo = nlo;
        hi = nhi;
    }
    double* dsh = (double*)(sh + 4);
    if (t64 == 0) dsh[half] = 0.5 * (lo + hi);
    __syncthreads();
    emax = dsh[0];
    emin = dsh[1];
    __syncthreads();
}

// smem: rowptr int[n + 1] | cnt (aliases rowptr + 1) | ent u32[cap] | rs, v1, q0, q1, z double[n] each |
//       al, be2 double[KMAX] | slots double[3 * 4] | misc
__device__ void radb_mcc_lanczos_cta(const RadbParams& p, long long cta, unsigned char* smem)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = RADB_NTZ / 32;
    const int NA = p.n_angles;
    const long long patch = cta / NA;
    const int a = (int)(cta - patch * NA);
    if (patch >= p.B || p.off_glcm < 0) return;
    if (p.status[radb_row(p, patch)] != 0) return;
    unsigned char* rec = p.ws + patch * (long long)p.rec_bytes;
    int* misc = (int*)(rec + (p.o_misc - p.o_rec));
    const int n = misc[8];
    const int* P = (const int*)(rec + (p.o_glcm - p.o_rec)) + (long long)a * n * n;
    int* rowptr = (int*)(smem + p.z_rowptr);
    unsigned* ent = (unsigned*)(smem + p.z_ent);
    double* rs = (double*)(smem + p.z_vec);
    double* v1 = rs + p.max_ng;
    double* q0 = v1 + p.max_ng;
    double* q1 = q0 + p.max_ng;
    double* z = q1 + p.max_ng;
    double* al = (double*)(smem + p.z_tri);
    double* be2 = al + RADB_LZ_KMAX;
    double* slots = be2 + RADB_LZ_KMAX;
    int* ishare = (int*)(slots + 16);      // 4 ints + 2 doubles of lz_extremes, then the CSR totals
    double* dshare = (double*)(ishare + 4) + 2;
    double* out_mcc = (double*)(misc + RADB_REC_MCC_INT) + a;

    // ---- CSR build, pass 1: non-zeros per row and row sums (warp per row, coalesced reads)
    int* cnt = rowptr + 1;
    int* pxi = (int*)z;  // row sums as integers (z is free until the iteration starts)
    for (int r = warp; r < n; r += NW) {
        int c = 0, s = 0;
        for (int j = lane; j < n; j += 32) {
            const int v = RADB_LDG(P + r * n + j);
            c += v != 0;
            s += v;
        }
        c = warp_sum_i(c);
        s = warp_sum_i(s);
        if (lane == 0) { cnt[r] = c; pxi[r] = s; }
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the row counts -> rowptr; totals
        int carry = 0, m = 0;
        long long N = 0;
        for (int b0 = 0; b0 < n; b0 += 32) {
            const int i = b0 + lane;
            const int c = i < n ? cnt[i] : 0;
            const int ex = warp_excl_scan_i(c, lane);
            const int tot = warp_sum_i(c);
            m += __popc(__ballot_sync(FULLMASK, i < n && pxi[i] > 0));
            N += warp_sum_ll(i < n ? (long long)pxi[i] : 0LL);
            __syncwarp();
            if (i < n) rowptr[i + 1] = carry + ex + c;  // cnt[i] aliases rowptr[i + 1]: each lane overwrites its own slot
            carry += tot;
        }
        if (lane == 0) { rowptr[0] = 0; ishare[0] = m; dshare[0] = (double)N; }
    }
    __syncthreads();
    const int m = ishare[0];
    const double Ntot = dshare[0];
    if (m < 2 || rowptr[n] > p.z_cap) {
        // no voxel pair in this angle: NaN (the angle is left out of the mean, A.6); one level only: no second
        // eigenvalue, 0 (capacity: cannot happen, see radb_layout)
        if (tid == 0) *out_mcc = Ntot > 0.0 ? 0.0 : nan_f64();
        return;
    }
    // ---- pass 2: fill the entries in (row, column) order; scaling vectors
    for (int r = warp; r < n; r += NW) {
        int base = rowptr[r];
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            const int v = j < n ? RADB_LDG(P + r * n + j) : 0;
            const unsigned b = __ballot_sync(FULLMASK, v != 0);
            if (v) ent[base + __popc(b & ((1u << lane) - 1u))] = ((unsigned)j << 16) | (unsigned)v;
            base += __popc(b);
        }
    }
    for (int i = tid; i < n; i += RADB_NTZ) {
        const int s = pxi[i];
        const double dv = (double)s;
        rs[i] = s > 0 ? radb_div(1.0, radb_sqrt(dv)) : 0.0;
        v1[i] = s > 0 ? radb_sqrt(radb_div(dv, Ntot)) : 0.0;
    }
    __syncthreads();
    // ---- start vector: deterministic pseudo-random on the levels present, orthogonal to v1, unit length
    {
        double acc[2] = {0, 0};
        for (int i = tid; i < n; i += RADB_NTZ) {
            const unsigned h = (unsigned)(i + 1) * 2654435761u;
            const double x = rs[i] != 0.0 ? ((double)((h >> 8) & 0xffffu) * (1.0 / 65536.0) - 0.5) + 0.0078125 : 0.0;
            q1[i] = x;
            q0[i] = 0.0;
            acc[0] += x * v1[i];
        }
        lz_cta_sum(acc, slots, tid);
        const double pr = acc[0];
        acc[0] = 0;
        for (int i = tid; i < n; i += RADB_NTZ) {
            const double x = q1[i] - pr * v1[i];
            q1[i] = x;
            acc[0] += x * x;
        }
        lz_cta_sum(acc, slots, tid);
        const double rn = radb_div(1.0, radb_sqrt(acc[0]));
        for (int i = tid; i < n; i += RADB_NTZ) q1[i] *= rn;
    }
    __syncthreads();
    // ---- Lanczos on B = A - v1 v1^T.  Rows are dealt so that all threads work: up to 128 levels a row is shared by
    // tpr = 128 / npad consecutive lanes (entries interleaved, partial sums combined by xor shuffles); above that a
    // thread owns rows tid and tid + 128.
    int npad = 32;
    while (npad < n) npad <<= 1;
    const int tpr = npad <= RADB_NTZ ? RADB_NTZ / npad : 1;
    const int sub = tid & (tpr - 1), row0 = tpr > 1 ? tid / tpr : tid;
    const int nrows = npad > RADB_NTZ ? 2 : 1;
    for (int i = tid; i < n; i += RADB_NTZ) z[i] = rs[i] * q1[i];
    __syncthreads();
    double beta_prev = 0.0, rho = 0.0, emax = 0.0, emin = 0.0;
    int K = 0;
    bool done = false, have_prev = false;
    const int kcap = m < RADB_LZ_KMAX ? m : RADB_LZ_KMAX;  // k = m: the Krylov space is the whole space
#pragma unroll 1
    for (int k = 0; k < kcap && !done; k++) {
        double acc[3] = {0, 0, 0};  // v1 . y, q . y, v1 . q
        double yv[2] = {0, 0};
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int i = row0 + u * RADB_NTZ;
            if (u < nrows) {
                double y = 0;
                if (i < n) {
                    const int e1 = rowptr[i + 1];
                    for (int e = rowptr[i] + sub; e < e1; e += tpr) {
                        const unsigned w = ent[e];
                        y += (double)(w & 0xffffu) * z[w >> 16];
                    }
                }
                for (int mm = 1; mm < tpr; mm <<= 1) y += __shfl_xor_sync(FULLMASK, y, mm);
                if (i < n && sub == 0) {
                    y *= rs[i];
                    const double qi = q1[i], vi = v1[i];
                    acc[0] += vi * y;
                    acc[1] += qi * y;
                    acc[2] += vi * qi;
                }
                yv[u] = y;
            }
        }
        lz_cta_sum(acc, slots, tid);
        const double s1 = acc[0], alpha = acc[1] - acc[0] * acc[2];
        double nb[1] = {0};
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int i = row0 + u * RADB_NTZ;
            if (u < nrows && i < n && sub == 0) {
                const double w = yv[u] - s1 * v1[i] - alpha * q1[i] - beta_prev * q0[i];
                yv[u] = w;
                nb[0] += w * w;
            }
        }
        lz_cta_sum(nb, slots, tid);
        const double beta = radb_sqrt(nb[0]);
        if (tid == 0) { al[k] = alpha; be2[k] = nb[0]; }
        K = k + 1;
        const bool breakdown = !(beta > 1e-13);
        if (!breakdown) {
            const double rb = radb_div(1.0, beta);
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int i = row0 + u * RADB_NTZ;
                if (u < nrows && i < n && sub == 0) {
                    const double qn = yv[u] * rb;
                    q0[i] = q1[i];
                    q1[i] = qn;
                    z[i] = rs[i] * qn;
                }
            }
        }
        beta_prev = beta;
        __syncthreads();
        // check points: every 16 steps from 32 on, the last step, breakdown
        const bool check = breakdown || K == kcap || (K >= 32 && (K & 7) == 0);
        if (check) {
            // The extreme Ritz values move outwards monotonically with K (interlacing), so "nothing of T_K lies beyond the
            // previous extremes + 1e-12" proves convergence with two Sturm counts; only otherwise are they recomputed
            // (8 rounds of CTA-wide multisection)
            bool same = false;
            if (have_prev) {
                if (tid == 0) ishare[0] = sturm_count(al, be2, K, emax + 1e-12) == K;
                if (tid == 32) ishare[1] = sturm_count(al, be2, K, emin - 1e-12) == 0;
                __syncthreads();
                same = ishare[0] && ishare[1];
                __syncthreads();
            }
            // recomputed at K = 32, 48, 64, ... (and at the end); the check points in between only run the two-count test
            const bool full = !same && (!have_prev || breakdown || K == kcap || (K & 15) == 0);
            if (full) lz_extremes(al, be2, K, ishare, tid, emax, emin);  // Ritz values of T_K lie inside [-1, 1] like B's spectrum
            if (!same && !full) continue;
            have_prev = true;
            rho = fmax(fabs(emax), fabs(emin));
            if (breakdown || K == kcap || same) done = true;
        }
    }
    if (tid == 0) *out_mcc = rho;
}

// MCC column of the output row = mean of the per-angle values the Lanczos kernel left in the record header over the
// non-empty angles (NaN = empty angle), 1 for a one-level ROI (glcm.py's flat-region rule); one thread per patch.
// Runs after both radb_mcc_lanczos_kernel and radb_angle_kernel (which writes a placeholder into the column).
__device__ void radb_mcc_combine_thread(const RadbParams& p, long long patch)
{
    if (patch >= p.B || p.off_glcm < 0) return;
    const long long row = radb_row(p, patch);
    if (p.status[row] != 0) return;
    const unsigned char* rec = p.ws + patch * (long long)p.rec_bytes;
    const int* misc = (const int*)(rec + (p.o_misc - p.o_rec));
    const double* v = (const double*)(misc + RADB_REC_MCC_INT);
    double s = 0.0;
    int k = 0;
    for (int a = 0; a < p.n_angles; a++)
        if (v[a] == v[a]) { s += v[a]; k++; }
    double r = k ? s / (double)k : nan_f64();
    if (misc[9] < 2) r = 1.0;
    p.out[row * (long long)p.F + p.off_glcm + 19] = r;
}
