/* Plain-C restatement of pyradiomics 3.1.0 `radiomics/src/cmatrices.c` matrix builders for
 * 2-D arrays.  TEST INFRASTRUCTURE ONLY (the checker; never the product path).
 *
 * PARITY UNPINNED: the upstream source is not under /root/reference (third-party dependency,
 * pinned only by /root/reference/params.yml:24).  This file restates the published algorithm
 *  independently of oracle/radiomics_oracle.py; the two must
 * agree bit-exactly (tests/test_oracle.py).  Reached in the reference through
 * RadiomicExtractor.py:38,42,45,48 -> RadiomicsFeatureExtractor.execute -> cMatrices.calculate_*.
 *
 * Layout: `lev` is an int32 [H][W] discretised image, 0 outside the ROI, levels 1..Ng inside.
 * `ang` is int32 [Na][2] = (dy, dx).  All outputs are caller-zeroed.
 */
#include <stdint.h>
#include <stdlib.h>
#include <math.h>

#define IN(y, x) ((y) >= 0 && (y) < H && (x) >= 0 && (x) < W)

/* A.6 calculate_glcm: out int64 [Ng][Ng][Na]; symmetric => P += P^T (glcm.py:_applyMatrixOptions) */
void orc_glcm(const int32_t *lev, int H, int W, int Ng, const int32_t *ang, int Na, int symmetric,
              int64_t *out)
{
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int g = lev[y * W + x];
            if (!g) continue;
            for (int a = 0; a < Na; a++) {
                int yy = y + ang[2 * a], xx = x + ang[2 * a + 1];
                if (!IN(yy, xx)) continue;
                int n = lev[yy * W + xx];
                if (!n) continue;
                out[((int64_t)(g - 1) * Ng + (n - 1)) * Na + a] += 1;
                if (symmetric) out[((int64_t)(n - 1) * Ng + (g - 1)) * Na + a] += 1;
            }
        }
}

/* A.7 calculate_glrlm: out int64 [Ng][Nr][Na].  Walk every line of the array along the angle. */
void orc_glrlm(const int32_t *lev, int H, int W, int Ng, int Nr, const int32_t *ang, int Na,
               int64_t *out)
{
    (void)Ng;
    for (int a = 0; a < Na; a++) {
        int dy = ang[2 * a], dx = ang[2 * a + 1];
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                /* (y,x) starts a line iff its predecessor is outside the array */
                if (IN(y - dy, x - dx)) continue;
                int cy = y, cx = x, cur = 0, len = 0;
                while (IN(cy, cx)) {
                    int g = lev[cy * W + cx];
                    if (g == cur && g) {
                        len++;
                    } else {
                        if (cur) out[((int64_t)(cur - 1) * Nr + (len - 1)) * Na + a] += 1;
                        cur = g;
                        len = g ? 1 : 0;
                    }
                    cy += dy;
                    cx += dx;
                }
                if (cur) out[((int64_t)(cur - 1) * Nr + (len - 1)) * Na + a] += 1;
            }
    }
}

/* A.8 calculate_glszm: out int64 [Ng][Ns]; region growing with an explicit stack over the
 * bidirectional angle set `ang` [Nb][2]. */
int orc_glszm(const int32_t *lev, int H, int W, int Ng, int Ns, const int32_t *ang, int Nb,
              int64_t *out)
{
    (void)Ng;
    uint8_t *seen = (uint8_t *)calloc((size_t)H * W, 1);
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * (size_t)H * W);
    if (!seen || !stack) { free(seen); free(stack); return -1; }
    for (int p0 = 0; p0 < H * W; p0++) {
        int g = lev[p0];
        if (!g || seen[p0]) continue;
        int sp = 0, size = 0;
        stack[sp++] = p0;
        seen[p0] = 1;
        while (sp) {
            int p = stack[--sp];
            int y = p / W, x = p % W;
            size++;
            for (int a = 0; a < Nb; a++) {
                int yy = y + ang[2 * a], xx = x + ang[2 * a + 1];
                if (!IN(yy, xx)) continue;
                int q = yy * W + xx;
                if (seen[q] || lev[q] != g) continue;
                seen[q] = 1;
                stack[sp++] = q;
            }
        }
        out[(int64_t)(g - 1) * Ns + (size - 1)] += 1;
    }
    free(seen);
    free(stack);
    return 0;
}

/* A.9 calculate_gldm: out int64 [Ng][Nb+1] */
void orc_gldm(const int32_t *lev, int H, int W, int Ng, const int32_t *ang, int Nb, int alpha,
              int64_t *out)
{
    (void)Ng;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int g = lev[y * W + x];
            if (!g) continue;
            int dep = 0;
            for (int a = 0; a < Nb; a++) {
                int yy = y + ang[2 * a], xx = x + ang[2 * a + 1];
                if (!IN(yy, xx)) continue;
                int n = lev[yy * W + xx];
                if (n && abs(n - g) <= alpha) dep++;
            }
            out[(int64_t)(g - 1) * (Nb + 1) + dep] += 1;
        }
}

/* A.9 calculate_ngtdm: n int64 [Ng], s double [Ng] */
void orc_ngtdm(const int32_t *lev, int H, int W, int Ng, const int32_t *ang, int Nb, int64_t *n,
               double *s)
{
    (void)Ng;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int g = lev[y * W + x];
            if (!g) continue;
            int cnt = 0;
            double sum = 0;
            for (int a = 0; a < Nb; a++) {
                int yy = y + ang[2 * a], xx = x + ang[2 * a + 1];
                if (!IN(yy, xx)) continue;
                int v = lev[yy * W + xx];
                if (v) { cnt++; sum += v; }
            }
            if (cnt) {
                n[g - 1] += 1;
                s[g - 1] += fabs((double)g - sum / cnt);
            }
        }
}

/* A.3 binImage for a fixed bin width: levels via the fp64 edges numpy.arange would produce
 * (edge_k = low + k*bw) and numpy.digitize (number of edges <= x).  Returns Ng, or 0 if the
 * ROI is empty.  `img` is double [H][W], `roi` uint8 [H][W] (1 inside). */
int orc_bin_image(const double *img, const uint8_t *roi, int H, int W, double bw, int32_t *lev)
{
    double mn = 0, mx = 0;
    int any = 0;
    for (int p = 0; p < H * W; p++)
        if (roi[p]) {
            if (!any || img[p] < mn) mn = img[p];
            if (!any || img[p] > mx) mx = img[p];
            any = 1;
        }
    if (!any) return 0;
    double r = fmod(mn, bw);
    if (r != 0 && ((r < 0) != (bw < 0))) r += bw; /* Python modulo: sign of the divisor */
    double low = mn - r;
    int Ng = 0;
    for (int p = 0; p < H * W; p++) {
        lev[p] = 0;
        if (!roi[p]) continue;
        long k = (long)floor((img[p] - low) / bw);
        while (low + (double)k * bw > img[p]) k--;
        while (low + (double)(k + 1) * bw <= img[p]) k++;
        lev[p] = (int32_t)(k + 1);
        if (lev[p] > Ng) Ng = lev[p];
    }
    return Ng;
}
