// Filtered image types of the pyradiomics parameter file (/root/reference/params.yml:138-145): Gradient, LoG and
// Wavelet of uint8 images, as device kernels feeding radb_extract (RADB_DTYPE_F32 / RADB_DTYPE_F64).
// Included by radb_kernels.cuh.  Algorithms and operation order follow oracle/image_filters.py (restatements of
// ITK's GradientMagnitude / LaplacianRecursiveGaussian filters and PyWavelets' level-1 stationary transform);
// products and sums are rounded separately (__dmul_rn / __dadd_rn: no fused multiply-add), as the libraries'
// x86-64 builds do, so the kernels can be compared with the oracle bit for bit.
#pragma once

enum { RADB_IT_GRADIENT = 5, RADB_IT_LOG = 6, RADB_IT_WAVELET = 7 };

// ------------------------------------------------------------------ Gradient
// sitk.GradientMagnitudeImageFilter: central differences 0.5 * (f[i+1] - f[i-1]) per axis, ZeroFluxNeumann border
// (out-of-range neighbours take the border pixel), magnitude in double, float32 output pixels.
__device__ __forceinline__ void radb_gradient_px(const unsigned char* img, long long n_images, int H, int W, float* out,
                                                 long long t)
{
    const long long HW = (long long)H * W;
    if (t >= n_images * HW) return;
    const long long im = t / HW;
    const int r = (int)(t - im * HW), y = r / W, x = r - y * W;
    const unsigned char* s = img + im * HW;
    const int xm = x > 0 ? x - 1 : 0, xp = x < W - 1 ? x + 1 : W - 1, ym = y > 0 ? y - 1 : 0, yp = y < H - 1 ? y + 1 : H - 1;
    const double dx = 0.5 * ((double)s[y * W + xp] - (double)s[y * W + xm]);
    const double dy = 0.5 * ((double)s[yp * W + x] - (double)s[ym * W + x]);
    out[t] = (float)sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// ------------------------------------------------------------------ Wavelet (coif1, level 1, periodic)
// out[n] = sum_j taps[j] * x[(n + 3 - j) mod Np], accumulated in tap order; odd sizes are padded by one wrapped
// sample (index N -> 0) and cropped afterwards (pyradiomics imageoperations._swt3).
__device__ __forceinline__ double radb_coif1(int hi, int j)
{
    const double lo[6] = {-0.01565572813546454, -0.0727326195128539, 0.38486484686420286, 0.8525720202122554,
                          0.3378976624578092, -0.0727326195128539};
    const double hh[6] = {0.0727326195128539, 0.3378976624578092, -0.8525720202122554, 0.38486484686420286,
                          0.0727326195128539, -0.01565572813546454};
    return hi ? hh[j] : lo[j];
}
__device__ __forceinline__ int radb_wrap_pad(int i, int N, int Np)
{
    // index into the padded periodic signal of length Np (= N or N + 1) -> index into the original samples
    i %= Np;
    if (i < 0) i += Np;
    return i == N ? 0 : i;
}
// row pass: src uint8 [n][H][W] -> tmp double [n][2][Hp][Wp] (band 0 = low / 'a', band 1 = high / 'd' along x);
// the padded row H (when H is odd) repeats row 0
__device__ __forceinline__ void radb_wavelet_rows_px(const unsigned char* img, long long n_images, int H, int W, double* tmp,
                                                     long long t)
{
    const int Hp = H + (H & 1), Wp = W + (W & 1);
    const long long HWp = (long long)Hp * Wp;
    if (t >= n_images * HWp) return;
    const long long im = t / HWp;
    const int r = (int)(t - im * HWp), y = r / Wp, x = r - y * Wp;
    const unsigned char* s = img + im * (long long)H * W + (long long)(y == H ? 0 : y) * W;
    double a = 0.0, d = 0.0;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        const double v = (double)s[radb_wrap_pad(x + 3 - j, W, Wp)];
        a = __dadd_rn(a, __dmul_rn(radb_coif1(0, j), v));
        d = __dadd_rn(d, __dmul_rn(radb_coif1(1, j), v));
    }
    tmp[(im * 2 + 0) * HWp + r] = a;
    tmp[(im * 2 + 1) * HWp + r] = d;
}
// x-only transform (force2D on a 2-D array, oracle/U1_ANGLES.md): out double [n][2][H][W] = wavelet-H, wavelet-L
__device__ __forceinline__ void radb_wavelet_x_px(const unsigned char* img, long long n_images, int H, int W, double* out,
                                                  long long t)
{
    const long long HW = (long long)H * W;
    if (t >= n_images * HW) return;
    const int Wp = W + (W & 1);
    const long long im = t / HW;
    const int r = (int)(t - im * HW), y = r / W, x = r - y * W;
    const unsigned char* s = img + im * HW + (long long)y * W;
    double a = 0.0, d = 0.0;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        const double v = (double)s[radb_wrap_pad(x + 3 - j, W, Wp)];
        a = __dadd_rn(a, __dmul_rn(radb_coif1(0, j), v));
        d = __dadd_rn(d, __dmul_rn(radb_coif1(1, j), v));
    }
    out[(im * 2 + 0) * HW + r] = d;  // pyradiomics yields the detail band(s) first, the approximation last
    out[(im * 2 + 1) * HW + r] = a;
}
// column pass: tmp [n][2][Hp][Wp] -> out double [n][4][H][W] in pyradiomics' order LH, HL, HH, LL
// (first letter <-> x, second <-> y)
__device__ __forceinline__ void radb_wavelet_cols_px(const double* tmp, long long n_images, int H, int W, double* out, long long t)
{
    const long long HW = (long long)H * W;
    if (t >= n_images * HW) return;
    const int Hp = H + (H & 1), Wp = W + (W & 1);
    const long long HWp = (long long)Hp * Wp;
    const long long im = t / HW;
    const int r = (int)(t - im * HW), y = r / W, x = r - y * W;
    double v[2][2] = {{0.0, 0.0}, {0.0, 0.0}};  // [x band][y band]
#pragma unroll
    for (int j = 0; j < 6; j++) {
        int yy = (y + 3 - j) % Hp;
        if (yy < 0) yy += Hp;  // tmp holds the padded row explicitly
#pragma unroll
        for (int bx = 0; bx < 2; bx++) {
            const double s = tmp[(im * 2 + bx) * HWp + (long long)yy * Wp + x];
            v[bx][0] = __dadd_rn(v[bx][0], __dmul_rn(radb_coif1(0, j), s));
            v[bx][1] = __dadd_rn(v[bx][1], __dmul_rn(radb_coif1(1, j), s));
        }
    }
    out[(im * 4 + 0) * HW + r] = v[0][1];  // LH
    out[(im * 4 + 1) * HW + r] = v[1][0];  // HL
    out[(im * 4 + 2) * HW + r] = v[1][1];  // HH
    out[(im * 4 + 3) * HW + r] = v[0][0];  // LL
}

// ------------------------------------------------------------------ LoG: ITK recursive Gaussian, one thread per line
struct RadbIir {
    double N[4], D[4], M[4], BN[4], BM[4];
};
#define RADB_E4(a, b, c, d, e, f, g, h) __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a, b), __dmul_rn(c, d)), __dmul_rn(e, f)), __dmul_rn(g, h))
// One line of ITK's RecursiveSeparableImageFilter::FilterDataArray: the anti-causal recursion goes to `scr` (double,
// same indexing as the line), then the causal recursion runs and writes float32(causal + anti-causal) to dst
// (added to dst's previous content in float32 when `accumulate`).  `src` is uint8 (SRC8) or float32.
template <bool SRC8>
__device__ void radb_iir_line(const void* src, float* dst, double* scr, int ln, long long step, const RadbIir& c, int accumulate)
{
    auto X = [&](int i) -> double {
        return SRC8 ? (double)((const unsigned char*)src)[i * step] : (double)((const float*)src)[i * step];
    };
    const double w = X(ln - 1);
    double s0, s1, s2, s3;  // anti-causal outputs at i, i+1, i+2, i+3
    {
        double t3 = RADB_E4(w, c.M[0], w, c.M[1], w, c.M[2], w, c.M[3]);
        double t2 = RADB_E4(X(ln - 1), c.M[0], w, c.M[1], w, c.M[2], w, c.M[3]);
        double t1 = RADB_E4(X(ln - 2), c.M[0], X(ln - 1), c.M[1], w, c.M[2], w, c.M[3]);
        double t0 = RADB_E4(X(ln - 3), c.M[0], X(ln - 2), c.M[1], X(ln - 1), c.M[2], w, c.M[3]);
        t3 = __dadd_rn(t3, -RADB_E4(w, c.BM[0], w, c.BM[1], w, c.BM[2], w, c.BM[3]));
        t2 = __dadd_rn(t2, -RADB_E4(t3, c.D[0], w, c.BM[1], w, c.BM[2], w, c.BM[3]));
        t1 = __dadd_rn(t1, -RADB_E4(t2, c.D[0], t3, c.D[1], w, c.BM[2], w, c.BM[3]));
        t0 = __dadd_rn(t0, -RADB_E4(t1, c.D[0], t2, c.D[1], t3, c.D[2], w, c.BM[3]));
        scr[(ln - 1) * step] = t3;
        scr[(ln - 2) * step] = t2;
        scr[(ln - 3) * step] = t1;
        scr[(ln - 4) * step] = t0;
        s0 = t0; s1 = t1; s2 = t2; s3 = t3;
    }
    for (int i = ln - 4; i > 0; i--) {
        double t = RADB_E4(X(i), c.M[0], X(i + 1), c.M[1], X(i + 2), c.M[2], X(i + 3), c.M[3]);
        t = __dadd_rn(t, -RADB_E4(s0, c.D[0], s1, c.D[1], s2, c.D[2], s3, c.D[3]));
        scr[(i - 1) * step] = t;
        s3 = s2; s2 = s1; s1 = s0; s0 = t;
    }
    const double v = X(0);
    double c0, c1, c2, c3;  // causal outputs at i-1, i-2, i-3, i-4 (c0 most recent)
    auto emit = [&](int i, double causal) {
        const float r = (float)__dadd_rn(causal, scr[i * step]);
        dst[i * step] = accumulate ? __fadd_rn(dst[i * step], r) : r;
    };
    {
        double t0 = RADB_E4(v, c.N[0], v, c.N[1], v, c.N[2], v, c.N[3]);
        double t1 = RADB_E4(X(1), c.N[0], v, c.N[1], v, c.N[2], v, c.N[3]);
        double t2 = RADB_E4(X(2), c.N[0], X(1), c.N[1], v, c.N[2], v, c.N[3]);
        double t3 = RADB_E4(X(3), c.N[0], X(2), c.N[1], X(1), c.N[2], v, c.N[3]);
        t0 = __dadd_rn(t0, -RADB_E4(v, c.BN[0], v, c.BN[1], v, c.BN[2], v, c.BN[3]));
        t1 = __dadd_rn(t1, -RADB_E4(t0, c.D[0], v, c.BN[1], v, c.BN[2], v, c.BN[3]));
        t2 = __dadd_rn(t2, -RADB_E4(t1, c.D[0], t0, c.D[1], v, c.BN[2], v, c.BN[3]));
        t3 = __dadd_rn(t3, -RADB_E4(t2, c.D[0], t1, c.D[1], t0, c.D[2], v, c.BN[3]));
        emit(0, t0); emit(1, t1); emit(2, t2); emit(3, t3);
        c0 = t3; c1 = t2; c2 = t1; c3 = t0;
    }
    for (int i = 4; i < ln; i++) {
        double t = RADB_E4(X(i), c.N[0], X(i - 1), c.N[1], X(i - 2), c.N[2], X(i - 3), c.N[3]);
        t = __dadd_rn(t, -RADB_E4(c0, c.D[0], c1, c.D[1], c2, c.D[2], c3, c.D[3]));
        emit(i, t);
        c3 = c2; c2 = c1; c1 = c0; c0 = t;
    }
}
// lines along x (along_y = 0: one thread per row) or along y (one thread per column) of every image
template <bool SRC8>
__device__ __forceinline__ void radb_iir_thread(const void* src, float* dst, double* scr, long long n_images, int H, int W,
                                                int along_y, RadbIir c, int accumulate, long long t)
{
    const int lines = along_y ? W : H, ln = along_y ? H : W;
    if (t >= n_images * lines) return;
    const long long im = t / lines;
    const int l = (int)(t - im * lines);
    const long long base = im * (long long)H * W + (along_y ? l : (long long)l * W);
    const long long step = along_y ? W : 1;
    const void* s = SRC8 ? (const void*)((const unsigned char*)src + base) : (const void*)((const float*)src + base);
    radb_iir_line<SRC8>(s, dst + base, scr + base, ln, step, c, accumulate);
}

#ifndef RADB_EMU
__global__ void radb_gradient_kernel(const unsigned char* img, long long n, int H, int W, float* out)
{
    radb_gradient_px(img, n, H, W, out, (long long)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void radb_wavelet_rows_kernel(const unsigned char* img, long long n, int H, int W, double* tmp)
{
    radb_wavelet_rows_px(img, n, H, W, tmp, (long long)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void radb_wavelet_x_kernel(const unsigned char* img, long long n, int H, int W, double* out)
{
    radb_wavelet_x_px(img, n, H, W, out, (long long)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void radb_wavelet_cols_kernel(const double* tmp, long long n, int H, int W, double* out)
{
    radb_wavelet_cols_px(tmp, n, H, W, out, (long long)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void radb_iir_u8_kernel(const unsigned char* src, float* dst, double* scr, long long n, int H, int W, int along_y,
                                   RadbIir c, int accumulate)
{
    radb_iir_thread<true>(src, dst, scr, n, H, W, along_y, c, accumulate, (long long)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void radb_iir_f32_kernel(const float* src, float* dst, double* scr, long long n, int H, int W, int along_y,
                                    RadbIir c, int accumulate)
{
    radb_iir_thread<false>(src, dst, scr, n, H, W, along_y, c, accumulate, (long long)blockIdx.x * blockDim.x + threadIdx.x);
}
#endif
