// CPU emulation harness: runs the *device* code of multimodal-isic_b200/csrc/radb_kernels.cuh
// thread-for-thread on host threads (tests/emu/cuda_emu.h).  TEST INFRASTRUCTURE ONLY -- used
// by tests/ (marker: not gpu) to check kernel logic against the oracle without a GPU.
// Not part of the product; the product library has no CPU path.
#include "cuda_emu.h"
#include <string.h>
#include <string>
#include "../../multimodal-isic_b200/csrc/radb_host.h"
#include "../../multimodal-isic_b200/csrc/radb_kernels.cuh"

static std::string g_err;

extern "C" const char* radb_emu_last_error() { return g_err.c_str(); }

extern "C" int radb_emu_feature_count(const radb_settings* s)
{
    radb::Plan pl;
    if (radb::make_plan(*s, pl, g_err)) return -1;
    return pl.F;
}
extern "C" int radb_emu_is_wide(const radb_settings* s, int H, int W)
{
    radb::Plan pl;
    if (radb::make_plan(*s, pl, g_err)) return -1;
    RadbParams p;
    if (radb::fill_params(pl, H, W, RADB_DTYPE_U8, p, g_err)) return -1;
    return p.wide;
}
extern "C" int radb_emu_max_ng(const radb_settings* s)
{
    radb::Plan pl;
    if (radb::make_plan(*s, pl, g_err)) return -1;
    return pl.max_ng;
}

// the launches radb_api.cu issues for one same-size batch, CTA by CTA on host threads
static int emu_launch(RadbParams& p, int dtype, int64_t B)
{
    std::vector<double> inv2, tlog;
    radb::make_tables(p.ninv, inv2, tlog);
    if (getenv("RADB_EMU_PERTURB_TABLES"))  // host-libm vs device-libm: the tables may differ by an ulp
        for (size_t k = 2; k < tlog.size(); k++) tlog[k] = nextafter(tlog[k], (k & 1) ? 1e9 : -1e9);
    p.g_inv2 = inv2.data();
    p.g_tlog = tlog.data();
    std::vector<unsigned char> ws((size_t)B * p.rec_bytes + 64), scr((size_t)B * p.scr_bytes + 64);
    p.ws = (unsigned char*)(((uintptr_t)ws.data() + 15) & ~(uintptr_t)15);
    p.ws_scr = (unsigned char*)(((uintptr_t)scr.data() + 15) & ~(uintptr_t)15);
    int mx = p.smem_total > p.a_smem_total ? p.smem_total : p.a_smem_total;
    mx = mx > p.m_smem_total ? mx : p.m_smem_total;
    mx = mx > p.s_smem_total ? mx : p.s_smem_total;
    if (getenv("RADB_NO_LANE")) p.use_lane = 0;
    if (p.use_lane && p.l_smem_total > mx) mx = p.l_smem_total;
    if (p.use_lane == 2 && p.g8_smem_total > mx) mx = p.g8_smem_total;
    if (p.use_lane) p.use_lanczos = 0;
    if (p.use_lanczos && p.z_smem_total > mx) mx = p.z_smem_total;
    p.only_big_ovf = (!getenv("RADB_NO_LANE") && p.ml_smem_total <= 96 * 1024) ? 1 : 0;
    if (p.only_big_ovf && p.ml_smem_total > mx) mx = p.ml_smem_total;
    std::vector<unsigned char> smem((size_t)mx + 64);
    unsigned char* sm = (unsigned char*)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    // the same three launches radb_api.cu issues, CTA by CTA on host threads
#define EMU_BUILD(PT)                                                                                             \
    if (p.lev_bytes == 2)                                                                                         \
        emu::launch((unsigned)B, RADB_NTB, [&]() { radb_build_cta<PT, true, true, true>(p, (long long)blockIdx.x, sm); }); \
    else if (p.wide)                                                                                                   \
        emu::launch((unsigned)B, RADB_NTB, [&]() { radb_build_cta<PT, true, true>(p, (long long)blockIdx.x, sm); }); \
    else                                                                                                          \
        emu::launch((unsigned)B, RADB_NTB, [&]() { radb_build_cta<PT, true, false>(p, (long long)blockIdx.x, sm); });
    switch (dtype) {
        case RADB_DTYPE_U8: EMU_BUILD(unsigned char) break;
        case RADB_DTYPE_U16: EMU_BUILD(unsigned short) break;
        case RADB_DTYPE_F32: EMU_BUILD(float) break;
        case RADB_DTYPE_F64: EMU_BUILD(double) break;
        default: g_err = "unknown dtype"; return -1;
    }
    if (p.use_lanczos && p.off_glcm >= 0)
        emu::launch((unsigned)(B * p.n_angles), RADB_NTZ, [&]() { radb_mcc_lanczos_cta(p, (long long)blockIdx.x, sm); });
    if (p.use_lane == 2 && p.off_glcm >= 0)
    {
        const int na = p.n_angles;
        const long long g8_warps = (na == 1 || na == 2 || na == 4) ? (B * na + 3) / 4 : B;  // as radb_api.cu: launch
        emu::launch((unsigned)((g8_warps + RADB_NTM / 32 - 1) / (RADB_NTM / 32)), RADB_NTM, [&]() { radb_mcc_g8_cta(p, (long long)blockIdx.x, sm); });
    }
    if (p.use_lane)
        emu::launch((unsigned)((B * p.l_nap + RADB_NTL - 1) / RADB_NTL), RADB_NTL, [&]() { radb_angle_lane_cta(p, (long long)blockIdx.x, sm); });
    else
        emu::launch((unsigned)B, RADB_NT, [&]() { radb_angle_cta(p, (long long)blockIdx.x, sm); });
    if (p.use_lanczos && p.off_glcm >= 0)
        for (long long b = 0; b < B; b++) radb_mcc_combine_thread(p, b);
    if (p.only_big_ovf)
        emu::launch((unsigned)((B + RADB_NT - 1) / RADB_NT), RADB_NT, [&]() { radb_misc_lane_cta(p, (long long)blockIdx.x, sm); });
    emu::launch((unsigned)B, RADB_NT, [&]() { radb_misc_cta(p, (long long)blockIdx.x, sm); });
    if (p.off_shape >= 0) emu::launch((unsigned)B, RADB_NT, [&]() { radb_shape_cta(p, (long long)blockIdx.x, sm); });
    return 0;
}

static int g_mask_bits = 0;
// the next radb_emu_extract call reads `mask` as bit-packed streams (radb_extract_packed)
extern "C" void radb_emu_set_mask_bits(int on) { g_mask_bits = on ? 1 : 0; }

// Host-pointer twin of radb_debug_matrices.
extern "C" int radb_emu_extract(const radb_settings* s, const void* img, int dtype, const uint8_t* mask, int64_t B,
                                int H, int W, int64_t img_stride_b, int64_t mask_stride_b, double* out,
                                int32_t* status, int32_t* levels, int32_t* glcm, int32_t* glrlm, int32_t* glszm,
                                int32_t* gldm, int32_t* ngtdm_n, double* ngtdm_s, int32_t* ng)
{
    radb::Plan pl;
    int rc = radb::make_plan(*s, pl, g_err);
    if (rc) return rc;
    RadbParams p;
    rc = radb::fill_params(pl, H, W, dtype, p, g_err);
    if (rc) return rc;
    p.img = img;
    p.mask = mask;
    p.img_stride = img_stride_b;
    p.mask_stride = mask_stride_b;
    p.mask_bits = g_mask_bits;
    p.out = out;
    p.status = status;
    p.B = B;
    p.dbg_levels = levels;
    p.dbg_glcm = glcm;
    p.dbg_glrlm = glrlm;
    p.dbg_glszm = glszm;
    p.dbg_gldm = gldm;
    p.dbg_ngn = ngtdm_n;
    p.dbg_ngs = ngtdm_s;
    p.dbg_ng = ng;
    return emu_launch(p, dtype, B);
}

// Host twin of radb_bgr_planes_kernel (one "thread" per 4 pixels).
extern "C" void radb_emu_bgr_planes(const uint8_t* bgr, uint8_t* planes, int64_t n_images, int64_t HW)
{
    const long long threads = n_images * ((HW + 3) / 4);
    for (long long q = 0; q < threads; q++) radb_bgr_planes_thread(bgr, planes, n_images, HW, q);
}

// Host twin of radb_resize_mask (same reciprocal scales as radb_api.cu, same per-pixel function).
extern "C" void radb_emu_resize_mask(const uint8_t* src, int64_t n, int sH, int sW, uint8_t* dst, int dH, int dW)
{
    const double ify = 1.0 / ((double)dH / (double)sH), ifx = 1.0 / ((double)dW / (double)sW);
    const long long threads = n * (long long)dH * dW;
    for (long long q = 0; q < threads; q++) radb_resize_mask_thread(src, sH, sW, dst, dH, dW, n, ify, ifx, q);
}

// Host twin of radb_image_max_kernel + radb_derive_kernel.
extern "C" void radb_emu_derive(const uint8_t* img, int64_t n_images, int64_t HW, int type, double* out)
{
    for (int64_t i = 0; i < n_images; i++) {
        int m = 0;
        for (int64_t k = 0; k < HW; k++) m = img[i * HW + k] > m ? img[i * HW + k] : m;
        for (int64_t k = 0; k < HW; k++) out[i * HW + k] = radb_derive_px(type, (double)img[i * HW + k], (double)m);
    }
}

// Host-pointer twin of radb_extract_ragged: same grouping (radb::group_ragged), same per-group launches.
extern "C" int radb_emu_extract_ragged(const radb_settings* s, const void* img_pool, int dtype, const uint8_t* mask_pool,
                                       int64_t n, const int64_t* img_off, const int64_t* mask_off, const int32_t* hw,
                                       double* out, int32_t* status)
{
    radb::Plan pl;
    int rc = radb::make_plan(*s, pl, g_err);
    if (rc) return rc;
    std::vector<radb::RaggedGroup> groups;
    rc = radb::group_ragged(n, hw, groups, g_err);
    if (rc) return rc;
    for (const auto& g : groups) {
        RadbParams p;
        rc = radb::fill_params(pl, g.H, g.W, dtype, p, g_err);
        if (rc) return rc;
        std::vector<long long> io, mo, rows;
        for (long long i : g.idx) { io.push_back(img_off[i]); mo.push_back(mask_off[i]); rows.push_back(i); }
        p.img = img_pool;
        p.mask = mask_pool;
        p.img_off = io.data();
        p.mask_off = mo.data();
        p.rows = rows.data();
        p.out = out;
        p.status = status;
        p.B = (long long)g.idx.size();
        rc = emu_launch(p, dtype, p.B);
        if (rc) return rc;
    }
    return 0;
}

// Host twin of radb_filter_image (Gradient / LoG / Wavelet kernels of radb_filters.cuh), one "thread" per loop index.
extern "C" int radb_emu_filter(const uint8_t* img, int64_t n, int H, int W, int type, double param, int flags, void* out,
                               void* scratch)
{
    const long long HW = (long long)H * W, npx = n * HW;
    if (type == RADB_IT_GRADIENT) {
        for (long long t = 0; t < npx; t++) radb_gradient_px(img, n, H, W, (float*)out, t);
    } else if (type == RADB_IT_WAVELET) {
        if (flags & 1) {
            for (long long t = 0; t < npx; t++) radb_wavelet_x_px(img, n, H, W, (double*)out, t);
        } else {
            const long long npp = n * (long long)(H + (H & 1)) * (W + (W & 1));
            for (long long t = 0; t < npp; t++) radb_wavelet_rows_px(img, n, H, W, (double*)scratch, t);
            for (long long t = 0; t < npx; t++) radb_wavelet_cols_px((const double*)scratch, n, H, W, (double*)out, t);
        }
    } else if (type == RADB_IT_LOG) {
        if (H < 4 || W < 4) return -1;
        RadbIir c0, c2;
        radb::deriche_coefficients(param, 0, c0.N, c0.D, c0.M, c0.BN, c0.BM);
        radb::deriche_coefficients(param, 2, c2.N, c2.D, c2.M, c2.BN, c2.BM);
        double* scr = (double*)scratch;
        float* tmp = (float*)(scr + npx);
        for (long long t = 0; t < n * H; t++) radb_iir_thread<true>(img, tmp, scr, n, H, W, 0, c2, 0, t);
        for (long long t = 0; t < n * W; t++) radb_iir_thread<false>(tmp, (float*)out, scr, n, H, W, 1, c0, 0, t);
        for (long long t = 0; t < n * W; t++) radb_iir_thread<true>(img, tmp, scr, n, H, W, 1, c2, 0, t);
        for (long long t = 0; t < n * H; t++) radb_iir_thread<false>(tmp, (float*)out, scr, n, H, W, 0, c0, 1, t);
    } else {
        return -3;
    }
    return 0;
}
