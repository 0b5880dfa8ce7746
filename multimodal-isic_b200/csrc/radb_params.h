// Launch parameters + shared-memory layout shared by the C-ABI host code (radb_api.cu),
// the kernels (radb_kernels.cuh) and the CPU emulation harness under tests/emu.
#pragma once
#include <stdint.h>
#include <stdlib.h>

#define RADB_NT 128          // threads per CTA of the reduction kernels (4 warps); one CTA per patch
#ifndef RADB_NTB
#define RADB_NTB 256         // threads per CTA of the build kernel
#endif
#ifndef RADB_NTB_MINB
#define RADB_NTB_MINB 5      // min resident build CTAs per SM (register cap: 48)
#endif
#ifndef RADB_WALK_UNROLL
#define RADB_WALK_UNROLL 4   // unroll factor of the line-walk loops of the build kernel
#endif
#define RADB_PRAGMA_(x) _Pragma(#x)
#define RADB_UNROLL(n) RADB_PRAGMA_(unroll n)
#ifndef RADB_NTL
#define RADB_NTL 64          // threads per CTA of the lane kernel: one THREAD per (patch, angle)
#endif
#define RADB_LSTRIDE RADB_NTL   // fp64 slot stride of its per-thread scratch (slot-major)
#define RADB_NTM 64          // threads per CTA of the MCC kernel: one WARP per patch, 8 lanes per angle
#define RADB_MAX_ANGLES 4    // unidirectional offsets at distance 1 in a plane
#define RADB_GLCM_NF 24
#define RADB_GLRLM_NF 16
#define RADB_FSC_STRIDE 40   // per-angle scratch: 24 GLCM + 16 GLRLM features
#define RADB_REC_MCC_INT 16  // record header ints [16, 24): the per-angle MCC values (4 doubles) of radb_mcc_g8_kernel

enum { RADB_U8 = 0, RADB_U16 = 1, RADB_F32 = 2, RADB_F64 = 3 };

struct RadbParams {
    const void* img;
    const uint8_t* mask;
    long long img_stride;   // bytes between patches
    long long mask_stride;  // bytes between masks
    int mask_group;         // consecutive patches sharing one mask (1; 4 for the gray/R/G/B planes of an image)
    // ragged batches (radb_extract_ragged): patch k of a same-size group reads its pixels / mask at byte
    // offsets img_off[k] / mask_off[k] from img / mask and writes row rows[k] of out / status (null: dense)
    const long long* img_off;
    const long long* mask_off;
    const long long* rows;
    double* out;            // [B][F]
    int* status;            // [B]
    long long B;
    int H, W, WP, HW;
    int lp;        // pitch (words) of the union-find array: odd in narrow mode so that a warp whose lanes walk 32 rows hits 32 banks
    int mask_bits; // 1: `mask` is bit-packed (bit i of a patch's stream <-> pixel i, set = ROI), mask_stride in bytes of that stream
    int xo;        // column of pixel x = 0 inside a padded level-image row (left border width)
    int vec4;      // uint8 narrow patches with W % 4 == 0: 4 pixels per shared-memory load in the discretise phases
    int lev4;      // narrow u8 level image with W % 4 == 0: pixel x at byte 4 + x of a row whose word pitch is odd
                   // (the neighbourhood pass then works on aligned 4-pixel words, radb_build_cta)
    int glcm_pad;  // 1: the build kernel counts into a (Ng + 1)^2 GLCM whose row / column 0 collect the pairs with a
                   // pixel outside the ROI (no test per increment); it is written compact (Ng^2) to the global record
    int uq_cap;    // entries per warp of the union request queues
    int label;
    int n_angles;
    int ang_y[RADB_MAX_ANGLES], ang_x[RADB_MAX_ANGLES];
    int symmetric;
    int alpha;
    double bin_width;
    int bin_count; // > 0: binCount binning (the ROI range split into this many bins), bin_width is ignored
    int bw_int;    // bin_width when it is an integer in 1..255 (uint8 fast path of the level LUT), else 0
    double shift;
    int max_ng;
    int nr;        // GLRLM columns = max(H, W)
    int nrp;       // pitch of a GLRLM row (nr rounded up to even: a row of packed u16 counters starts on a word)
    int s0;        // dense GLSZM columns (zone sizes 1..s0); larger zones go to the overflow list
    int ovf_cap;
    int F;
    int off_fo, off_glcm, off_gldm, off_glrlm, off_glszm, off_ngtdm;  // output column of each class, -1 = off
    int off_shape;                                                    // shape2D (always first when enabled)
    int use_tma;
    int pix_bytes; // 1 uint8, 2 uint16, 4 float32, 8 float64
    int wide;      // 1: whole-image mode (level image, union-find words, GLRLM and overflow list in global memory)
    int big;       // 1 (implies wide): many gray levels -- the GLCM counters and the MCC workspace live in global memory too
    int lev_bytes; // bytes per pixel of the level image: 1, or 2 when max_ng > 255
    long long g_mcc;  // big mode: byte offset of the per-angle MCC workspaces inside the per-patch global scratch
    // ---- build kernel: shared-memory byte offsets.  [o_rec, o_rec + rec_bytes) is the per-patch
    // RECORD (header + every integer matrix); the build kernel copies it to the global workspace
    // and the reduction kernels read it from there at the same relative offsets.
    int o_runs;    // narrow mode: u16 list of the row-run start pixels (-1: not kept, the zone phases scan the bbox)
    int o_stage, o_mask, o_mbar, o_zero, o_lev, o_uq, o_lut, o_fo, o_rec, o_misc, o_hist, o_lhist, o_glcm, o_glrlm,
        o_gldm, o_ngc, o_ngn, o_szm, o_ovf, smem_total;
    int rec_bytes;       // record size in the global workspace
    int rec_copy_bytes;  // leading part of the record that is built in shared memory and copied out
    long long scr_bytes; // wide mode: per-patch global scratch (level image + union-find words)
    long long g_lev, g_lab;  // offsets inside that scratch
    int glrlm_stride;  // bytes per angle of the packed-u16 GLRLM counters
    int glrlm_dense;   // 16: narrow mode counts runs of length <= 16 in u32 [na][ng][16] shared-memory counters; 0: packed u16 in place
    int mcc_stride;    // doubles per angle in the MCC workspace
    int ninv;          // entries of the 1/k^2 table
    // ---- angle kernel (GLRLM + GLCM + MCC, one warp per angle): per-warp scratch + CTA scratch
    int a_px, a_py, a_padd, a_psub, a_pr, a_idx, a_mcc, a_red, a_warp_bytes, a_fsc, a_valid, a_smem_total;
    // ---- misc kernel (GLSZM, GLDM, NGTDM, first-order: one warp each)
    int m_pg, m_ovf2, m_ngp, m_qv, m_red, m_smem_total;
    int s_smem_total;  // shape kernel
    // ---- lane kernel (radb_lane.cuh): one thread per (patch, angle), l_doubles fp64 slots of shared memory each
    int use_lane, l_nap, l_doubles, l_smem_total;
    // use_lane: 0 warp-per-angle kernel | 1 thread-per-angle kernel incl. MCC (Ng <= 14) | 2 thread-per-angle kernel
    // without the matrix storage + radb_mcc_g8_kernel for the eigenproblems (Ng <= 40)
    int g8_px, g8_idx, g8_mcc, g8_group_bytes, g8_smem_total;  // MCC kernel: per-group shared-memory layout
    // ---- Lanczos MCC kernel (radb_lanczos.cuh): more than 40 gray levels, symmetric GLCM, one CTA per (patch, angle)
    int use_lanczos, z_rowptr, z_ent, z_vec, z_tri, z_cap, z_smem_total;
    // ---- misc lane kernel: one thread per (patch, class); the warp-level misc kernel then only serves GLSZM
    // patches with a long overflow list (only_big_ovf = 1)
    int ml_doubles, ml_smem_total, only_big_ovf;
    // global workspace + tables (device pointers)
    unsigned char* ws;        // [B][rec_bytes]
    unsigned char* ws_scr;    // [B][scr_bytes] (wide mode)
    const double* g_inv2;     // [ninv]  1/k^2
    const double* g_tlog;     // [2048]  log2(k)
    // optional debug outputs (device pointers, may be null); dims use max_ng
    int* dbg_levels;   // [B][H][W]
    int* dbg_glcm;     // [B][Na][max_ng][max_ng]
    int* dbg_glrlm;    // [B][Na][max_ng][nr]
    int* dbg_glszm;    // [B][max_ng][HW]
    int* dbg_gldm;     // [B][max_ng][2*Na+1]
    int* dbg_ngn;      // [B][max_ng]
    double* dbg_ngs;   // [B][max_ng]
    int* dbg_ng;       // [B]
};

static inline int radb_align(int v, int a) { return (v + a - 1) / a * a; }
// A/B switch (tests, profiling): RADB_NO_LANCZOS=1 keeps the dense Householder MCC of the warp-per-angle kernel
static inline int radb_no_lanczos(void) { return getenv("RADB_NO_LANCZOS") != 0; }

// Fills WP/HW/nr/s0/ovf_cap and every shared-memory / record offset from H, W, max_ng, n_angles
// and p->wide.  Record layout (both modes): header, hist, lhist, glcm, gldm, ngc, ngn, szm | ovf, glrlm.
static inline void radb_layout(RadbParams* p, int pix_bytes)
{
    const int H = p->H, W = p->W, ng = p->max_ng, na = p->n_angles, wide = p->wide, big = p->big;
    p->lev_bytes = ng > 255 ? 2 : 1;
    p->pix_bytes = pix_bytes;
    p->HW = H * W;
    p->lp = (!wide && W % 2 == 0) ? W + 1 : W;
    p->lev4 = (!wide && p->lev_bytes == 1 && W % 4 == 0) ? 1 : 0;
    p->vec4 = (p->lev4 && pix_bytes == 1) ? 1 : 0;
    p->glcm_pad = big ? 0 : 1;
    p->uq_cap = 96;
    p->xo = p->lev4 ? 4 : 1;
    p->WP = radb_align(W + p->xo + 1, 4);
    if (p->lev4 && (p->WP / 4) % 2 == 0) p->WP += 4;  // odd word stride: row walks stay bank-conflict free
    p->nr = H > W ? H : W;
    p->nrp = (p->nr + 1) & ~1;
    p->s0 = wide ? 64 : 16;
    p->ovf_cap = p->HW / (p->s0 + 1) + 1;
    p->ninv = ng > p->nr ? ng : p->nr;
    if (p->ninv < p->s0) p->ninv = p->s0;
    // ---- lane kernel: replaces the angle kernel for symmetric GLCMs whose per-thread workspace fits
    p->l_nap = na <= 1 ? 1 : (na <= 2 ? 2 : 4);
    p->l_doubles = ng * (ng + 1) / 2 + 2 * ng;
    if (p->l_doubles < (p->nr + 1) / 2 + 1) p->l_doubles = (p->nr + 1) / 2 + 1;
    p->l_smem_total = p->l_doubles * 8 * RADB_LSTRIDE;
    // (>= 3 resident CTAs per SM: with fewer the serial per-thread chains are latency-bound and the
    // warp-per-angle kernel wins -- measured at Ng 26: 1.55 ms vs 1.16 ms per 8192 patches)
#ifdef RADB_FORCE_G8   // A/B build: the 8-lanes-per-angle MCC kernel also below 15 gray levels
    p->use_lane = 0;
#else
    p->use_lane = (p->symmetric && p->l_smem_total <= 72 * 1024) ? 1 : 0;
#endif
    if (!p->use_lane && p->symmetric && !big && ng <= 40) {
        // mid-size matrices: the per-thread scratch only holds the marginals (2 * ng slots); the eigenproblems go
        // to the MCC kernel, one warp per patch
        p->use_lane = 2;
        p->l_doubles = 2 * ng;
        if (p->l_doubles < (p->nr + 1) / 2 + 1) p->l_doubles = (p->nr + 1) / 2 + 1;
        p->l_smem_total = p->l_doubles * 8 * RADB_LSTRIDE;
        if (p->l_smem_total > 72 * 1024) p->use_lane = 0;
    }
    // Lanczos MCC kernel: CSR copy of one angle's non-zeros (<= min(ng^2, 2 * #voxel pairs) entries, counts < 2^16)
    {
        long long cap = (long long)ng * ng, pairs = 2LL * p->HW;
        if (cap > pairs) cap = pairs;
        int o = 0;
        p->z_rowptr = o; o += radb_align((ng + 2) * 4, 16);
        p->z_ent = o; o += radb_align((int)(cap < (1 << 20) ? cap : (1 << 20)) * 4, 16);
        p->z_vec = o; o += 5 * ng * 8;
        p->z_tri = o; o += (2 * 448 + 16) * 8 + 64;  // al, be2 [RADB_LZ_KMAX], slots, shared scalars (kernel: al + RADB_LZ_KMAX)
        p->z_cap = (int)(cap < (1 << 20) ? cap : (1 << 20));
        p->z_smem_total = o;
        p->use_lanczos = (!p->use_lane && p->symmetric && ng <= 256 && p->HW <= 32767 && o <= 200 * 1024 && !radb_no_lanczos()) ? 1 : 0;
    }
    p->mcc_stride = ng * (ng + 1) / 2 + 4 * ng;  // dense MCC workspace of the warp-per-angle kernel (doubles per angle)
    if (p->use_lanczos) p->mcc_stride = 0;       // the Lanczos kernel solves the eigenproblems: only the marginal tables remain
    if (p->mcc_stride < 2 * ng + 8) p->mcc_stride = 2 * ng + 8;
    int o = 0;
    // ---- build kernel
    p->o_stage = o;                       // raw pixels (TMA destination); later union-find words (u32[HW])
    p->o_mask = o;
    if (!wide) {
        int stage_bytes = radb_align(p->HW * pix_bytes, 16);
        p->o_mask = o + stage_bytes;      // raw mask (TMA destination)
        int both = stage_bytes + radb_align(p->HW, 16);
        int lab_bytes = radb_align(H * p->lp * 4, 16);   // union-find words (parent | size << 16), pitch lp
        o += both > lab_bytes ? both : lab_bytes;
    }
    p->o_mbar = o; o += 16;
    p->o_zero = o;                        // everything from here on is zeroed at CTA start
    p->o_lev = o;
    if (!wide) o += radb_align((H + 2) * p->WP * p->lev_bytes, 16);
    p->o_uq = o; o += (RADB_NTB / 32) * p->uq_cap * (wide ? 8 : 4);   // per-warp union request queues
    p->o_lut = o; o += 256;
    p->o_fo = o; o += pix_bytes == 1 ? 16 : radb_align(64 * 8 + 16 * 8 + 10 * 8 + 10 * 8 + 10 * 4 + 10 * 4 + 8 + 10 * 256 * 4 + 64, 16);  // RADB_FO_SCRATCH
    p->o_rec = o;
    p->o_misc = o; o += 32 * 4;           // record header: [0] Np, [5] #overflow zones, [8] Ng, [9] #levels present, [10+a] longest run of angle a
    p->o_hist = o; o += 256 * 4;
    p->o_lhist = o; o += radb_align(ng * 4, 16);
    p->o_glcm = o; if (!big) o += radb_align(na * (ng + 1) * (ng + 1) * 4, 16);  // padded while it is built
    // GLDM / NGTDM: a few garbage words in front of each, hit by the (branch-free) updates of the non-ROI pixels
    // of a 4-pixel word (level 0 -> row -1)
    o += radb_align((2 * na + 1) * 4, 16);
    p->o_gldm = o; o += radb_align(ng * (2 * na + 1) * 4, 16);
    o += radb_align((2 * na + 1) * 4, 16);
    p->o_ngc = o; o += radb_align(ng * 2 * na * 4, 16);   // [ng][2na] voxel counts per neighbour count
    o += radb_align((2 * na + 1) * 4, 16);
    p->o_ngn = o; o += radb_align(ng * 2 * na * 4, 16);   // [ng][2na] sum |cnt*i - sum(neigh)|
    p->o_szm = o; o += radb_align(ng * p->s0 * 4, 16);
    p->rec_copy_bytes = o - p->o_rec;
    p->glrlm_stride = radb_align(ng * p->nrp * (wide ? 4 : 2), 16);
    p->o_ovf = o; o += radb_align(p->ovf_cap * 4, 16);     // wide: lives in the global record only
    // GLRLM, last in the record.  Record: packed u16 [na][ng][nrp] (u32 in wide mode).  Narrow mode builds it DENSE in
    // shared memory when it can -- u32 [na][ng][16] for run lengths <= 16, incremented by 1 (ATOMS.POPC.INC: a fifth of
    // the shared-memory wavefronts of an add of 1 << 16 to a packed pair) -- and sends the rare longer runs straight to
    // the record with global atomics; the dense part is packed into the record at the end.
    p->glrlm_dense = (!wide && p->nrp % 8 == 0 && p->nrp >= 16) ? 16 : 0;
    p->o_glrlm = o;
    const int smem_end = o + (p->glrlm_dense ? radb_align(na * ng * p->glrlm_dense * 4, 16) : na * p->glrlm_stride);
    o += na * p->glrlm_stride;                             // wide: lives in the global record only
    if (big) { p->o_glcm = o; o += radb_align(na * ng * ng * 4, 16); }  // big: GLCM in the global record only
    p->rec_bytes = o - p->o_rec;
    p->smem_total = wide ? p->o_rec + p->rec_copy_bytes : smem_end;
    if (!wide) p->rec_copy_bytes = p->glrlm_dense ? p->o_glrlm - p->o_rec : p->rec_bytes;
    // Run list: the along-row walk appends the start pixel of every run, and the zone phases (fold run lengths
    // into roots, emit roots) visit the ~HW/5 runs instead of scanning the bounding box twice.  Kept only when
    // its 2 * HW bytes do not cost a resident CTA (5 per SM at most: the register cap of the build kernel).
    p->o_runs = -1;
    if (!wide && p->HW <= 65535) {
        const int sm = 227 * 1024, with = p->smem_total + radb_align(p->HW * 2, 16);
        int before = sm / (p->smem_total + 1024), after = sm / (with + 1024);
        if (before > 5) before = 5;
        if (after >= before && after >= 1) {
            p->o_runs = p->smem_total;
            p->smem_total = with;
        }
    }
    p->g_lev = 0;
    p->g_lab = radb_align((H + 2) * p->WP * p->lev_bytes, 16);
    p->scr_bytes = wide ? (p->g_lab + (long long)p->HW * 8 + 15) / 16 * 16 : 0;  // 16-byte multiple: uint4 stores
    p->g_mcc = 0;  // big: the MCC workspaces re-use the scratch once the build kernel is done with it
    if (big && p->scr_bytes < (long long)na * p->mcc_stride * 8) p->scr_bytes = (long long)na * p->mcc_stride * 8;
    o = 0;
    p->g8_px = o; o += radb_align(ng * 4, 16);
    p->g8_idx = o; o += radb_align(ng, 16);
    p->g8_mcc = o; o += radb_align((ng * (ng + 1) / 2 + 2 * ng) * 8, 16);  // M + (d | v) + (e2 | w), see mcc_task_g8
    p->g8_group_bytes = o;
    p->g8_smem_total = (RADB_NTM / 32) * 4 * p->g8_group_bytes;
    // ---- angle kernel
    o = 0;
    p->a_px = o; o += radb_align(ng * 4, 16);
    p->a_py = o; o += radb_align(ng * 4, 16);
    p->a_padd = o; o += radb_align(2 * ng * 4, 16);
    p->a_psub = o; o += radb_align(ng * 4, 16);
    p->a_pr = o; o += radb_align(p->nr * 4, 16);
    p->a_idx = o; o += radb_align(ng, 16);
    p->a_mcc = o; if (!big) o += radb_align(p->mcc_stride * 8, 16);
    p->a_red = o; o += 14 * 33 * 8;       // RADB_RED_DOUBLES
    p->a_warp_bytes = o;
    o = (RADB_NT / 32) * p->a_warp_bytes;
    p->a_fsc = o; o += RADB_MAX_ANGLES * RADB_FSC_STRIDE * 8;
    p->a_valid = o; o += 16 * 4;
    p->a_smem_total = o;
    p->ml_doubles = 2 * ng + 16 > 32 ? 2 * ng + 16 : 32;  // >= RADB_LANE_MAX_OVF / 2 slots for the sorted overflow list
    p->ml_smem_total = 4 * p->ml_doubles * 32 * 8;
    // ---- misc kernel
    o = 0;
    p->m_pg = o; o += radb_align(ng * 4, 16);
    p->m_ovf2 = o; o += radb_align(p->ovf_cap * 4, 16);
    p->m_ngp = o; o += radb_align(2 * ng * 8, 16);
    p->m_qv = o; o += 16 * 8;
    p->m_red = o; o += (RADB_NT / 32) * 14 * 33 * 8;
    p->m_smem_total = o;
    // ---- shape kernel: row extremes of the contour vertices + integer totals
    p->s_smem_total = radb_align(2 * (2 * H + 1) * 4, 16) + 10 * 8 + (RADB_NT / 32) * 2 * 4 + 16;
}
