// Host-side helpers shared by the C-ABI (radb_api.cu) and the emulation harness (tests/emu):
// settings validation, feature-name table (SURVEY.md A.2 order), launch-parameter setup.
#pragma once
#include <math.h>
#include <string>
#include <vector>
#include "../../include/radb.h"
#include "radb_params.h"

static_assert(sizeof(radb_settings) == 72, "radb_settings layout is part of the ABI (ctypes mirror in _abi.py)");

namespace radb {

struct ClassInfo {
    unsigned bit;
    const char* name;
    std::vector<const char*> feats;
};

// A.2: alphabetical get<Name>FeatureValue order, deprecated features excluded.
static inline const std::vector<ClassInfo>& classes()
{
    static const std::vector<ClassInfo> c = {
        {RADB_CLASS_FIRSTORDER, "firstorder",
         {"10Percentile", "90Percentile", "Energy", "Entropy", "InterquartileRange", "Kurtosis", "Maximum",
          "MeanAbsoluteDeviation", "Mean", "Median", "Minimum", "Range", "RobustMeanAbsoluteDeviation",
          "RootMeanSquared", "Skewness", "TotalEnergy", "Uniformity", "Variance"}},
        {RADB_CLASS_GLCM, "glcm",
         {"Autocorrelation", "ClusterProminence", "ClusterShade", "ClusterTendency", "Contrast", "Correlation",
          "DifferenceAverage", "DifferenceEntropy", "DifferenceVariance", "Id", "Idm", "Idmn", "Idn", "Imc1",
          "Imc2", "InverseVariance", "JointAverage", "JointEnergy", "JointEntropy", "MCC",
          "MaximumProbability", "SumAverage", "SumEntropy", "SumSquares"}},
        {RADB_CLASS_GLDM, "gldm",
         {"DependenceEntropy", "DependenceNonUniformity", "DependenceNonUniformityNormalized",
          "DependenceVariance", "GrayLevelNonUniformity", "GrayLevelVariance", "HighGrayLevelEmphasis",
          "LargeDependenceEmphasis", "LargeDependenceHighGrayLevelEmphasis",
          "LargeDependenceLowGrayLevelEmphasis", "LowGrayLevelEmphasis", "SmallDependenceEmphasis",
          "SmallDependenceHighGrayLevelEmphasis", "SmallDependenceLowGrayLevelEmphasis"}},
        {RADB_CLASS_GLRLM, "glrlm",
         {"GrayLevelNonUniformity", "GrayLevelNonUniformityNormalized", "GrayLevelVariance",
          "HighGrayLevelRunEmphasis", "LongRunEmphasis", "LongRunHighGrayLevelEmphasis",
          "LongRunLowGrayLevelEmphasis", "LowGrayLevelRunEmphasis", "RunEntropy", "RunLengthNonUniformity",
          "RunLengthNonUniformityNormalized", "RunPercentage", "RunVariance", "ShortRunEmphasis",
          "ShortRunHighGrayLevelEmphasis", "ShortRunLowGrayLevelEmphasis"}},
        {RADB_CLASS_GLSZM, "glszm",
         {"GrayLevelNonUniformity", "GrayLevelNonUniformityNormalized", "GrayLevelVariance",
          "HighGrayLevelZoneEmphasis", "LargeAreaEmphasis", "LargeAreaHighGrayLevelEmphasis",
          "LargeAreaLowGrayLevelEmphasis", "LowGrayLevelZoneEmphasis", "SizeZoneNonUniformity",
          "SizeZoneNonUniformityNormalized", "SmallAreaEmphasis", "SmallAreaHighGrayLevelEmphasis",
          "SmallAreaLowGrayLevelEmphasis", "ZoneEntropy", "ZonePercentage", "ZoneVariance"}},
        {RADB_CLASS_NGTDM, "ngtdm", {"Busyness", "Coarseness", "Complexity", "Contrast", "Strength"}},
    };
    return c;
}

struct Plan {
    radb_settings s;
    int max_ng;
    int F;
    int off[6];
    int off_shape;
    std::vector<std::string> names;
};

// Validates settings; returns 0 or a RADB_E_* code with a message.
static inline int make_plan(const radb_settings& s, Plan& pl, std::string& err)
{
    if (s.bin_count < 0) { err = "bin_count must be >= 0"; return RADB_E_INVALID; }
    if (s.bin_count == 0 && (!(s.bin_width > 0) || !isfinite(s.bin_width))) { err = "bin_width must be > 0"; return RADB_E_INVALID; }
    if (s.n_angles < 1 || s.n_angles > RADB_MAX_ANGLES) { err = "n_angles must be 1..4 (distance-1 offsets in a plane)"; return RADB_E_INVALID; }
    for (int a = 0; a < s.n_angles; a++) {
        int dy = s.angles[a][0], dx = s.angles[a][1];
        if (dy < -1 || dy > 1 || dx < -1 || dx > 1 || (dy == 0 && dx == 0)) {
            err = "angle offsets must have infinity-norm 1 (distances other than [1] are not implemented)";
            return RADB_E_UNSUPPORTED;
        }
        for (int b = 0; b < a; b++)
            if ((s.angles[b][0] == dy && s.angles[b][1] == dx) || (s.angles[b][0] == -dy && s.angles[b][1] == -dx)) {
                err = "duplicate angle";
                return RADB_E_INVALID;
            }
    }
    if ((s.class_mask & RADB_CLASS_ALL) == 0) { err = "no feature class enabled"; return RADB_E_INVALID; }
    if (s.gldm_alpha < 0) { err = "gldm_alpha must be >= 0"; return RADB_E_INVALID; }
    pl.s = s;
    if (s.n_angles == 4) {
        // four distinct distance-1 offsets up to sign = the in-plane set: bring them into the canonical
        // order (1,1), (0,1), (-1,1), (1,0) (pyradiomics' generator order) the kernel's fast path assumes
        static const int canon[4][2] = {{1, 1}, {0, 1}, {-1, 1}, {1, 0}};
        for (int a = 0; a < 4; a++) {
            bool found = false;
            for (int b = 0; b < 4; b++)
                if ((s.angles[b][0] == canon[a][0] && s.angles[b][1] == canon[a][1]) ||
                    (s.angles[b][0] == -canon[a][0] && s.angles[b][1] == -canon[a][1]))
                    found = true;
            if (!found) { err = "four angles must be the in-plane set"; return RADB_E_INVALID; }
            pl.s.angles[a][0] = (int8_t)canon[a][0];
            pl.s.angles[a][1] = (int8_t)canon[a][1];
        }
    }
    int ng = s.max_ng;
    if (s.bin_count > 0) ng = s.bin_count;  // binCount: the top bin always holds the ROI maximum, so Ng = binCount
    if (ng <= 0) ng = (int)floor(255.0 / s.bin_width) + 1;  // uint8 pixels: levels 1..floor(255/bw)+1
    if (ng > 256) { err = "more than 256 gray levels"; return RADB_E_UNSUPPORTED; }
    pl.max_ng = ng;
    pl.F = 0;
    pl.names.clear();
    pl.off_shape = -1;
    if (s.class_mask & RADB_CLASS_SHAPE2D) {  // A.1 step 3: shape keys come first
        static const char* shape_names[9] = {"Elongation", "MajorAxisLength", "MaximumDiameter", "MeshSurface",
                                             "MinorAxisLength", "Perimeter", "PerimeterSurfaceRatio", "PixelSurface",
                                             "Sphericity"};
        pl.off_shape = 0;
        for (auto f : shape_names) pl.names.push_back(std::string("original_shape2D_") + f);
        pl.F = 9;
    }
    int k = 0;
    for (const auto& c : classes()) {
        if (s.class_mask & c.bit) {
            pl.off[k] = pl.F;
            for (auto f : c.feats) pl.names.push_back(std::string("original_") + c.name + "_" + f);
            pl.F += (int)c.feats.size();
        } else
            pl.off[k] = -1;
        k++;
    }
    return 0;
}

static inline int fill_params(const Plan& pl, int H, int W, int dtype, RadbParams& p, std::string& err)
{
    int pix_bytes = 0;
    switch (dtype) {
        case RADB_DTYPE_U8: pix_bytes = 1; break;
        case RADB_DTYPE_U16: pix_bytes = 2; break;
        case RADB_DTYPE_F32: pix_bytes = 4; break;
        case RADB_DTYPE_F64: pix_bytes = 8; break;
        default: err = "unknown pixel dtype"; return RADB_E_INVALID;
    }
    if (dtype != RADB_DTYPE_U8 && pl.s.max_ng <= 0 && pl.s.bin_count <= 0) {
        err = "max_ng must be given for non-uint8 pixels (the gray-level count cannot be bounded from the dtype)";
        return RADB_E_INVALID;
    }
    if (H < 1 || W < 1 || H > 4096 || W > 4096 || (long long)H * W >= (1 << 24)) {
        err = "image must be 1..4096 pixels per side and < 2^24 pixels";
        return RADB_E_INVALID;
    }
    memset(&p, 0, sizeof(p));
    p.H = H;
    p.W = W;
    p.mask_group = 1;
    p.label = pl.s.label;
    p.n_angles = pl.s.n_angles;
    for (int a = 0; a < pl.s.n_angles; a++) { p.ang_y[a] = pl.s.angles[a][0]; p.ang_x[a] = pl.s.angles[a][1]; }
    p.symmetric = pl.s.symmetrical_glcm ? 1 : 0;
    p.alpha = (int)floor(pl.s.gldm_alpha);
    p.bin_width = pl.s.bin_count > 0 ? 1.0 : pl.s.bin_width;
    p.bin_count = pl.s.bin_count;
    p.bw_int = (pl.s.bin_count <= 0 && pl.s.bin_width >= 1.0 && pl.s.bin_width <= 255.0 && pl.s.bin_width == floor(pl.s.bin_width)) ? (int)pl.s.bin_width : 0;
    p.shift = pl.s.voxel_array_shift;
    p.max_ng = pl.max_ng;
    p.F = pl.F;
    p.off_fo = pl.off[0];
    p.off_glcm = pl.off[1];
    p.off_gldm = pl.off[2];
    p.off_glrlm = pl.off[3];
    p.off_glszm = pl.off[4];
    p.off_ngtdm = pl.off[5];
    p.off_shape = pl.off_shape;
    // narrow mode (everything in shared memory) when the patch fits with >= 2 CTAs per SM,
    // otherwise wide mode (level image, union-find words, GLRLM, overflow list in global memory)
    // and big mode (GLCM counters + MCC workspace in global memory as well) when the gray-level count makes
    // even that too large (Ng >~ 90 with 4 angles: the binWidth sweep of BASELINE.json configs[4])
    p.wide = 0;
    p.big = 0;
    radb_layout(&p, pix_bytes);
    if ((long long)H * (W + 1) > 65535 || p.smem_total > 110 * 1024 || pl.max_ng > 255) {
        p.wide = 1;
        radb_layout(&p, pix_bytes);
        if (p.smem_total > 200 * 1024 || p.a_smem_total > 200 * 1024 || pl.max_ng > 255) {
            p.big = 1;
            radb_layout(&p, pix_bytes);
        }
    }
    if (p.lev_bytes == 2 && pix_bytes == 1) { err = "uint8 pixels cannot have more than 255 gray levels"; return RADB_E_INVALID; }
    if (p.smem_total > 227 * 1024 || p.a_smem_total > 227 * 1024 || p.m_smem_total > 227 * 1024 || p.s_smem_total > 227 * 1024) {
        err = "image size x gray levels need more than 227 KB of shared memory";
        return RADB_E_SMEM;
    }
    return 0;
}

// radb_extract_ragged: patches of one (H, W) are launched together.  Groups keep the order in which their
// size first appears and every group keeps its members in input order.
struct RaggedGroup {
    int H, W;
    std::vector<long long> idx;  // input positions
};
static inline int group_ragged(long long n, const int32_t* hw, std::vector<RaggedGroup>& groups, std::string& err)
{
    groups.clear();
    for (long long i = 0; i < n; i++) {
        const int H = hw[2 * i], W = hw[2 * i + 1];
        if (H < 1 || W < 1) { err = "ragged batch: non-positive patch size"; return RADB_E_INVALID; }
        RaggedGroup* g = nullptr;
        for (auto& e : groups)
            if (e.H == H && e.W == W) { g = &e; break; }
        if (!g) {
            groups.push_back({H, W, {}});
            g = &groups.back();
        }
        g->idx.push_back(i);
    }
    return 0;
}

// ITK RecursiveGaussianImageFilter::SetUp (Deriche's 4th-order recursive Gaussian; spacing 1, NormalizeAcrossScale
// on): the coefficient sets of the zero-order (order = 0) and second-order (order = 2) filters.  Same formulas, in
// the same order, as oracle/image_filters.py:_deriche_coefficients.  out: N[4], D[4], M[4], BN[4], BM[4].
static inline void deriche_coefficients(double sigma, int order, double* N, double* D, double* M, double* BN, double* BM)
{
    const double sd = sigma;
    const double W1 = 0.6681, L1 = -1.3932, W2 = 2.0787, L2 = -1.3732;
    const double A1[3] = {1.3530, -0.6724, -1.3563}, B1[3] = {1.8151, -3.4327, 5.2318};
    const double A2[3] = {-0.3531, 0.6724, 0.3446}, B2[3] = {0.0902, 0.6100, -2.2355};
    const double c1 = cos(W1 / sd), c2 = cos(W2 / sd), s1 = sin(W1 / sd), s2 = sin(W2 / sd);
    const double e1 = exp(L1 / sd), e2 = exp(L2 / sd);
    auto ncoef = [&](double a1, double b1, double a2, double b2, double* n, double& SN, double& DN, double& EN) {
        n[0] = a1 + a2;
        n[1] = e2 * (b2 * s2 - (a2 + 2 * a1) * c2);
        n[1] += e1 * (b1 * s1 - (a1 + 2 * a2) * c1);
        n[2] = (a1 + a2) * c2 * c1;
        n[2] -= b1 * c2 * s1 + b2 * c1 * s2;
        n[2] *= 2 * e1 * e2;
        n[2] += a2 * e1 * e1 + a1 * e2 * e2;
        n[3] = e2 * e1 * e1 * (b2 * s2 - a2 * c2);
        n[3] += e1 * e2 * e2 * (b1 * s1 - a1 * c1);
        SN = n[0] + n[1] + n[2] + n[3];
        DN = n[1] + 2 * n[2] + 3 * n[3];
        EN = n[1] + 4 * n[2] + 9 * n[3];
    };
    D[3] = e1 * e1 * e2 * e2;
    D[2] = -2 * c1 * e1 * e2 * e2;
    D[2] += -2 * c2 * e2 * e1 * e1;
    D[1] = 4 * c2 * c1 * e1 * e2;
    D[1] += e1 * e1 + e2 * e2;
    D[0] = -2 * (e2 * c2 + e1 * c1);
    const double SD = 1.0 + D[0] + D[1] + D[2] + D[3];
    const double DD = D[0] + 2 * D[1] + 3 * D[2] + 4 * D[3];
    const double ED = D[0] + 4 * D[1] + 9 * D[2] + 16 * D[3];
    if (order == 0) {
        double n[4], SN, DN, EN;
        ncoef(A1[0], B1[0], A2[0], B2[0], n, SN, DN, EN);
        const double alpha0 = 2 * SN / SD - n[0];
        for (int k = 0; k < 4; k++) N[k] = n[k] / alpha0;
    } else {
        const double scale = sigma * sigma;
        double a[4], b[4], SN0, DN0, EN0, SN2, DN2, EN2;
        ncoef(A1[0], B1[0], A2[0], B2[0], a, SN0, DN0, EN0);
        ncoef(A1[2], B1[2], A2[2], B2[2], b, SN2, DN2, EN2);
        const double beta = -(2 * SN2 - SD * b[0]) / (2 * SN0 - SD * a[0]);
        const double SN = SN2 + beta * SN0, DN = DN2 + beta * DN0, EN = EN2 + beta * EN0;
        const double alpha2 = (EN * SD * SD - ED * SN * SD - 2 * DN * DD * SD + 2 * DD * DD * SN) / (SD * SD * SD);
        for (int k = 0; k < 4; k++) N[k] = (b[k] + beta * a[k]) * (scale / alpha2);
    }
    M[0] = N[1] - D[0] * N[0];
    M[1] = N[2] - D[1] * N[0];
    M[2] = N[3] - D[2] * N[0];
    M[3] = -D[3] * N[0];
    const double SNs = N[0] + N[1] + N[2] + N[3], SMs = M[0] + M[1] + M[2] + M[3], SDs = 1.0 + D[0] + D[1] + D[2] + D[3];
    for (int k = 0; k < 4; k++) { BN[k] = D[k] * SNs / SDs; BM[k] = D[k] * SMs / SDs; }
}

// 1/k^2 and log2(k) tables read by the reduction kernels (device copies are made at radb_create)
static inline void make_tables(int ninv, std::vector<double>& inv2, std::vector<double>& tlog)
{
    inv2.resize(ninv);
    for (int k = 0; k < ninv; k++) inv2[k] = 1.0 / ((double)(k + 1) * (double)(k + 1));
    tlog.resize(2048);  // RADB_TLOG_N
    for (int k = 0; k < 2048; k++) tlog[k] = k >= 2 ? log2((double)k) : 0.0;
}

}  // namespace radb
