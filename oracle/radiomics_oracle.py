"""NumPy restatement of the pyradiomics 3.1.0 feature path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see ``oracle/__init__.py``): pyradiomics is the un-vendored
third-party engine behind ``/root/reference/RadiomicExtractor.py:8,15,38-48``
(version pin: ``/root/reference/params.yml:24``).  Every function below cites
the upstream pyradiomics module it restates (SURVEY.md Appendix A section in
brackets) and the reference call site that reaches it.

Conventions: images are 2-D ``(y, x)`` arrays, exactly what
``sitk.GetImageFromArray(im_gray)`` hands to pyradiomics at
``RadiomicExtractor.py:31``; all arithmetic is int64 / float64 (numpy-1
promotion semantics, SURVEY.md A.3 pitfall).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

EPS = float(np.spacing(1))  # pyradiomics' ``eps = numpy.spacing(1)``

CLASS_ORDER = ("firstorder", "glcm", "gldm", "glrlm", "glszm", "ngtdm")
SHAPE2D_NAMES = ["Elongation", "MajorAxisLength", "MaximumDiameter", "MeshSurface", "MinorAxisLength", "Perimeter",
                 "PerimeterSurfaceRatio", "PixelSurface", "Sphericity"]  # [A.2] (SphericalDisproportion is deprecated)

# [A.2] feature names, alphabetical by getXxxFeatureValue (inspect.getmembers order),
# deprecated features excluded (they are off when the class list is ``[]``,
# params.yml:164-171).
FEATURE_NAMES = {
    "firstorder": ["10Percentile", "90Percentile", "Energy", "Entropy", "InterquartileRange",
                   "Kurtosis", "Maximum", "MeanAbsoluteDeviation", "Mean", "Median", "Minimum",
                   "Range", "RobustMeanAbsoluteDeviation", "RootMeanSquared", "Skewness",
                   "TotalEnergy", "Uniformity", "Variance"],
    "glcm": ["Autocorrelation", "ClusterProminence", "ClusterShade", "ClusterTendency", "Contrast",
             "Correlation", "DifferenceAverage", "DifferenceEntropy", "DifferenceVariance", "Id",
             "Idm", "Idmn", "Idn", "Imc1", "Imc2", "InverseVariance", "JointAverage", "JointEnergy",
             "JointEntropy", "MCC", "MaximumProbability", "SumAverage", "SumEntropy", "SumSquares"],
    "gldm": ["DependenceEntropy", "DependenceNonUniformity", "DependenceNonUniformityNormalized",
             "DependenceVariance", "GrayLevelNonUniformity", "GrayLevelVariance",
             "HighGrayLevelEmphasis", "LargeDependenceEmphasis",
             "LargeDependenceHighGrayLevelEmphasis", "LargeDependenceLowGrayLevelEmphasis",
             "LowGrayLevelEmphasis", "SmallDependenceEmphasis",
             "SmallDependenceHighGrayLevelEmphasis", "SmallDependenceLowGrayLevelEmphasis"],
    "glrlm": ["GrayLevelNonUniformity", "GrayLevelNonUniformityNormalized", "GrayLevelVariance",
              "HighGrayLevelRunEmphasis", "LongRunEmphasis", "LongRunHighGrayLevelEmphasis",
              "LongRunLowGrayLevelEmphasis", "LowGrayLevelRunEmphasis", "RunEntropy",
              "RunLengthNonUniformity", "RunLengthNonUniformityNormalized", "RunPercentage",
              "RunVariance", "ShortRunEmphasis", "ShortRunHighGrayLevelEmphasis",
              "ShortRunLowGrayLevelEmphasis"],
    "glszm": ["GrayLevelNonUniformity", "GrayLevelNonUniformityNormalized", "GrayLevelVariance",
              "HighGrayLevelZoneEmphasis", "LargeAreaEmphasis", "LargeAreaHighGrayLevelEmphasis",
              "LargeAreaLowGrayLevelEmphasis", "LowGrayLevelZoneEmphasis", "SizeZoneNonUniformity",
              "SizeZoneNonUniformityNormalized", "SmallAreaEmphasis",
              "SmallAreaHighGrayLevelEmphasis", "SmallAreaLowGrayLevelEmphasis", "ZoneEntropy",
              "ZonePercentage", "ZoneVariance"],
    "ngtdm": ["Busyness", "Coarseness", "Complexity", "Contrast", "Strength"],
}

DEFAULT_SETTINGS = {  # [A.1] pyradiomics defaults
    "label": 1, "binWidth": 25, "binCount": None, "force2D": False, "force2Ddimension": 0,
    "distances": [1], "symmetricalGLCM": True, "weightingNorm": None, "gldm_a": 0,
    "voxelArrayShift": 0, "minimumROIDimensions": 2, "minimumROISize": None,
}


# --------------------------------------------------------------------------- angles
def angles(ndim, distances=(1,), force2D=False, force2Ddimension=0):
    """[A.4] pyradiomics ``src/cmatrices.c:build_angles`` driven by ``_cmatrices.c``.

    Returns ``(unidirectional, bidirectional)`` lists of offset tuples in the
    enumeration order of the C generator (axis 0 varies fastest, offsets run
    +D..-D).  GLCM/GLRLM use the first list, GLSZM/GLDM/NGTDM the second.
    """
    if not force2D:
        force2Ddimension = -1
    D = int(max(distances))
    stride = 2 * D + 1
    n_all = 1
    for d in range(ndim):
        if d != force2Ddimension:
            n_all *= stride
    n_all -= 1  # the all-zero offset is never generated
    dist_set = set(int(d) for d in distances)

    def gen(count):
        out = []
        for a_idx in range(count):
            a_off = 1
            a_dist = 0
            ang = []
            for d in range(ndim):
                if d == force2Ddimension:
                    ang.append(0)
                else:
                    off = D - (a_idx // a_off) % stride
                    ang.append(off)
                    a_off *= stride
                    a_dist = max(a_dist, abs(off))
            if a_dist in dist_set:
                out.append(tuple(ang))
        return out

    uni = gen(n_all // 2)
    bi = uni + [tuple(-c for c in a) for a in uni]
    return uni, bi


# --------------------------------------------------------------------------- mask / binning
def check_mask(mask, label, minimumROIDimensions=2, minimumROISize=None):
    """[A.1 step 2] pyradiomics ``imageoperations.checkMask`` (reached from
    ``RadiomicExtractor.py:38``): bounding box + ROI validity, ValueError on failure."""
    mask_arr = (np.asarray(mask) == label)
    if not mask_arr.any():
        raise ValueError("Label (%s) not present in mask" % (label,))
    idx = np.nonzero(mask_arr)
    bbox = [(int(i.min()), int(i.max())) for i in idx]
    ndims = sum(1 for lo, hi in bbox if hi - lo + 1 > 1)
    if ndims == 0:
        raise ValueError("mask only contains 1 segmented voxel! Cannot extract features for a single voxel.")
    if ndims < minimumROIDimensions:
        raise ValueError("mask has too few dimensions (number of dimensions %d, minimum required %d)"
                         % (ndims, minimumROIDimensions))
    if minimumROISize is not None and int(mask_arr.sum()) <= minimumROISize:
        raise ValueError("Size of the ROI is too small (minimum size: %g)" % minimumROISize)
    return bbox


def get_bin_edges(values, binWidth=25, binCount=None):
    """[A.3] pyradiomics ``imageoperations.getBinEdges`` with numpy-1 promotion
    (values are cast to float64 before any edge arithmetic)."""
    v = np.asarray(values, dtype=np.float64)
    if binCount is not None:
        edges = np.histogram(v, int(binCount))[1]
        edges[-1] += 1
        return edges
    minimum = float(v.min())
    maximum = float(v.max())
    low = minimum - (minimum % binWidth)
    high = maximum + 2 * binWidth
    edges = np.arange(low, high, binWidth)
    if len(edges) == 1:
        edges = np.array([edges[0] - .5, edges[0] + .5])
    return edges


def bin_image(image, mask_arr, binWidth=25, binCount=None):
    """[A.3] ``imageoperations.binImage`` + ``base._applyBinning``: discretised
    image (0 outside ROI), present gray levels, Ng = max level."""
    out = np.zeros(image.shape, dtype=np.int64)
    edges = get_bin_edges(image[mask_arr], binWidth, binCount)
    out[mask_arr] = np.digitize(np.asarray(image[mask_arr], dtype=np.float64), edges)
    gray_levels = np.unique(out[mask_arr])
    return out, gray_levels, int(gray_levels.max()), edges


# --------------------------------------------------------------------------- matrix builders
def _shifted(levels, off):
    """levels at p + off (0 where p + off leaves the array)."""
    out = np.zeros_like(levels)
    src = []
    dst = []
    for n, o in zip(levels.shape, off):
        if o >= 0:
            src.append(slice(o, n))
            dst.append(slice(0, n - o))
        else:
            src.append(slice(0, n + o))
            dst.append(slice(-o, n))
    out[tuple(dst)] = levels[tuple(src)]
    return out


def glcm_matrix(levels, Ng, uni_angles, symmetrical=True):
    """[A.6] ``cmatrices.c:calculate_glcm`` + ``glcm.py:_applyMatrixOptions`` symmetrisation.
    Returns int64 ``[Ng, Ng, Na]`` counts (not normalised)."""
    P = np.zeros((Ng, Ng, len(uni_angles)), dtype=np.int64)
    for a, off in enumerate(uni_angles):
        nb = _shifted(levels, off)
        ok = (levels > 0) & (nb > 0)
        np.add.at(P[:, :, a], (levels[ok] - 1, nb[ok] - 1), 1)
    if symmetrical:
        P = P + P.transpose(1, 0, 2)
    return P


def glrlm_matrix(levels, Ng, uni_angles):
    """[A.7] ``cmatrices.c:calculate_glrlm``: int64 ``[Ng, Nr, Na]``, Nr = max array dim."""
    H, W = levels.shape
    Nr = max(H, W)
    P = np.zeros((Ng, Nr, len(uni_angles)), dtype=np.int64)
    for a, (dy, dx) in enumerate(uni_angles):
        prev = _shifted(levels, (-dy, -dx))
        starts = np.argwhere((levels > 0) & (prev != levels))
        for y, x in starts:
            g = levels[y, x]
            n = 0
            while 0 <= y < H and 0 <= x < W and levels[y, x] == g:
                n += 1
                y += dy
                x += dx
            P[g - 1, n - 1, a] += 1
    return P


def glszm_matrix(levels, Ng, bi_angles):
    """[A.8] ``cmatrices.c:calculate_glszm`` (stack-based region growing over the
    bidirectional angle set): int64 ``[Ng, Ns]``, Ns = ROI voxel count."""
    H, W = levels.shape
    Ns = int((levels > 0).sum())
    P = np.zeros((Ng, max(Ns, 1)), dtype=np.int64)
    seen = levels == 0
    for y0 in range(H):
        for x0 in range(W):
            if seen[y0, x0]:
                continue
            g = levels[y0, x0]
            stack = [(y0, x0)]
            seen[y0, x0] = True
            size = 0
            while stack:
                y, x = stack.pop()
                size += 1
                for dy, dx in bi_angles:
                    yy, xx = y + dy, x + dx
                    if 0 <= yy < H and 0 <= xx < W and not seen[yy, xx] and levels[yy, xx] == g:
                        seen[yy, xx] = True
                        stack.append((yy, xx))
            P[g - 1, size - 1] += 1
    return P


def gldm_matrix(levels, Ng, bi_angles, alpha=0):
    """[A.9] ``cmatrices.c:calculate_gldm``: int64 ``[Ng, Na_bi + 1]``; column = dependence count."""
    P = np.zeros((Ng, len(bi_angles) + 1), dtype=np.int64)
    roi = levels > 0
    dep = np.zeros(levels.shape, dtype=np.int64)
    for off in bi_angles:
        nb = _shifted(levels, off)
        dep += ((nb > 0) & roi & (np.abs(nb - levels) <= alpha))
    np.add.at(P, (levels[roi] - 1, dep[roi]), 1)
    return P


def ngtdm_matrix(levels, Ng, bi_angles):
    """[A.9] ``cmatrices.c:calculate_ngtdm``: ``n_i`` int64 ``[Ng]`` and ``s_i`` float64 ``[Ng]``."""
    roi = levels > 0
    cnt = np.zeros(levels.shape, dtype=np.int64)
    tot = np.zeros(levels.shape, dtype=np.int64)
    for off in bi_angles:
        nb = _shifted(levels, off)
        cnt += (nb > 0)
        tot += nb
    ok = roi & (cnt > 0)
    n = np.zeros(Ng, dtype=np.int64)
    s = np.zeros(Ng, dtype=np.float64)
    g = levels[ok]
    diff = np.abs(g.astype(np.float64) - tot[ok].astype(np.float64) / cnt[ok].astype(np.float64))
    np.add.at(n, g - 1, 1)
    np.add.at(s, g - 1, diff)
    return n, s


# --------------------------------------------------------------------------- feature classes
def firstorder_features(x, levels_roi, voxelArrayShift=0, spacing_prod=1.0):
    """[A.5] pyradiomics ``firstorder.py``; ``x`` = ROI raw values, ``levels_roi`` = their gray levels."""
    x = np.asarray(x, dtype=np.float64)
    N = x.size
    hist = np.bincount(np.asarray(levels_roi, dtype=np.int64))[1:].astype(np.float64)
    p = hist / hist.sum()
    xs = x + voxelArrayShift
    mean = np.mean(x)
    p10, p25, p75, p90 = (np.percentile(x, q) for q in (10, 25, 75, 90))
    m2 = np.mean((x - mean) ** 2)
    m3 = np.mean((x - mean) ** 3)
    m4 = np.mean((x - mean) ** 4)
    inner = x[(x >= p10) & (x <= p90)]
    f = OrderedDict()
    f["10Percentile"] = p10
    f["90Percentile"] = p90
    f["Energy"] = np.sum(xs ** 2)
    f["Entropy"] = -np.sum(p * np.log2(p + EPS))
    f["InterquartileRange"] = p75 - p25
    f["Kurtosis"] = 0.0 if m2 == 0 else m4 / m2 ** 2.0
    f["Maximum"] = np.max(x)
    f["MeanAbsoluteDeviation"] = np.mean(np.abs(x - mean))
    f["Mean"] = mean
    f["Median"] = np.median(x)
    f["Minimum"] = np.min(x)
    f["Range"] = np.max(x) - np.min(x)
    f["RobustMeanAbsoluteDeviation"] = np.mean(np.abs(inner - np.mean(inner)))
    f["RootMeanSquared"] = np.sqrt(np.sum(xs ** 2) / N)
    f["Skewness"] = 0.0 if m2 == 0 else m3 / m2 ** 1.5
    f["TotalEnergy"] = spacing_prod * np.sum(xs ** 2)
    f["Uniformity"] = np.sum(p ** 2)
    f["Variance"] = np.std(x) ** 2
    return f


def _nanmean(v):
    v = np.asarray(v, dtype=np.float64)
    if v.size == 0 or np.all(np.isnan(v)):
        return float("nan")
    return float(np.nanmean(v))


def glcm_features(P_counts, gray_levels, Ng):
    """[A.6] pyradiomics ``glcm.py``; ``P_counts`` int ``[Ng, Ng, Na]`` (already symmetrised)."""
    gl = np.asarray(gray_levels, dtype=np.int64)
    P = P_counts[np.ix_(gl - 1, gl - 1)].astype(np.float64)  # delete absent levels
    sums = P.sum((0, 1))
    if P.shape[2] > 1:
        keep = sums != 0
        P = P[:, :, keep]
        sums = sums[keep]
    sums = sums.copy()
    sums[sums == 0] = np.nan
    P = P / sums
    Na = P.shape[2]
    iv = gl.astype(np.float64)
    i = iv[:, None, None]
    j = iv[None, :, None]
    kSum = np.arange(2, 2 * Ng + 1, dtype=np.float64)
    kDiff = np.arange(0, Ng, dtype=np.float64)
    px = P.sum(1, keepdims=True)
    py = P.sum(0, keepdims=True)
    ux = np.sum(i * P, (0, 1), keepdims=True)
    uy = np.sum(j * P, (0, 1), keepdims=True)
    ipj = (iv[:, None] + iv[None, :])
    imj = np.abs(iv[:, None] - iv[None, :])
    pxAddy = np.array([P[ipj == k, :].sum(0) for k in kSum]).reshape(len(kSum), Na)
    pxSuby = np.array([P[imj == k, :].sum(0) for k in kDiff]).reshape(len(kDiff), Na)
    HXY = -np.sum(P * np.log2(P + EPS), (0, 1))
    f = OrderedDict()
    f["Autocorrelation"] = _nanmean(np.sum(P * (i * j), (0, 1)))
    f["ClusterProminence"] = _nanmean(np.sum(P * ((i + j - ux - uy) ** 4), (0, 1)))
    f["ClusterShade"] = _nanmean(np.sum(P * ((i + j - ux - uy) ** 3), (0, 1)))
    f["ClusterTendency"] = _nanmean(np.sum(P * ((i + j - ux - uy) ** 2), (0, 1)))
    f["Contrast"] = _nanmean(np.sum(P * (np.abs(i - j) ** 2), (0, 1)))
    sigx = np.sum(P * ((i - ux) ** 2), (0, 1), keepdims=True) ** 0.5
    sigy = np.sum(P * ((j - uy) ** 2), (0, 1), keepdims=True) ** 0.5
    corm = np.sum(P * (i - ux) * (j - uy), (0, 1), keepdims=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        corr = corm / (sigx * sigy + EPS)
    corr[sigx * sigy == 0] = 1
    f["Correlation"] = _nanmean(corr)
    diffavg = np.sum(kDiff[:, None] * pxSuby, 0, keepdims=True)
    f["DifferenceAverage"] = _nanmean(diffavg)
    f["DifferenceEntropy"] = _nanmean(-np.sum(pxSuby * np.log2(pxSuby + EPS), 0))
    f["DifferenceVariance"] = _nanmean(np.sum(pxSuby * ((kDiff[:, None] - diffavg) ** 2), 0))
    f["Id"] = _nanmean(np.sum(pxSuby / (1 + kDiff[:, None]), 0))
    f["Idm"] = _nanmean(np.sum(pxSuby / (1 + kDiff[:, None] ** 2), 0))
    f["Idmn"] = _nanmean(np.sum(pxSuby / (1 + (kDiff[:, None] ** 2) / (Ng ** 2)), 0))
    f["Idn"] = _nanmean(np.sum(pxSuby / (1 + kDiff[:, None] / Ng), 0))
    HX = -np.sum(px * np.log2(px + EPS), (0, 1))
    HY = -np.sum(py * np.log2(py + EPS), (0, 1))
    HXY1 = -np.sum(P * np.log2(px * py + EPS), (0, 1))
    div = np.fmax(HX, HY)
    imc1 = HXY - HXY1
    nz = div != 0
    imc1[nz] = imc1[nz] / div[nz]
    imc1[div == 0] = 0
    f["Imc1"] = _nanmean(imc1)
    HXY2 = -np.sum((px * py) * np.log2(px * py + EPS), (0, 1))
    imc2 = 1 - np.e ** (-2 * (HXY2 - HXY))
    imc2[imc2 < 0] = 0  # [A.11 U2] HXY > HXY2 by rounding -> 0
    f["Imc2"] = _nanmean(imc2 ** 0.5)
    f["InverseVariance"] = _nanmean(np.sum(pxSuby[1:, :] / kDiff[1:, None] ** 2, 0))
    f["JointAverage"] = _nanmean(ux)
    f["JointEnergy"] = _nanmean(np.sum(P ** 2, (0, 1)))
    f["JointEntropy"] = _nanmean(HXY)
    if P.shape[0] < 2:
        f["MCC"] = 1.0
    else:
        mcc = []
        for a in range(Na):
            Pa = P[:, :, a]
            if np.isnan(Pa).any():
                mcc.append(np.nan)
                continue
            pxa = Pa.sum(1)
            pya = Pa.sum(0)
            Q = np.zeros((Pa.shape[0], Pa.shape[0]))
            for k in range(Pa.shape[1]):
                Q += (Pa[:, None, k] * Pa[None, :, k]) / (pxa[:, None] * pya[k] + EPS)
            ev = np.linalg.eigvals(Q)
            ev = np.sort(ev)
            mcc.append(np.sqrt(ev[-2]).real)
        f["MCC"] = _nanmean(mcc)
    f["MaximumProbability"] = _nanmean(np.amax(P, (0, 1)))
    f["SumAverage"] = _nanmean(np.sum(kSum[:, None] * pxAddy, 0))
    f["SumEntropy"] = _nanmean(-np.sum(pxAddy * np.log2(pxAddy + EPS), 0))
    f["SumSquares"] = _nanmean(np.sum(P * ((i - ux) ** 2), (0, 1)))
    return f


def glrlm_features(P_counts, gray_levels):
    """[A.7] pyradiomics ``glrlm.py``; ``P_counts`` int ``[Ng, Nr, Na]``."""
    gl = np.asarray(gray_levels, dtype=np.int64)
    P = P_counts[gl - 1].astype(np.float64)
    Nr = P.sum((0, 1))
    if P.shape[2] > 1:
        keep = Nr != 0
        P = P[:, :, keep]
        Nr = Nr[keep]
    Nr = Nr.copy()
    Nr[Nr == 0] = np.nan
    pr = P.sum(0)
    pg = P.sum(1)
    iv = gl.astype(np.float64)
    jv = np.arange(1, P.shape[1] + 1, dtype=np.float64)
    keepj = pr.sum(1) != 0
    P = P[:, keepj, :]
    jv = jv[keepj]
    pr = pr[keepj]
    i2 = (iv ** 2)[:, None, None]
    j2 = (jv ** 2)[None, :, None]
    f = OrderedDict()
    f["GrayLevelNonUniformity"] = _nanmean(np.sum(pg ** 2, 0) / Nr)
    f["GrayLevelNonUniformityNormalized"] = _nanmean(np.sum(pg ** 2, 0) / Nr ** 2)
    pgn = pg / Nr
    u_i = np.sum(pgn * iv[:, None], 0, keepdims=True)
    f["GrayLevelVariance"] = _nanmean(np.sum(pgn * (iv[:, None] - u_i) ** 2, 0))
    f["HighGrayLevelRunEmphasis"] = _nanmean(np.sum(pg * (iv ** 2)[:, None], 0) / Nr)
    f["LongRunEmphasis"] = _nanmean(np.sum(pr * (jv ** 2)[:, None], 0) / Nr)
    f["LongRunHighGrayLevelEmphasis"] = _nanmean(np.sum(P * i2 * j2, (0, 1)) / Nr)
    f["LongRunLowGrayLevelEmphasis"] = _nanmean(np.sum(P * j2 / i2, (0, 1)) / Nr)
    f["LowGrayLevelRunEmphasis"] = _nanmean(np.sum(pg / (iv ** 2)[:, None], 0) / Nr)
    p = P / Nr
    f["RunEntropy"] = _nanmean(-np.sum(p * np.log2(p + EPS), (0, 1)))
    f["RunLengthNonUniformity"] = _nanmean(np.sum(pr ** 2, 0) / Nr)
    f["RunLengthNonUniformityNormalized"] = _nanmean(np.sum(pr ** 2, 0) / Nr ** 2)
    Np = np.sum(pr * jv[:, None], 0)
    f["RunPercentage"] = _nanmean(Nr / Np)
    prn = pr / Nr
    u_j = np.sum(prn * jv[:, None], 0, keepdims=True)
    f["RunVariance"] = _nanmean(np.sum(prn * (jv[:, None] - u_j) ** 2, 0))
    f["ShortRunEmphasis"] = _nanmean(np.sum(pr / (jv ** 2)[:, None], 0) / Nr)
    f["ShortRunHighGrayLevelEmphasis"] = _nanmean(np.sum(P * i2 / j2, (0, 1)) / Nr)
    f["ShortRunLowGrayLevelEmphasis"] = _nanmean(np.sum(P / (i2 * j2), (0, 1)) / Nr)
    return f


def _zone_like_features(P_counts, gray_levels):
    """Shared sums of ``glszm.py`` / ``gldm.py`` (same algebra on an ``[Ng, Nj]`` matrix)."""
    gl = np.asarray(gray_levels, dtype=np.int64)
    P = P_counts[gl - 1].astype(np.float64)
    pj = P.sum(0)
    pg = P.sum(1)
    iv = gl.astype(np.float64)
    jv = np.arange(1, P.shape[1] + 1, dtype=np.float64)
    Nz = P.sum()
    if Nz == 0:
        Nz = 1.0
    Np = np.sum(pj * jv)
    if Np == 0:
        Np = 1.0
    keep = pj != 0
    P = P[:, keep]
    jv = jv[keep]
    pj = pj[keep]
    i2 = (iv ** 2)[:, None]
    j2 = (jv ** 2)[None, :]
    r = dict(P=P, pj=pj, pg=pg, iv=iv, jv=jv, Nz=Nz, Np=Np, i2=i2, j2=j2)
    r["small"] = np.sum(pj / jv ** 2) / Nz
    r["large"] = np.sum(pj * jv ** 2) / Nz
    r["gln"] = np.sum(pg ** 2) / Nz
    r["glnn"] = np.sum(pg ** 2) / Nz ** 2
    r["jn"] = np.sum(pj ** 2) / Nz
    r["jnn"] = np.sum(pj ** 2) / Nz ** 2
    pgn = pg / Nz
    u_i = np.sum(pgn * iv)
    r["glv"] = np.sum(pgn * (iv - u_i) ** 2)
    pjn = pj / Nz
    u_j = np.sum(pjn * jv)
    r["jv_var"] = np.sum(pjn * (jv - u_j) ** 2)
    p = P / Nz
    r["entropy"] = -np.sum(p * np.log2(p + EPS))
    r["lgl"] = np.sum(pg / iv ** 2) / Nz
    r["hgl"] = np.sum(pg * iv ** 2) / Nz
    r["small_lgl"] = np.sum(P / (i2 * j2)) / Nz
    r["small_hgl"] = np.sum(P * i2 / j2) / Nz
    r["large_lgl"] = np.sum(P * j2 / i2) / Nz
    r["large_hgl"] = np.sum(P * i2 * j2) / Nz
    return r


def glszm_features(P_counts, gray_levels):
    """[A.8] pyradiomics ``glszm.py``; ``P_counts`` int ``[Ng, Ns]``."""
    r = _zone_like_features(P_counts, gray_levels)
    f = OrderedDict()
    f["GrayLevelNonUniformity"] = r["gln"]
    f["GrayLevelNonUniformityNormalized"] = r["glnn"]
    f["GrayLevelVariance"] = r["glv"]
    f["HighGrayLevelZoneEmphasis"] = r["hgl"]
    f["LargeAreaEmphasis"] = r["large"]
    f["LargeAreaHighGrayLevelEmphasis"] = r["large_hgl"]
    f["LargeAreaLowGrayLevelEmphasis"] = r["large_lgl"]
    f["LowGrayLevelZoneEmphasis"] = r["lgl"]
    f["SizeZoneNonUniformity"] = r["jn"]
    f["SizeZoneNonUniformityNormalized"] = r["jnn"]
    f["SmallAreaEmphasis"] = r["small"]
    f["SmallAreaHighGrayLevelEmphasis"] = r["small_hgl"]
    f["SmallAreaLowGrayLevelEmphasis"] = r["small_lgl"]
    f["ZoneEntropy"] = r["entropy"]
    f["ZonePercentage"] = r["Nz"] / r["Np"]
    f["ZoneVariance"] = r["jv_var"]
    return f


def gldm_features(P_counts, gray_levels):
    """[A.9] pyradiomics ``gldm.py``; ``P_counts`` int ``[Ng, Na_bi+1]``."""
    r = _zone_like_features(P_counts, gray_levels)
    f = OrderedDict()
    f["DependenceEntropy"] = r["entropy"]
    f["DependenceNonUniformity"] = r["jn"]
    f["DependenceNonUniformityNormalized"] = r["jnn"]
    f["DependenceVariance"] = r["jv_var"]
    f["GrayLevelNonUniformity"] = r["gln"]
    f["GrayLevelVariance"] = r["glv"]
    f["HighGrayLevelEmphasis"] = r["hgl"]
    f["LargeDependenceEmphasis"] = r["large"]
    f["LargeDependenceHighGrayLevelEmphasis"] = r["large_hgl"]
    f["LargeDependenceLowGrayLevelEmphasis"] = r["large_lgl"]
    f["LowGrayLevelEmphasis"] = r["lgl"]
    f["SmallDependenceEmphasis"] = r["small"]
    f["SmallDependenceHighGrayLevelEmphasis"] = r["small_hgl"]
    f["SmallDependenceLowGrayLevelEmphasis"] = r["small_lgl"]
    return f


def ngtdm_features(n, s):
    """[A.9] pyradiomics ``ngtdm.py``; ``n`` int ``[Ng]``, ``s`` float ``[Ng]`` (level = index+1)."""
    n = np.asarray(n, dtype=np.float64)
    s = np.asarray(s, dtype=np.float64)
    keep = n != 0
    iv = (np.arange(len(n), dtype=np.float64) + 1)[keep]
    n = n[keep]
    s = s[keep]
    f = OrderedDict()
    Nvp = n.sum()
    if Nvp == 0:  # no voxel has a valid neighbour: upstream divides 0/0
        for k in FEATURE_NAMES["ngtdm"]:
            f[k] = float("nan")
        return f
    p = n / Nvp
    Ngp = len(n)
    sum_ps = np.sum(p * s)
    ipi = iv * p
    absdiff = np.sum(np.abs(ipi[:, None] - ipi[None, :]))
    f["Busyness"] = sum_ps / absdiff if absdiff != 0 else 0.0
    f["Coarseness"] = 1.0 / sum_ps if sum_ps != 0 else 1e6
    num = (p * s)[:, None] + (p * s)[None, :]
    den = p[:, None] + p[None, :]
    f["Complexity"] = np.sum(np.abs(iv[:, None] - iv[None, :]) * num / den) / Nvp
    div = Ngp * (Ngp - 1)
    c = np.sum(p[:, None] * p[None, :] * (iv[:, None] - iv[None, :]) ** 2) * np.sum(s) / Nvp
    f["Contrast"] = c / div if div != 0 else 0.0
    st = np.sum((p[:, None] + p[None, :]) * (iv[:, None] - iv[None, :]) ** 2)
    sum_s = np.sum(s)
    f["Strength"] = st / sum_s if sum_s != 0 else 0.0
    return f


# --------------------------------------------------------------------------- shape2D
# pyradiomics ``src/cshape.c`` marching-squares tables (corner order, segments per case, edge midpoints)
_GRID_ANGLES_2D = ((0, 0), (0, 1), (1, 1), (1, 0))
_LINE_TABLE_2D = ((), (3, 0), (0, 1), (3, 1), (1, 2), (1, 2, 3, 0), (0, 2), (3, 2), (2, 3), (2, 0), (0, 1, 2, 3),
                  (2, 1), (1, 3), (1, 0), (0, 3), ())
_VERT_LIST_2D = ((0.0, 0.5), (0.5, 1.0), (1.0, 0.5), (0.5, 0.0))


def shape2d_coefficients(mask_arr, spacing=(1.0, 1.0)):
    """pyradiomics ``cshape.c:calculate_coefficients2D`` on the zero-padded mask (shape2D.py pads by 1):
    marching-squares perimeter, mesh surface (|shoelace sum| / 2) and the maximum distance between
    contour vertices.  (SURVEY.md section 8 f rank 1; reached from RadiomicExtractor.py:38 when
    ``shape2D`` is enabled, params.yml:165.)"""
    m = np.pad(np.asarray(mask_arr, dtype=bool), 1)
    H, W = m.shape
    perimeter = 0.0
    surface = 0.0
    verts = []
    for iy in range(H - 1):
        for ix in range(W - 1):
            idx = 0
            for a, (dy, dx) in enumerate(_GRID_ANGLES_2D):
                if m[iy + dy, ix + dx]:
                    idx |= 1 << a
            seg = _LINE_TABLE_2D[idx]
            for t in range(0, len(seg), 2):
                a = [(iy + _VERT_LIST_2D[seg[t]][0]) * spacing[0], (ix + _VERT_LIST_2D[seg[t]][1]) * spacing[1]]
                b = [(iy + _VERT_LIST_2D[seg[t + 1]][0]) * spacing[0], (ix + _VERT_LIST_2D[seg[t + 1]][1]) * spacing[1]]
                surface += a[0] * b[1] - b[0] * a[1]
                perimeter += np.sqrt((a[0] - b[0]) ** 2 + (a[1] - b[1]) ** 2)
                verts.append(a)
                verts.append(b)
    surface = abs(surface) / 2.0
    v = np.unique(np.asarray(verts), axis=0) if verts else np.zeros((0, 2))
    diameter = 0.0
    for i in range(0, len(v), 512):
        d = np.sqrt(((v[i:i + 512, None, :] - v[None, :, :]) ** 2).sum(-1))
        diameter = max(diameter, float(d.max()))
    return perimeter, surface, diameter


def shape2d_features(mask_arr, spacing=(1.0, 1.0)):
    """pyradiomics ``shape2D.py`` (9 non-deprecated features, A.2 order)."""
    mask_arr = np.asarray(mask_arr, dtype=bool)
    perimeter, surface, diameter = shape2d_coefficients(mask_arr, spacing)
    coords = np.array(np.nonzero(mask_arr), dtype="int").transpose((1, 0))
    Np = len(coords)
    phys = coords * np.asarray(spacing, dtype=np.float64)[None, :]
    phys = phys - np.mean(phys, axis=0)
    phys = phys / np.sqrt(Np)
    cov = np.dot(phys.T.copy(), phys)
    ev = np.linalg.eigvals(cov)
    ev[(ev < 0) & (ev > -1e-10)] = 0
    ev = np.sort(ev)
    f = OrderedDict()
    f["Elongation"] = float("nan") if (ev[0] < 0 or ev[1] < 0) else float(np.sqrt(ev[0] / ev[1]))
    f["MajorAxisLength"] = float("nan") if ev[1] < 0 else float(np.sqrt(ev[1]) * 4)
    f["MaximumDiameter"] = diameter
    f["MeshSurface"] = surface
    f["MinorAxisLength"] = float("nan") if ev[0] < 0 else float(np.sqrt(ev[0]) * 4)
    f["Perimeter"] = perimeter
    f["PerimeterSurfaceRatio"] = perimeter / surface
    f["PixelSurface"] = Np * float(spacing[0] * spacing[1])
    f["Sphericity"] = (2 * np.sqrt(np.pi * surface)) / perimeter
    return f


# --------------------------------------------------------------------------- derived image types
def derived_image(image, image_type):
    """pyradiomics ``imageoperations.get{Square,SquareRoot,Logarithm,Exponential}Image`` (params.yml:141-144):
    point-wise transforms of the WHOLE image in float64, rescaled to the original intensity range.
    Returns ``(array, name)`` with the lower-case name pyradiomics prefixes the feature keys with."""
    im = np.asarray(image).astype("float64")
    if image_type == "Original":
        return np.asarray(image), "original"
    if image_type == "Square":
        coeff = 1 / np.sqrt(np.max(np.abs(im)))
        return (coeff * im) ** 2, "square"
    if image_type == "SquareRoot":
        coeff = np.max(np.abs(im))
        out = im.copy()
        out[im > 0] = np.sqrt(im[im > 0] * coeff)
        out[im < 0] = -np.sqrt(-im[im < 0] * coeff)
        return out, "squareroot"
    if image_type == "Logarithm":
        im_max = np.max(np.abs(im))
        out = im.copy()
        out[im > 0] = np.log(im[im > 0] + 1)
        out[im < 0] = -np.log(-(im[im < 0] - 1))
        return out * (im_max / np.max(np.abs(out))), "logarithm"
    if image_type == "Exponential":
        im_max = np.max(np.abs(im))
        coeff = np.log(im_max) / im_max
        return np.exp(coeff * im), "exponential"
    raise ValueError("image type %r is not restated in the oracle" % image_type)


# --------------------------------------------------------------------------- execute
def resolve_settings(settings=None):
    s = dict(DEFAULT_SETTINGS)
    if settings:
        s.update(settings)
    return s


def matrices(image, mask, settings=None, matrix_backend=None):
    """Discretised image + every integer matrix for one ``(image, mask)`` pair.
    ``matrix_backend`` may be the ctypes-loaded C restatement (``oracle.cmatrices``)."""
    s = resolve_settings(settings)
    image = np.asarray(image)
    mask_arr = np.asarray(mask) == s["label"]
    check_mask(mask, s["label"], s["minimumROIDimensions"], s["minimumROISize"])
    levels, gray_levels, Ng, edges = bin_image(image, mask_arr, s["binWidth"], s["binCount"])
    uni, bi = angles(image.ndim, s["distances"], s["force2D"], s["force2Ddimension"])
    mb = matrix_backend
    out = dict(levels=levels, gray_levels=gray_levels, Ng=Ng, uni=uni, bi=bi, mask=mask_arr)
    if mb is None:
        out["glcm"] = glcm_matrix(levels, Ng, uni, s["symmetricalGLCM"])
        out["glrlm"] = glrlm_matrix(levels, Ng, uni)
        out["glszm"] = glszm_matrix(levels, Ng, bi)
        out["gldm"] = gldm_matrix(levels, Ng, bi, s["gldm_a"])
        out["ngtdm_n"], out["ngtdm_s"] = ngtdm_matrix(levels, Ng, bi)
    else:
        out["glcm"] = mb.glcm(levels, Ng, uni, s["symmetricalGLCM"])
        out["glrlm"] = mb.glrlm(levels, Ng, uni)
        out["glszm"] = mb.glszm(levels, Ng, bi)
        out["gldm"] = mb.gldm(levels, Ng, bi, s["gldm_a"])
        out["ngtdm_n"], out["ngtdm_s"] = mb.ngtdm(levels, Ng, bi)
    return out


def execute_image_types(image, mask, settings=None, classes=CLASS_ORDER, image_types=("Original",), matrix_backend=None):
    """``execute`` over several enabled image types in file order: shape first (once), then one block of
    texture / first-order features per filtered image, keys prefixed with pyradiomics' image name.
    ``image_types``: names, or a mapping name -> options (``{"LoG": {"sigma": [1.0, 2.0]}}``) as in the
    ``imageType`` section of the parameter file (/root/reference/params.yml:137-145)."""
    from . import image_filters as flt

    s = resolve_settings(settings)
    out = OrderedDict()
    if "shape2D" in classes:
        out.update(execute(image, mask, settings, classes=("shape2D",)))
    rest = tuple(c for c in classes if c != "shape2D")
    items = image_types.items() if hasattr(image_types, "items") else [(t, {}) for t in image_types]
    for t, opt in items:
        if t in ("Gradient", "LoG", "Wavelet"):
            imgs = flt.filtered_images(image, t, opt or {}, s["force2D"], s["force2Ddimension"])
        else:
            arr, name = derived_image(image, t)
            imgs = {name: arr}
        for name, arr in imgs.items():
            out.update(execute(arr, mask, settings, classes=rest, image_type=name, matrix_backend=matrix_backend))
    return out


def execute(image, mask, settings=None, classes=CLASS_ORDER, image_type="original", matrix_backend=None):
    """[A.1] ``RadiomicsFeatureExtractor.execute`` for imageType Original, as called at
    ``RadiomicExtractor.py:38,42,45,48``.  Returns ``OrderedDict[name -> float]`` in the
    class order given (params.yml:164-171 order by default) with A.2 feature order."""
    s = resolve_settings(settings)
    m = matrices(image, mask, s, matrix_backend)
    gl, Ng = m["gray_levels"], m["Ng"]
    image = np.asarray(image)
    out = OrderedDict()
    if "shape2D" in classes:  # [A.1 step 3] shape keys come first, whatever the class order
        f = shape2d_features(m["mask"])
        for k in SHAPE2D_NAMES:
            out["%s_shape2D_%s" % (image_type, k)] = float(f[k])
    for cls in classes:
        if cls == "shape2D":
            continue
        if cls == "firstorder":
            f = firstorder_features(image[m["mask"]], m["levels"][m["mask"]], s["voxelArrayShift"])
        elif cls == "glcm":
            f = glcm_features(m["glcm"], gl, Ng)
        elif cls == "gldm":
            f = gldm_features(m["gldm"], gl)
        elif cls == "glrlm":
            f = glrlm_features(m["glrlm"], gl)
        elif cls == "glszm":
            f = glszm_features(m["glszm"], gl)
        elif cls == "ngtdm":
            f = ngtdm_features(m["ngtdm_n"], m["ngtdm_s"])
        else:
            raise ValueError("unknown feature class %r" % cls)
        for k in FEATURE_NAMES[cls]:
            out["%s_%s_%s" % (image_type, cls, k)] = float(np.real(f[k]))
    return out


def feature_names(classes=CLASS_ORDER, image_type="original"):
    names = ["%s_shape2D_%s" % (image_type, k) for k in SHAPE2D_NAMES] if "shape2D" in classes else []
    return names + ["%s_%s_%s" % (image_type, c, k) for c in classes if c != "shape2D" for k in FEATURE_NAMES[c]]
