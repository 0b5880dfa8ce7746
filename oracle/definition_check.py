"""Definition-level checker of the 93 + 9 radiomic features.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED like the rest of ``oracle/`` (pyradiomics is not installable here).  This module exists so
that every formula has TWO independent derivations inside the repo: ``radiomics_oracle.py`` restates
pyradiomics' vectorised NumPy code; this file is written from the *definitions* (IBSI reference manual
notation: explicit loops over voxel pairs, runs, zones and dictionary-of-cells probabilities, pure Python
floats) and shares no code, helper or array layout with it.  ``tests/test_definition_check.py`` compares the
two (hypothesis, 1e-10) and the golden fixtures.

Where pyradiomics 3.1.0 deviates from the IBSI definitions the deviation is applied here *explicitly* and
listed (``DEVIATIONS``), so that a reader can see which numbers are "IBSI" and which are "pyradiomics":
the reference consumes pyradiomics' numbers (/root/reference/RadiomicExtractor.py:38-48).
"""
from __future__ import annotations

import math
from collections import defaultdict
from fractions import Fraction
from itertools import groupby

EPS = 2.220446049250313e-16  # numpy.spacing(1)

DEVIATIONS = {
    "firstorder_Kurtosis": "pyradiomics returns m4/m2^2 (IBSI: excess kurtosis, -3)",
    "firstorder_Entropy/Uniformity": "computed on the discretised levels; log2(p + eps)",
    "firstorder_Variance/Skewness": "population moments (divide by N)",
    "*_Entropy": "every entropy is -sum p log2(p + eps) with eps = 2.2e-16 inside the logarithm",
    "glcm_Correlation": "denominator sigma_x sigma_y + eps; 1 when sigma_x sigma_y = 0",
    "glcm_Imc1": "0 when max(HX, HY) = 0",
    "glcm_Imc2": "sqrt(max(0, 1 - exp(-2 (HXY2 - HXY))))",
    "glcm_MCC": "sqrt of the second largest eigenvalue of Q, Q(i,j) = sum_k p(i,k) p(j,k) / (px(i) py(k) + eps); 1 for a one-level ROI",
    "glcm_Idmn/Idn": "normalised by Ng = the largest gray level, not the number of levels present",
    "glcm/glrlm aggregation": "features per direction, then the mean over the non-empty directions",
    "ngtdm_Coarseness": "1e6 when sum p_i s_i = 0",
    "gldm": "dependence counts the voxel's neighbours only (size j = dep + 1); alpha = gldm_a",
}

IN_PLANE_UNI = ((1, 1), (0, 1), (-1, 1), (1, 0))  # distance-1 offsets of a plane, one per +- pair
GLCM_NAMES = ("Autocorrelation", "ClusterProminence", "ClusterShade", "ClusterTendency", "Contrast", "Correlation",
              "DifferenceAverage", "DifferenceEntropy", "DifferenceVariance", "Id", "Idm", "Idmn", "Idn", "Imc1", "Imc2",
              "InverseVariance", "JointAverage", "JointEnergy", "JointEntropy", "MCC", "MaximumProbability", "SumAverage",
              "SumEntropy", "SumSquares")


def _sq(v):
    return v * v


def _entropy(probs):
    return -sum(p * math.log2(p + EPS) for p in probs)


# ------------------------------------------------------------------------------------------ discretisation
def discretise(image, mask, label, bin_width):
    """Fixed bin width, integer pixels: level = floor((x - low) / w) + 1 with low = min - (min mod w), exact
    rational arithmetic.  Returns ``levels`` (list of rows, 0 outside the ROI) and the ROI values."""
    H, W = len(image), len(image[0])
    roi = [(y, x) for y in range(H) for x in range(W) if int(mask[y][x]) == label]
    vals = [int(image[y][x]) for y, x in roi]
    w = Fraction(bin_width).limit_denominator(10 ** 6)
    lo = min(vals)
    low = lo - (Fraction(lo) - (Fraction(lo) // w) * w)  # Python's floor-mod: sign of the divisor
    levels = [[0] * W for _ in range(H)]
    for (y, x), v in zip(roi, vals):
        levels[y][x] = int((Fraction(v) - low) // w) + 1
    return levels, vals


# ------------------------------------------------------------------------------------------ first order
def _percentile(sorted_vals, q):
    pos = q / 100.0 * (len(sorted_vals) - 1)
    k = int(math.floor(pos))
    if k + 1 >= len(sorted_vals):
        return float(sorted_vals[-1])
    return sorted_vals[k] + (sorted_vals[k + 1] - sorted_vals[k]) * (pos - k)


def firstorder(vals, levels_of_vals, shift=0.0):
    x = [float(v) for v in vals]
    n = len(x)
    s = sorted(x)
    mean = math.fsum(x) / n
    m2 = math.fsum(_sq(v - mean) for v in x) / n
    m3 = math.fsum((v - mean) ** 3 for v in x) / n
    m4 = math.fsum((v - mean) ** 4 for v in x) / n
    p10, p25, p50, p75, p90 = (_percentile(s, q) for q in (10, 25, 50, 75, 90))
    inner = [v for v in x if p10 <= v <= p90]
    imean = math.fsum(inner) / len(inner)
    counts = defaultdict(int)
    for g in levels_of_vals:
        counts[g] += 1
    probs = [c / n for c in counts.values()]
    energy = math.fsum(_sq(v + shift) for v in x)
    return {
        "10Percentile": p10, "90Percentile": p90, "Energy": energy, "Entropy": _entropy(probs),
        "InterquartileRange": p75 - p25, "Kurtosis": 0.0 if m2 == 0 else m4 / (m2 * m2), "Maximum": s[-1],
        "MeanAbsoluteDeviation": math.fsum(abs(v - mean) for v in x) / n, "Mean": mean, "Median": p50, "Minimum": s[0],
        "Range": s[-1] - s[0], "RobustMeanAbsoluteDeviation": math.fsum(abs(v - imean) for v in inner) / len(inner),
        "RootMeanSquared": math.sqrt(energy / n), "Skewness": 0.0 if m2 == 0 else m3 / m2 ** 1.5, "TotalEnergy": energy,
        "Uniformity": math.fsum(p * p for p in probs), "Variance": m2,
    }


# ------------------------------------------------------------------------------------------ GLCM
def glcm_cells(levels, offset, symmetric=True):
    """{(i, j): count} over ordered voxel pairs (v, v + offset), both in the ROI."""
    H, W = len(levels), len(levels[0])
    dy, dx = offset
    cells = defaultdict(int)
    for y in range(H):
        for x in range(W):
            yy, xx = y + dy, x + dx
            if 0 <= yy < H and 0 <= xx < W and levels[y][x] and levels[yy][xx]:
                cells[(levels[y][x], levels[yy][xx])] += 1
                if symmetric:
                    cells[(levels[yy][xx], levels[y][x])] += 1
    return cells


def _second_largest_eigenvalue(p, px, py, n_roi_levels):
    """Second largest eigenvalue of Q(i,j) = sum_k p(i,k) p(j,k) / (px(i) py(k)) over the gray levels of the ROI.
    Q = Dx^-1 P Dy^-1 P^T is similar to the symmetric positive semi-definite S = Dx^-1/2 P Dy^-1 P^T Dx^-1/2, so
    the symmetric LAPACK solver (eigvalsh) is used on S -- a different route from the general eigvals(Q)
    pyradiomics (and radiomics_oracle.py) take.  Levels with px = 0 (possible for a non-symmetric matrix) or
    without any pair in this direction have an all-zero row in Q (pyradiomics: 0 / eps) and only contribute
    zero eigenvalues."""
    import numpy as np

    rows = sorted(i for i, v in px.items() if v > 0)
    cols = sorted(k for k, v in py.items() if v > 0)
    B = np.array([[p.get((i, k), 0.0) / math.sqrt(px[i] * py[k]) for k in cols] for i in rows], dtype=np.float64)
    ev = list(np.linalg.eigvalsh(B @ B.T)) + [0.0] * (n_roi_levels - len(rows))
    return sorted(ev)[-2]


def glcm_direction(cells, ng_max, n_roi_levels):
    """The 24 features of one direction from its {(i, j): count} cells; None for an empty direction."""
    total = sum(cells.values())
    if total == 0:
        return None
    p = {k: c / total for k, c in cells.items()}
    px, py, padd, psub = defaultdict(float), defaultdict(float), defaultdict(float), defaultdict(float)
    for (i, j), v in p.items():
        px[i] += v
        py[j] += v
        padd[i + j] += v
        psub[abs(i - j)] += v
    ux = sum(i * v for i, v in px.items())
    uy = sum(j * v for j, v in py.items())
    sx = math.sqrt(sum(_sq(i - ux) * v for i, v in px.items()))
    sy = math.sqrt(sum(_sq(j - uy) * v for j, v in py.items()))
    hxy = _entropy(p.values())
    hx, hy = _entropy(px.values()), _entropy(py.values())
    hxy1 = -sum(v * math.log2(px[i] * py[j] + EPS) for (i, j), v in p.items())
    # HXY2 runs over ALL (i, j) of the present levels, not only over the non-empty cells
    hxy2 = -sum(a * b * math.log2(a * b + EPS) for a in px.values() for b in py.values())
    da = sum(k * v for k, v in psub.items())
    f = {}
    f["Autocorrelation"] = sum(i * j * v for (i, j), v in p.items())
    f["ClusterProminence"] = sum((i + j - ux - uy) ** 4 * v for (i, j), v in p.items())
    f["ClusterShade"] = sum((i + j - ux - uy) ** 3 * v for (i, j), v in p.items())
    f["ClusterTendency"] = sum((i + j - ux - uy) ** 2 * v for (i, j), v in p.items())
    f["Contrast"] = sum(_sq(i - j) * v for (i, j), v in p.items())
    cov = sum((i - ux) * (j - uy) * v for (i, j), v in p.items())
    f["Correlation"] = 1.0 if sx * sy == 0 else cov / (sx * sy + EPS)
    f["DifferenceAverage"] = da
    f["DifferenceEntropy"] = _entropy(psub.values())
    f["DifferenceVariance"] = sum(_sq(k - da) * v for k, v in psub.items())
    f["Id"] = sum(v / (1 + k) for k, v in psub.items())
    f["Idm"] = sum(v / (1 + k * k) for k, v in psub.items())
    f["Idmn"] = sum(v / (1 + k * k / (ng_max * ng_max)) for k, v in psub.items())
    f["Idn"] = sum(v / (1 + k / ng_max) for k, v in psub.items())
    f["Imc1"] = 0.0 if max(hx, hy) == 0 else (hxy - hxy1) / max(hx, hy)
    f["Imc2"] = math.sqrt(max(0.0, 1 - math.exp(-2 * (hxy2 - hxy))))
    f["InverseVariance"] = sum(v / (k * k) for k, v in psub.items() if k > 0)
    f["JointAverage"] = ux
    f["JointEnergy"] = sum(v * v for v in p.values())
    f["JointEntropy"] = hxy
    f["MCC"] = math.sqrt(max(_second_largest_eigenvalue(p, px, py, n_roi_levels), 0.0)) if n_roi_levels > 1 else 1.0
    f["MaximumProbability"] = max(p.values())
    f["SumAverage"] = sum(k * v for k, v in padd.items())
    f["SumEntropy"] = _entropy(padd.values())
    f["SumSquares"] = sum(_sq(i - ux) * v for i, v in px.items())
    return f


def _mean_over_directions(per_dir):
    live = [d for d in per_dir if d is not None]
    if not live:
        return None
    return {k: sum(d[k] for d in live) / len(live) for k in live[0]}


def glcm(levels, offsets=IN_PLANE_UNI, symmetric=True):
    present = {g for row in levels for g in row if g}
    ng_max = max(present)
    out = _mean_over_directions([glcm_direction(glcm_cells(levels, o, symmetric), ng_max, len(present)) for o in offsets])
    if out is None:  # no voxel pair in any direction: pyradiomics normalises by a NaN sum -> every feature is NaN
        out = {k: float("nan") for k in GLCM_NAMES}
    if len(present) < 2:
        out["MCC"] = 1.0
    return out


# ------------------------------------------------------------------------------------------ GLRLM
def _lines(H, W, offset):
    """Every maximal straight line of the grid along ``offset``, as lists of (y, x)."""
    dy, dx = offset
    lines = []
    for y in range(H):
        for x in range(W):
            py_, px_ = y - dy, x - dx
            if 0 <= py_ < H and 0 <= px_ < W:
                continue  # not the first voxel of its line
            line, yy, xx = [], y, x
            while 0 <= yy < H and 0 <= xx < W:
                line.append((yy, xx))
                yy += dy
                xx += dx
            lines.append(line)
    return lines


def glrlm_runs(levels, offset):
    """[(level, length)] of the maximal same-level ROI runs along one direction."""
    runs = []
    for line in _lines(len(levels), len(levels[0]), offset):
        for g, grp in groupby(levels[y][x] for y, x in line):
            if g:
                runs.append((g, len(list(grp))))
    return runs


def _run_like(cells, n_voxels_hint=None):
    """Shared definition of the run-length / size-zone / dependence families: ``cells`` = {(i, j): count}."""
    n = sum(cells.values())
    pi, pj = defaultdict(int), defaultdict(int)
    for (i, j), c in cells.items():
        pi[i] += c
        pj[j] += c
    nv = sum(j * c for j, c in pj.items())
    mu_i = sum(i * c for i, c in pi.items()) / n
    mu_j = sum(j * c for j, c in pj.items()) / n
    return {
        "n": n, "nv": nv,
        "short": sum(c / (j * j) for j, c in pj.items()) / n,
        "long": sum(c * j * j for j, c in pj.items()) / n,
        "gl_nu": sum(c * c for c in pi.values()) / n,
        "gl_nu_n": sum(c * c for c in pi.values()) / (n * n),
        "j_nu": sum(c * c for c in pj.values()) / n,
        "j_nu_n": sum(c * c for c in pj.values()) / (n * n),
        "gl_var": sum(c * _sq(i - mu_i) for i, c in pi.items()) / n,
        "j_var": sum(c * _sq(j - mu_j) for j, c in pj.items()) / n,
        "entropy": _entropy([c / n for c in cells.values()]),
        "low_gl": sum(c / (i * i) for i, c in pi.items()) / n,
        "high_gl": sum(c * i * i for i, c in pi.items()) / n,
        "short_low": sum(c / (i * i * j * j) for (i, j), c in cells.items()) / n,
        "short_high": sum(c * i * i / (j * j) for (i, j), c in cells.items()) / n,
        "long_low": sum(c * j * j / (i * i) for (i, j), c in cells.items()) / n,
        "long_high": sum(c * i * i * j * j for (i, j), c in cells.items()) / n,
    }


def glrlm(levels, offsets=IN_PLANE_UNI):
    per_dir = []
    for o in offsets:
        cells = defaultdict(int)
        for g, ln in glrlm_runs(levels, o):
            cells[(g, ln)] += 1
        if not cells:
            per_dir.append(None)
            continue
        r = _run_like(cells)
        per_dir.append({
            "GrayLevelNonUniformity": r["gl_nu"], "GrayLevelNonUniformityNormalized": r["gl_nu_n"],
            "GrayLevelVariance": r["gl_var"], "HighGrayLevelRunEmphasis": r["high_gl"], "LongRunEmphasis": r["long"],
            "LongRunHighGrayLevelEmphasis": r["long_high"], "LongRunLowGrayLevelEmphasis": r["long_low"],
            "LowGrayLevelRunEmphasis": r["low_gl"], "RunEntropy": r["entropy"], "RunLengthNonUniformity": r["j_nu"],
            "RunLengthNonUniformityNormalized": r["j_nu_n"], "RunPercentage": r["n"] / r["nv"], "RunVariance": r["j_var"],
            "ShortRunEmphasis": r["short"], "ShortRunHighGrayLevelEmphasis": r["short_high"],
            "ShortRunLowGrayLevelEmphasis": r["short_low"]})
    return _mean_over_directions(per_dir) or {}


# ------------------------------------------------------------------------------------------ GLSZM
def glszm_zones(levels, offsets=IN_PLANE_UNI):
    """[(level, size)] of the connected same-level zones; connectivity = the +- offsets (8 in a plane).
    Labelling is delegated to scipy.ndimage.label, level by level (a third-party implementation)."""
    import numpy as np
    from scipy import ndimage

    L = np.asarray(levels)
    st = np.zeros((3, 3), dtype=int)
    st[1, 1] = 1
    for dy, dx in offsets:
        st[1 + dy, 1 + dx] = 1
        st[1 - dy, 1 - dx] = 1
    zones = []
    for g in sorted(set(L[L > 0].tolist())):
        lab, n = ndimage.label(L == g, structure=st)
        sizes = np.bincount(lab.ravel())[1:]
        zones += [(int(g), int(s)) for s in sizes]
    return zones


def glszm(levels, offsets=IN_PLANE_UNI):
    cells = defaultdict(int)
    for z in glszm_zones(levels, offsets):
        cells[z] += 1
    r = _run_like(cells)
    return {
        "GrayLevelNonUniformity": r["gl_nu"], "GrayLevelNonUniformityNormalized": r["gl_nu_n"], "GrayLevelVariance": r["gl_var"],
        "HighGrayLevelZoneEmphasis": r["high_gl"], "LargeAreaEmphasis": r["long"], "LargeAreaHighGrayLevelEmphasis": r["long_high"],
        "LargeAreaLowGrayLevelEmphasis": r["long_low"], "LowGrayLevelZoneEmphasis": r["low_gl"],
        "SizeZoneNonUniformity": r["j_nu"], "SizeZoneNonUniformityNormalized": r["j_nu_n"], "SmallAreaEmphasis": r["short"],
        "SmallAreaHighGrayLevelEmphasis": r["short_high"], "SmallAreaLowGrayLevelEmphasis": r["short_low"],
        "ZoneEntropy": r["entropy"], "ZonePercentage": r["n"] / r["nv"], "ZoneVariance": r["j_var"]}


# ------------------------------------------------------------------------------------------ GLDM / NGTDM
def _neighbours(levels, y, x, offsets):
    H, W = len(levels), len(levels[0])
    out = []
    for dy, dx in offsets:
        for s in (1, -1):
            yy, xx = y + s * dy, x + s * dx
            if 0 <= yy < H and 0 <= xx < W and levels[yy][xx]:
                out.append(levels[yy][xx])
    return out


def gldm(levels, offsets=IN_PLANE_UNI, alpha=0):
    cells = defaultdict(int)
    for y, row in enumerate(levels):
        for x, g in enumerate(row):
            if g:
                dep = sum(1 for v in _neighbours(levels, y, x, offsets) if abs(v - g) <= alpha)
                cells[(g, dep + 1)] += 1
    r = _run_like(cells)
    return {
        "DependenceEntropy": r["entropy"], "DependenceNonUniformity": r["j_nu"], "DependenceNonUniformityNormalized": r["j_nu_n"],
        "DependenceVariance": r["j_var"], "GrayLevelNonUniformity": r["gl_nu"], "GrayLevelVariance": r["gl_var"],
        "HighGrayLevelEmphasis": r["high_gl"], "LargeDependenceEmphasis": r["long"],
        "LargeDependenceHighGrayLevelEmphasis": r["long_high"], "LargeDependenceLowGrayLevelEmphasis": r["long_low"],
        "LowGrayLevelEmphasis": r["low_gl"], "SmallDependenceEmphasis": r["short"],
        "SmallDependenceHighGrayLevelEmphasis": r["short_high"], "SmallDependenceLowGrayLevelEmphasis": r["short_low"]}


def ngtdm(levels, offsets=IN_PLANE_UNI):
    n, s = defaultdict(int), defaultdict(float)
    for y, row in enumerate(levels):
        for x, g in enumerate(row):
            if not g:
                continue
            nb = _neighbours(levels, y, x, offsets)
            if nb:
                n[g] += 1
                s[g] += abs(g - sum(nb) / len(nb))
    nvp = sum(n.values())
    if nvp == 0:
        return {k: float("nan") for k in ("Busyness", "Coarseness", "Complexity", "Contrast", "Strength")}
    lv = sorted(n)
    p = {i: n[i] / nvp for i in lv}
    ngp = len(lv)
    sum_ps = sum(p[i] * s[i] for i in lv)
    sum_s = sum(s[i] for i in lv)
    busy_den = sum(abs(i * p[i] - j * p[j]) for i in lv for j in lv)
    cplx = sum(abs(i - j) * (p[i] * s[i] + p[j] * s[j]) / (p[i] + p[j]) for i in lv for j in lv) / nvp
    contrast = 0.0
    if ngp > 1:
        contrast = sum(p[i] * p[j] * _sq(i - j) for i in lv for j in lv) / (ngp * (ngp - 1)) * sum_s / nvp
    strength = sum((p[i] + p[j]) * _sq(i - j) for i in lv for j in lv) / sum_s if sum_s != 0 else 0.0
    return {"Busyness": sum_ps / busy_den if busy_den != 0 else 0.0, "Coarseness": 1 / sum_ps if sum_ps != 0 else 1e6,
            "Complexity": cplx, "Contrast": contrast, "Strength": strength}


# ------------------------------------------------------------------------------------------ shape2D
def shape2d(mask, label):
    """Mask-only descriptors.  Contour = marching squares over the zero-padded mask, derived here per 2x2 cell
    from first principles: the contour crosses every cell edge whose two corners differ, at the edge midpoint;
    area of the polygon part inside a cell by corner count (1 -> 1/8, 2 adjacent -> 1/2, 2 diagonal -> 2/8
    (two separate corner triangles, as pyradiomics' line table cuts it), 3 -> 7/8, 4 -> 1)."""
    H, W = len(mask), len(mask[0])
    inside = lambda y, x: 0 <= y < H and 0 <= x < W and int(mask[y][x]) == label
    pts = [(y, x) for y in range(H) for x in range(W) if inside(y, x)]
    n = len(pts)
    area = 0.0
    perim = 0.0
    verts = set()
    for y in range(-1, H):
        for x in range(-1, W):
            c = [inside(y, x), inside(y, x + 1), inside(y + 1, x + 1), inside(y + 1, x)]  # clockwise from top-left
            k = sum(c)
            if k == 0:
                continue
            if k == 4:
                area += 1.0
                continue
            mids = [(y, x + 0.5), (y + 0.5, x + 1), (y + 1, x + 0.5), (y + 0.5, x)]  # top, right, bottom, left edge
            crossed = [mids[e] for e in range(4) if c[e] != c[(e + 1) % 4]]
            verts.update(crossed)
            if k == 2 and c[0] == c[2]:  # diagonal pair: two corner cuts
                area += 2 / 8
                perim += 2 * math.sqrt(0.5)
            elif k == 2:
                area += 0.5
                perim += 1.0
            else:
                area += 1 / 8 if k == 1 else 7 / 8
                perim += math.sqrt(0.5)
    diam = max((math.dist(a, b) for a in verts for b in verts), default=0.0)
    my = sum(p[0] for p in pts) / n
    mx = sum(p[1] for p in pts) / n
    syy = sum(_sq(p[0] - my) for p in pts) / n
    sxx = sum(_sq(p[1] - mx) for p in pts) / n
    sxy = sum((p[0] - my) * (p[1] - mx) for p in pts) / n
    tr, det = syy + sxx, syy * sxx - sxy * sxy
    disc = math.sqrt(max(tr * tr / 4 - det, 0.0))
    l1, l0 = tr / 2 + disc, max(tr / 2 - disc, 0.0)
    return {"Elongation": math.sqrt(l0 / l1), "MajorAxisLength": 4 * math.sqrt(l1), "MaximumDiameter": diam,
            "MeshSurface": area, "MinorAxisLength": 4 * math.sqrt(l0), "Perimeter": perim,
            "PerimeterSurfaceRatio": perim / area, "PixelSurface": float(n),
            "Sphericity": 2 * math.sqrt(math.pi * area) / perim}


# ------------------------------------------------------------------------------------------ everything
def all_features(image, mask, label=255, bin_width=25, offsets=IN_PLANE_UNI, symmetric=True, alpha=0, shape=False):
    """{'original_<class>_<Feature>': value} for an integer image (lists or arrays), in no particular order."""
    image = [list(map(int, r)) for r in image]
    mask = [list(map(int, r)) for r in mask]
    levels, vals = discretise(image, mask, label, bin_width)
    lv_of_vals = [g for row in levels for g in row if g]
    out = {}
    if shape:
        out.update({"original_shape2D_" + k: v for k, v in shape2d(mask, label).items()})
    for cls, feats in (("firstorder", firstorder(vals, lv_of_vals)), ("glcm", glcm(levels, offsets, symmetric)),
                       ("gldm", gldm(levels, offsets, alpha)), ("glrlm", glrlm(levels, offsets)),
                       ("glszm", glszm(levels, offsets)), ("ngtdm", ngtdm(levels, offsets))):
        out.update({"original_%s_%s" % (cls, k): v for k, v in feats.items()})
    return out
