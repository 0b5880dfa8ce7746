"""Pins the oracle: docstring known answers, NumPy vs C matrix builders
bit-exact, structural properties, regression fixtures."""
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import cmatrices, radiomics_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def a10():
    with open(os.path.join(GOLD, "a10_matrices.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("backend", ["numpy", "c"])
def test_a10_known_answers(a10, backend, built):
    mb = cmatrices if backend == "c" else orc
    g = (lambda name: getattr(cmatrices, name)) if backend == "c" else (lambda name: getattr(orc, name + "_matrix"))
    uni, bi = orc.angles(2)
    I1, I2, I3 = (np.array(a10[k]) for k in ("I1", "I2", "I3"))
    np.testing.assert_array_equal(g("glcm")(I1, 5, [(0, 1)], True)[:, :, 0], a10["glcm_sym_angle01"])
    np.testing.assert_array_equal(g("glrlm")(I2, 5, [(0, 1)])[:, :, 0], a10["glrlm_angle01"])
    np.testing.assert_array_equal(g("glszm")(I2, 5, bi)[:, :5], a10["glszm_8conn"])
    np.testing.assert_array_equal(g("gldm")(I2, 5, bi, 0)[:, :4], a10["gldm_alpha0_8nb"])
    n, s = g("ngtdm")(I3, 5, bi)
    np.testing.assert_array_equal(n, a10["ngtdm_n"])
    np.testing.assert_allclose(s, a10["ngtdm_s"], rtol=1e-12)


def test_angle_generator():
    # A.4: counts and the literal force2D-on-2D case (U1)
    assert orc.angles(2)[0] == [(1, 1), (0, 1), (-1, 1), (1, 0)]
    assert len(orc.angles(2)[1]) == 8
    assert orc.angles(2, force2D=True, force2Ddimension=0) == ([(0, 1)], [(0, 1), (0, -1)])
    assert orc.angles(2, force2D=True, force2Ddimension=1)[0] == [(1, 0)]
    assert len(orc.angles(3)[0]) == 13 and len(orc.angles(3, force2D=True)[0]) == 4
    assert len(orc.angles(2, distances=[1, 2])[0]) == 12


def test_feature_count_matches_reference_pin():
    # /root/reference/dataset.py:42 -> 102 = 9 shape2D + 93
    assert len(orc.feature_names()) == 93
    assert [len(orc.FEATURE_NAMES[c]) for c in orc.CLASS_ORDER] == [18, 24, 14, 16, 16, 5]
    for c in orc.CLASS_ORDER:
        assert orc.FEATURE_NAMES[c] == sorted(orc.FEATURE_NAMES[c], key=lambda n: 'get%sFeatureValue' % n)  # inspect.getmembers order


def test_binning_numpy1_semantics():
    # A.3 pitfall: uint8 arithmetic must not wrap (min 250, binWidth 20)
    img = np.array([[250, 255], [251, 253]], dtype=np.uint8)
    lev, gl, Ng, edges = orc.bin_image(img, np.ones((2, 2), bool), 20)
    assert edges[0] == 240 and Ng == 1
    lev, gl, Ng, _ = orc.bin_image(np.array([[0, 9], [10, 255]], np.uint8), np.ones((2, 2), bool), 10)
    np.testing.assert_array_equal(lev, [[1, 1], [2, 26]])


arrays = st.integers(3, 14).flatmap(lambda h: st.integers(3, 14).map(lambda w: (h, w)))


@settings(max_examples=40, deadline=None)
@given(arrays, st.integers(0, 2 ** 31 - 1), st.sampled_from([5, 10, 25, 64]), st.booleans())
def test_matrix_properties_and_backends_agree(hw, seed, bw, force2D, ):
    H, W = hw
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (H, W)).astype(np.uint8)
    if seed % 3 == 0:
        img = (img // 64 * 64).astype(np.uint8)
    mask = np.where(rng.random((H, W)) < 0.75, 255, 0).astype(np.uint8)
    s = dict(label=255, binWidth=bw, force2D=force2D)
    try:
        a = orc.matrices(img, mask, s)
    except ValueError:
        return
    b = orc.matrices(img, mask, s, matrix_backend=cmatrices)
    for k in ("glcm", "glrlm", "glszm", "gldm", "ngtdm_n"):
        np.testing.assert_array_equal(a[k], b[k])
    np.testing.assert_allclose(a["ngtdm_s"], b["ngtdm_s"], rtol=1e-12, atol=1e-12)
    lev_c, ng_c = cmatrices.bin_image(img, a["mask"], bw)
    assert ng_c == a["Ng"]
    np.testing.assert_array_equal(lev_c, a["levels"])
    Np = int(a["mask"].sum())
    P = a["glcm"]
    np.testing.assert_array_equal(P, P.transpose(1, 0, 2))                      # symmetric GLCM
    j = np.arange(1, a["glrlm"].shape[1] + 1)
    assert ((a["glrlm"].sum(0) * j[:, None]).sum(0) == Np).all()               # runs tile the ROI per angle
    assert (a["glszm"].sum(0) * np.arange(1, a["glszm"].shape[1] + 1)).sum() == Np
    assert a["gldm"].sum() == Np
    assert a["ngtdm_n"].sum() <= Np
    np.testing.assert_array_equal(a["gldm"].sum(1), np.bincount(a["levels"][a["mask"]], minlength=a["Ng"] + 1)[1:])


def test_mcc_top_eigenvalue_is_one():
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (20, 20)).astype(np.uint8)
    m = orc.matrices(img, np.full((20, 20), 255, np.uint8), dict(label=255, binWidth=32))
    P = m["glcm"][:, :, 0].astype(float)
    P /= P.sum()
    px, py = P.sum(1), P.sum(0)
    Q = (P / px[:, None]) @ (P / py[None, :]).T
    ev = np.sort(np.linalg.eigvals(Q).real)
    assert abs(ev[-1] - 1) < 1e-12


def test_oracle_regression_fixture(built):
    z = np.load(os.path.join(GOLD, "oracle_features_seed0.npz"))
    assert list(z["names"]) == orc.feature_names()
    for name, s in (("inplane_bw10", dict(label=255, binWidth=10, force2D=False)),
                    ("literal_bw10", dict(label=255, binWidth=10, force2D=True)),
                    ("inplane_bw25", dict(label=255, binWidth=25, force2D=False))):
        for b in (0, 3):
            got = np.array(list(orc.execute(z["images"][b], z["masks"][b], s, matrix_backend=cmatrices).values()))
            np.testing.assert_allclose(got, z[name][b], rtol=1e-10, atol=1e-12)


def test_check_mask_errors():
    m = np.zeros((5, 5), np.uint8)
    with pytest.raises(ValueError, match="not present"):
        orc.check_mask(m, 255)
    m[2, 2] = 255
    with pytest.raises(ValueError, match="1 segmented voxel"):
        orc.check_mask(m, 255)
    m[2, 3] = 255
    with pytest.raises(ValueError, match="too few dimensions"):
        orc.check_mask(m, 255)
    m[3, 3] = 255
    assert orc.check_mask(m, 255) == [(2, 3), (2, 3)]


def test_shape2d_oracle_against_independent_area_and_known_shapes():
    """The marching-squares tables are typed from memory of pyradiomics' cshape.c; their
    orientation is validated by comparing the shoelace surface with an independent count of
    eighths per 2x2 cell, and by closed-form values for simple shapes."""
    def area_indep(m):
        m = np.pad(np.asarray(m, bool), 1)
        a, b, c, d = (m[:-1, :-1].astype(int), m[:-1, 1:].astype(int), m[1:, 1:].astype(int), m[1:, :-1].astype(int))
        n = a + b + c + d
        diag = (a == c) & (b == d) & (a != b)
        e = np.zeros(n.shape)
        e[n == 1] = 1
        e[n == 2] = 4
        e[(n == 2) & diag] = 2
        e[n == 3] = 7
        e[n == 4] = 8
        return e.sum() / 8
    p, s, d = orc.shape2d_coefficients(np.ones((1, 1)))
    assert np.isclose(p, 4 * np.sqrt(0.5)) and s == 0.5 and d == 1.0
    p, s, d = orc.shape2d_coefficients(np.ones((3, 5)))
    assert np.isclose(p, 2 * (2 + 4) + 4 * np.sqrt(0.5)) and s == 15 - 0.5 and np.isclose(d, np.hypot(5, 2))
    rng = np.random.default_rng(3)
    for _ in range(10):
        m = rng.random((9, 11)) < 0.6
        assert np.isclose(orc.shape2d_coefficients(m)[1], area_indep(m))
    f = orc.shape2d_features(np.ones((4, 10), bool))
    assert np.isclose(f["Elongation"], np.sqrt((4 ** 2 - 1) / (10 ** 2 - 1))) and f["PixelSurface"] == 40


@settings(max_examples=25, deadline=None)
@given(arrays, st.integers(0, 2 ** 31 - 1), st.sampled_from([10, 25, 64]))
def test_oracle_against_independent_scipy_implementations(hw, seed, bw):
    """Third-party cross-checks of the restated algorithm (pyradiomics itself is not installable here): GLSZM
    zones = scipy.ndimage 8-connected components per gray level; GLCM = shifted-array pair counts; GLRLM rows =
    run-length encoding of each row; first-order moments / entropy = scipy.stats on the ROI values."""
    from scipy import ndimage, stats

    H, W = hw
    rng = np.random.default_rng(seed)
    img = (rng.integers(0, 256, (H, W)) // 32 * 32).astype(np.uint8)  # few levels: real zones and runs
    mask = np.where(rng.random((H, W)) < 0.8, 255, 0).astype(np.uint8)
    s = dict(label=255, binWidth=bw, force2D=False)
    try:
        m = orc.matrices(img, mask, s)
    except ValueError:
        return
    lev, roi, Ng = m["levels"], m["mask"], m["Ng"]
    # GLSZM: connected components (8-connectivity) of every level inside the ROI
    want = {}
    for g in range(1, Ng + 1):
        lab, n = ndimage.label((lev == g) & roi, structure=np.ones((3, 3), int))
        for size in ndimage.sum_labels(np.ones_like(lab), lab, index=np.arange(1, n + 1)).astype(int) if n else []:
            want[(g, size)] = want.get((g, size), 0) + 1
    got = {(i + 1, j + 1): int(c) for (i, j), c in np.ndenumerate(m["glszm"]) if c}
    assert got == want
    # GLCM, angle by angle: pairs (p, p + offset) with both pixels in the ROI, symmetrised
    offs = orc.angles(2)[0]
    for a, (dy, dx) in enumerate(offs):
        P = np.zeros((Ng, Ng), int)
        for y in range(H):
            for x in range(W):
                yy, xx = y + dy, x + dx
                if 0 <= yy < H and 0 <= xx < W and roi[y, x] and roi[yy, xx]:
                    P[lev[y, x] - 1, lev[yy, xx] - 1] += 1
        np.testing.assert_array_equal(m["glcm"][:, :, a], P + P.T)
    # GLRLM along rows (the (0, 1) angle): itertools-style run-length encoding
    a_row = [i for i, o in enumerate(offs) if tuple(o) == (0, 1)][0]
    R = np.zeros_like(m["glrlm"][:, :, a_row])
    for y in range(H):
        x = 0
        while x < W:
            g = lev[y, x] if roi[y, x] else 0
            e = x
            while e + 1 < W and (lev[y, e + 1] if roi[y, e + 1] else 0) == g:
                e += 1
            if g:
                R[g - 1, e - x] += 1
            x = e + 1
    np.testing.assert_array_equal(m["glrlm"][:, :, a_row], R)
    # first-order statistics on the raw ROI values
    f = orc.execute(img, mask, s, classes=("firstorder",))
    x = img[roi].astype(np.float64)
    np.testing.assert_allclose(f["original_firstorder_Skewness"], stats.skew(x) if x.std() > 0 else 0.0, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(f["original_firstorder_Kurtosis"], stats.kurtosis(x, fisher=False) if x.std() > 0 else 0.0, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(f["original_firstorder_Variance"], x.var(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(f["original_firstorder_InterquartileRange"], stats.iqr(x), rtol=1e-12, atol=1e-12)
    p = np.bincount(lev[roi])[1:] / roi.sum()
    np.testing.assert_allclose(f["original_firstorder_Entropy"], stats.entropy(p[p > 0], base=2), rtol=1e-9, atol=1e-9)
