"""CPU check of the CUDA kernel SOURCE: multimodal-isic_b200/csrc/radb_kernels.cuh compiled
under the thread-level emulation in tests/emu and compared with the oracle (bit-exact
matrices, features within rtol 1e-6 / atol 1e-9).  The GPU parity tests (-m gpu) make the
same comparisons through the real C-ABI library."""
import os

import numpy as np
import pytest

from multimodal_isic_b200 import synth
from oracle import radiomics_oracle as orc
from tests.emu_runner import compare_with_oracle, edge_case_batch, word_pass_stress_batch

GOLD = os.path.join(os.path.dirname(__file__), "golden")
INPLANE = orc.angles(2)[0]
LITERAL = orc.angles(2, force2D=True)[0]


def test_emu_synthetic_inplane(emu):
    imgs, masks = synth.make_patches(3, 64, seed=0)
    r = emu.run(imgs, masks, 10, 255, INPLANE)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=10, force2D=False)) == 3


def test_emu_literal_force2d_and_bw25(emu):
    imgs, masks = synth.make_patches(2, 32, seed=1)
    r = emu.run(imgs, masks, 25, 255, LITERAL)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=25, force2D=True)) == 2


def test_emu_nonsymmetric_glcm_and_alpha(emu):
    imgs, masks = synth.make_patches(2, 24, seed=2)
    r = emu.run(imgs, masks, 16, 255, INPLANE, symmetrical=False, alpha=1)
    s = dict(label=255, binWidth=16, force2D=False, symmetricalGLCM=False, gldm_a=1)
    assert compare_with_oracle(r, imgs, masks, s) == 2


def test_emu_edge_cases(emu):
    imgs, masks = edge_case_batch()
    r = emu.run(imgs, masks, 10, 255, INPLANE)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=10, force2D=False)) == 6
    assert list(r["status"][:3]) == [1, 2, 3]


@pytest.mark.parametrize("bw,angles,force2d", [(25, INPLANE, False), (20, LITERAL, True), (32, INPLANE, False)])
def test_emu_thread_per_angle_kernel(emu, monkeypatch, bw, angles, force2d):
    """Few gray levels select radb_angle_lane_kernel (one thread per (patch, angle)); RADB_NO_LANE forces the
    warp-per-angle kernel.  Both must match the oracle, edge cases included, and each other."""
    a, am = synth.make_patches(5, 20, seed=11)
    e, em_ = edge_case_batch(H=20, W=20)
    imgs, masks = np.concatenate([a, e]), np.concatenate([am, em_])
    r = emu.run(imgs, masks, bw, 255, angles)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=bw, force2D=force2d)) == 11
    monkeypatch.setenv("RADB_NO_LANE", "1")
    r2 = emu.run(imgs, masks, bw, 255, angles)
    assert compare_with_oracle(r2, imgs, masks, dict(label=255, binWidth=bw, force2D=force2d), check_matrices=False) == 11
    np.testing.assert_allclose(r["features"], r2["features"], rtol=1e-8, atol=1e-11, equal_nan=True)
    assert not np.array_equal(r["features"], r2["features"], equal_nan=True)  # they really are different code paths


def test_emu_large_zones_overflow_path(emu):
    # smooth image -> zones larger than the dense GLSZM columns (overflow list)
    H = W = 48
    yy, xx = np.mgrid[:H, :W]
    img = np.clip(100 + 40 * np.sin(xx / 9.0) + 30 * np.cos(yy / 7.0), 0, 255).astype(np.uint8)[None]
    mask = np.full((1, H, W), 255, np.uint8)
    r = emu.run(img, mask, 10, 255, INPLANE)
    assert compare_with_oracle(r, img, mask, dict(label=255, binWidth=10, force2D=False)) == 1
    assert (r["glszm"][0].sum(0)[16:] > 0).any()


def test_emu_many_large_zones_take_the_warp_level_glszm(emu):
    # 1-pixel stripes of alternating levels: 60 zones of 60 pixels each -> longer overflow list than the
    # thread-per-patch GLSZM task accepts (RADB_LANE_MAX_OVF = 48), so the warp-level task must take the patch;
    # the second patch has 30 such zones and stays on the thread-level task
    H = W = 60
    img = np.zeros((2, H, W), np.uint8)
    img[0] = np.where(np.arange(W) % 2 == 0, 40, 200)[None, :]
    img[1] = np.where((np.arange(W) // 2) % 2 == 0, 40, 200)[None, :]
    mask = np.full((2, H, W), 255, np.uint8)
    r = emu.run(img, mask, 25, 255, INPLANE)
    assert compare_with_oracle(r, img, mask, dict(label=255, binWidth=25, force2D=False)) == 2
    assert r["glszm"][0].sum() == 60 and r["glszm"][1].sum() == 30


def test_emu_ragged_mixed_sizes(emu):
    """BASELINE.json configs[3] in miniature: patches of mixed size and mask coverage in ONE call; rows come
    back in input order whatever the grouping, invalid ROIs keep their status, odd sizes lose the TMA path."""
    rng = np.random.default_rng(5)
    sizes = [(16, 16), (24, 20), (16, 16), (33, 17), (24, 20), (16, 16), (8, 40)]
    images, masks = [], []
    for i, (H, W) in enumerate(sizes):
        g, m = synth.make_patches(1, H, W, seed=20 + i)
        images.append(g[0])
        masks.append(m[0])
    masks[2] = np.zeros_like(masks[2])            # label absent
    masks[5] = (rng.random((16, 16)) < 0.9).astype(np.uint8) * 255  # near-full coverage
    out, status = emu.run_ragged(images, masks, 25, 255, INPLANE)
    assert list(status) == [0, 0, 1, 0, 0, 0, 0]
    s = orc.resolve_settings(dict(label=255, binWidth=25, force2D=False))
    names = orc.feature_names(orc.CLASS_ORDER)
    for i in range(len(sizes)):
        if status[i]:
            assert np.isnan(out[i]).all()
            continue
        f = orc.execute(images[i], masks[i], s)
        np.testing.assert_allclose(out[i], [f[k] for k in names], rtol=1e-6, atol=1e-9)
    # same rows as the dense entry point, group by group
    dense = emu.run(np.stack([images[0], images[5]]), np.stack([masks[0], masks[5]]), 25, 255, INPLANE)["features"]
    assert np.array_equal(dense, out[[0, 5]])


@pytest.mark.parametrize("ng_cap,bw", [(256, 8), (128, 16)])
def test_emu_many_gray_levels_big_mode(emu, ng_cap, bw):
    """BASELINE.json configs[4] (binWidth sweep on uint16 intensities in [0, 2048)): 128 and 256 gray levels.
    The GLCM counters and the MCC workspace no longer fit shared memory (big mode: global memory), and 256
    levels need the u16 level image."""
    g, masks = synth.make_patches(2, 20, seed=31, dtype=np.uint16, vmax=2047)
    g[1] = (np.arange(400).reshape(20, 20) * 5 % 2048).astype(np.uint16)  # every pixel its own level: Ng = 256 exactly
    masks[1] = 255
    r = emu.run(g, masks, bw, 255, INPLANE, max_ng=ng_cap)
    assert compare_with_oracle(r, g, masks, dict(label=255, binWidth=bw, force2D=False)) == 2
    assert r["ng"].max() > (200 if bw == 8 else 100)


@pytest.mark.parametrize("dt", [np.uint8, np.float64])
def test_emu_bin_count(emu, dt):
    """binCount binning (imageoperations.getBinEdges): the ROI range split into n bins by numpy.histogram's edges
    (last edge + 1), including ROIs whose edges fall exactly on pixel values, a flat ROI and a two-valued ROI."""
    g, masks = synth.make_patches(5, 24, 22, seed=13)
    g[1] = (g[1] // 16) * 16                 # pixel values on a lattice: edges hit pixel values exactly
    if dt == np.uint8:  # (flat *float* ROIs: Skewness / Kurtosis deviate on purpose)
        g[2][masks[2] == 255] = 77           # flat ROI: numpy.histogram uses the range min -+ 0.5
    g[3] = np.where(g[3] > 120, 200, 50)     # two values
    imgs = g if dt == np.uint8 else (np.sqrt(g.astype(np.float64)) * 3.0 - 7.0)
    for n in (1, 6, 20, 64):
        r = emu.run(imgs, masks, 25, 255, INPLANE, bin_count=n)
        assert compare_with_oracle(r, imgs, masks, dict(label=255, binCount=n, force2D=False)) == 5
        assert r["ng"].max() == n


def test_emu_golden_fixture(emu):
    z = np.load(os.path.join(GOLD, "oracle_features_seed0.npz"))
    r = emu.run(z["images"][:2], z["masks"][:2], 10, 255, INPLANE)
    np.testing.assert_allclose(r["features"], z["inplane_bw10"][:2], rtol=1e-6, atol=1e-9)


def test_emu_force2d_dimension1_column_only(emu):
    # force2Ddimension=1 removes axis 1: the only offset is (1, 0); zones are column runs
    imgs, masks = synth.make_patches(2, 24, 20, seed=3)
    ang = orc.angles(2, force2D=True, force2Ddimension=1)[0]
    assert ang == [(1, 0)]
    r = emu.run(imgs, masks, 10, 255, ang)
    s = dict(label=255, binWidth=10, force2D=True, force2Ddimension=1)
    assert compare_with_oracle(r, imgs, masks, s) == 2


def test_emu_bitwise_reproducible(emu):
    # thread interleaving differs between runs; the output must not (integer atomics only,
    # rank-sorted overflow list, fixed reduction trees)
    H = W = 40
    yy, xx = np.mgrid[:H, :W]
    img = np.clip(100 + 40 * np.sin(xx / 9.0) + 30 * np.cos(yy / 7.0), 0, 255).astype(np.uint8)[None]
    mask = np.full((1, H, W), 255, np.uint8)
    a = emu.run(img, mask, 10, 255, INPLANE)["features"]
    for _ in range(3):
        b = emu.run(img, mask, 10, 255, INPLANE)["features"]
        assert a.tobytes() == b.tobytes()


def test_emu_edge_cases_with_perturbed_tables(emu, monkeypatch):
    # the device log2 and the host-built log2 table may differ in the last bit: degenerate ROIs
    # (one gray level: Imc2 must be exactly 0) must not depend on it
    monkeypatch.setenv("RADB_EMU_PERTURB_TABLES", "1")
    imgs, masks = edge_case_batch()
    r = emu.run(imgs, masks, 10, 255, INPLANE)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=10, force2D=False)) == 6


def test_emu_wide_mode_whole_image(emu):
    # > 65535 pixels: level image, union-find words, GLRLM and the zone overflow list move to global
    # memory (whole dermoscopy images, RadiomicExtractor.py:29-38, are this size class)
    imgs, masks = synth.make_patches(1, 270, 300, seed=11)
    assert emu.is_wide(270, 300, 10, INPLANE) == 1
    r = emu.run(imgs, masks, 10, 255, INPLANE)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=10, force2D=False)) == 1


def test_emu_wide_mode_literal_and_smooth(emu):
    H, W = 260, 256
    yy, xx = np.mgrid[:H, :W]
    img = np.clip(120 + 50 * np.sin(xx / 23.0) + 40 * np.cos(yy / 17.0), 0, 255).astype(np.uint8)[None]
    mask = np.zeros((1, H, W), np.uint8)
    mask[0, 5:250, 3:251] = 255
    assert emu.is_wide(H, W, 25, LITERAL) == 1
    r = emu.run(img, mask, 25, 255, LITERAL)
    assert compare_with_oracle(r, img, mask, dict(label=255, binWidth=25, force2D=True)) == 1


ALL_CLASSES = ("shape2D",) + tuple(orc.CLASS_ORDER)


def test_emu_shape2d_102_features(emu):
    # the reference's full vector: 9 shape2D + 93 = 102 (dataset.py:42), shape keys first
    imgs, masks = synth.make_patches(3, 48, 40, seed=5)
    masks[2, 10:14, 10:14] = 0  # a hole in the ROI
    r = emu.run(imgs, masks, 10, 255, LITERAL, classes=ALL_CLASSES)
    assert r["features"].shape[1] == 102
    s = dict(label=255, binWidth=10, force2D=True)
    assert compare_with_oracle(r, imgs, masks, s, classes=ALL_CLASSES) == 3


def test_emu_shape2d_edge_cases(emu):
    imgs, masks = edge_case_batch()
    r = emu.run(imgs, masks, 10, 255, INPLANE, classes=ALL_CLASSES)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=10, force2D=False), classes=ALL_CLASSES) == 6


def test_emu_bgr_front_end_matches_cv2(emu):
    # RadiomicExtractor.py:29-30,41-47: gray = cv2.cvtColor(BGR2GRAY), R/G/B = channels 2/1/0
    import ctypes

    import cv2

    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, (3, 37, 53, 3)).astype(np.uint8)
    planes = np.zeros((3, 4, 37, 53), np.uint8)
    emu.lib.radb_emu_bgr_planes.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]
    emu.lib.radb_emu_bgr_planes(bgr.ctypes.data, planes.ctypes.data, 3, 37 * 53)
    for i in range(3):
        np.testing.assert_array_equal(planes[i, 0], cv2.cvtColor(bgr[i], cv2.COLOR_BGR2GRAY))
        for k, ch in ((1, 2), (2, 1), (3, 0)):
            np.testing.assert_array_equal(planes[i, k], bgr[i, :, :, ch])


def test_emu_mask_resize_matches_cv2(emu):
    # RadiomicExtractor.py:34-35: cv2.resize(mask, (W, H), interpolation=cv2.INTER_NEAREST)
    import ctypes

    import cv2

    rng = np.random.default_rng(3)
    emu.lib.radb_emu_resize_mask.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_int, ctypes.c_int]
    sizes = [((450, 600), (450, 600)), ((24, 32), (48, 64)), ((450, 600), (768, 1024)), ((1, 1), (5, 7)), ((7, 5), (1, 1)),
             ((333, 500), (450, 600)), ((1000, 1500), (450, 600))]
    sizes += [(tuple(rng.integers(1, 300, 2)), tuple(rng.integers(1, 300, 2))) for _ in range(40)]
    for (sh, sw), (dh, dw) in sizes:
        src = rng.integers(0, 256, (2, int(sh), int(sw))).astype(np.uint8)
        dst = np.zeros((2, int(dh), int(dw)), np.uint8)
        emu.lib.radb_emu_resize_mask(src.ctypes.data, 2, int(sh), int(sw), dst.ctypes.data, int(dh), int(dw))
        for i in range(2):
            np.testing.assert_array_equal(dst[i], cv2.resize(src[i], (int(dw), int(dh)), interpolation=cv2.INTER_NEAREST))


@pytest.mark.parametrize("dt", [np.uint16, np.float32, np.float64])
def test_emu_other_pixel_types(emu, dt):
    """uint16 / float pixels (derived image types, BASELINE.json configs[4]): per-pixel fp64 binning,
    first-order features from the raw values (radix-select percentiles)."""
    rng = np.random.default_rng(4)
    g, masks = synth.make_patches(3, 40, 36, seed=6)
    if dt == np.uint16:
        imgs = (g.astype(np.uint16) * 7 + rng.integers(0, 7, g.shape).astype(np.uint16))     # range ~[0, 1800)
        bw = 64
    else:
        imgs = (np.sqrt(g.astype(np.float64)) * 11.3 - 40.0 + rng.normal(0, 0.3, g.shape)).astype(dt)  # negatives too
        bw = 7.5
    # a flat patch.  (A dyadic value: numpy's mean of N copies is then exact.  For other values upstream's
    # Skewness/Kurtosis of a flat float ROI are rounding noise, +-1 or 0; the engine returns the documented 0.)
    imgs[2] = 37.25 if dt != np.uint16 else 640
    r = emu.run(imgs, masks, bw, 255, INPLANE, max_ng=40)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=bw, force2D=False)) == 3


def test_emu_derived_image_types(emu):
    # Square / SquareRoot / Logarithm / Exponential (params.yml:141-144): transform, then the f64 engine path
    import ctypes

    imgs, masks = synth.make_patches(2, 40, 44, seed=9)
    emu.lib.radb_emu_derive.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
    for code, name in ((1, "Square"), (2, "SquareRoot"), (3, "Logarithm"), (4, "Exponential")):
        out = np.zeros(imgs.shape, np.float64)
        emu.lib.radb_emu_derive(imgs.ctypes.data, 2, 40 * 44, code, out.ctypes.data)
        for b in range(2):
            ref, prefix = orc.derived_image(imgs[b], name)
            np.testing.assert_allclose(out[b], ref, rtol=1e-14, atol=1e-13)
            assert prefix == name.lower()
        r = emu.run(out, masks, 10, 255, LITERAL, max_ng=32)
        assert compare_with_oracle(r, out, masks, dict(label=255, binWidth=10, force2D=True)) == 2


def test_emu_address_sanitizer(tmp_path):
    """The kernel source, emulated thread for thread, under AddressSanitizer: out-of-bounds accesses to
    the global buffers / the shared-memory arena abort the run (stand-in for compute-sanitizer memcheck,
    which is closed on the GPU pool)."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not available")
    so = str(tmp_path / "libradb_emu_asan.so")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++20", "-fPIC", "-shared", "-pthread", "-fsanitize=address,alignment",
                           "-fno-sanitize-recover=alignment", "-fno-omit-frame-pointer", "-o", so, os.path.join(root, "tests", "emu", "radb_emu.cpp")])
    ubsan = subprocess.run(["gcc", "-print-file-name=libubsan.so"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ, LD_PRELOAD=asan + (":" + ubsan if os.path.isabs(ubsan) else ""), ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "emu", "asan_cases.py"), so], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-2000:]
    assert "u16 [0 0]" in r.stdout


@pytest.mark.parametrize("hw,dt,angles,force2d", [((64, 64), np.uint8, INPLANE, False), ((16, 12), np.uint8, INPLANE, False),
                                                  ((13, 7), np.uint8, LITERAL, True), ((24, 24), np.float32, INPLANE, False)])
def test_emu_packed_masks_give_identical_rows(emu, hw, dt, angles, force2d):
    """radb_extract_packed: the kernels read 1-bit masks directly; rows and every matrix must be bit-identical to the
    byte-mask call (TMA-size aligned and ragged sizes, vectorised and generic paths, non-uint8 first-order)."""
    H, W = hw
    if hw == (16, 12):
        imgs, masks = edge_case_batch()
    else:
        imgs, masks = synth.make_patches(3, H, W, seed=21)
    kw = dict(max_ng=64) if dt != np.uint8 else {}
    imgs = imgs.astype(dt)
    a = emu.run(imgs, masks, 10, 255, angles, classes=("shape2D",) + tuple(orc.CLASS_ORDER), **kw)
    b = emu.run(imgs, masks, 10, 255, angles, classes=("shape2D",) + tuple(orc.CLASS_ORDER), packed=True, **kw)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert (a["status"] == 0).any()


@pytest.mark.parametrize("bw,ng_cap,size", [(32, 64, 64), (8, 256, 48), (16, 128, 32)])
def test_emu_lanczos_mcc_matches_oracle_and_dense_path(emu, monkeypatch, bw, ng_cap, size):
    """More than 40 gray levels: radb_mcc_lanczos_kernel (deflated Lanczos over the CSR non-zeros, one CTA per
    (patch, angle)) replaces the dense Householder tridiagonalisation of the warp-per-angle kernel.  Both must match
    the oracle's eigvals-based MCC; RADB_NO_LANCZOS forces the dense path (different arithmetic, same value)."""
    g, masks = synth.make_patches(3, size, seed=41, dtype=np.uint16, vmax=2047)
    g[2] = (np.indices((size, size)).sum(0) * 37 % 2048).astype(np.uint16)  # diagonal stripes: few distinct pairs
    r = emu.run(g, masks, bw, 255, INPLANE, max_ng=ng_cap)
    s = dict(label=255, binWidth=bw, force2D=False)
    assert compare_with_oracle(r, g, masks, s) == 3
    monkeypatch.setenv("RADB_NO_LANCZOS", "1")
    r2 = emu.run(g, masks, bw, 255, INPLANE, max_ng=ng_cap)
    assert compare_with_oracle(r2, g, masks, s, check_matrices=False) == 3
    k = orc.feature_names().index("original_glcm_MCC")
    np.testing.assert_allclose(r["features"][:, k], r2["features"][:, k], rtol=1e-9, atol=1e-11)
    assert not np.array_equal(r["features"][:, k], r2["features"][:, k])  # really two code paths


def test_emu_lanczos_mcc_edge_spectra(emu):
    """Edge spectra through the Lanczos path (max_ng 64 selects it): a checkerboard (lambda = -1: the spectral radius
    of the deflated operator is |lambda_min|), a flat ROI (one level: MCC = 1 by definition), two level groups that
    never touch (lambda = 1 twice: MCC = 1), a 2x2 ROI, sparse ROIs with empty angles."""
    imgs, masks = edge_case_batch(H=20, W=20)
    imgs = imgs.astype(np.uint16) * 4
    yy, xx = np.mgrid[:20, :20]
    imgs[6] = np.where(xx < 10, 100 + 40 * ((yy + xx) % 2), 900 + 40 * (yy % 2))  # two blocks of levels ...
    masks[6] = 255
    masks[6, :, 9:11] = 0                                                           # ... separated by a gap
    r = emu.run(imgs, masks, 16, 255, INPLANE, max_ng=64)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=16, force2D=False)) == 6
    k = orc.feature_names().index("original_glcm_MCC")
    assert r["features"][6, k] == pytest.approx(1.0, abs=1e-9) and r["features"][5, k] == pytest.approx(1.0, abs=1e-9)


@pytest.mark.parametrize("hw", [(16, 20), (9, 13), (64, 64)])
def test_emu_filtered_image_types_bit_exact(emu, hw):
    """Gradient / LoG / Wavelet kernels (radb_filters.cuh) against oracle/image_filters.py: same operation order, no
    fused multiply-adds -> bit-identical pixels (even and odd sizes: the wavelet pads odd sizes by a wrapped sample)."""
    import ctypes

    from oracle import image_filters as flt

    H, W = hw
    rng = np.random.default_rng(7)
    n = 3
    imgs = rng.integers(0, 256, (n, H, W)).astype(np.uint8)
    imgs[1] = synth.make_patches(1, H, W, seed=3)[0][0]
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    emu.lib.radb_emu_filter.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    scratch = np.zeros(n * 2 * (H + 1) * (W + 1) + n * H * W * 2, np.float64)
    out = np.zeros((n, H, W), np.float32)
    assert emu.lib.radb_emu_filter(p(imgs), n, H, W, 5, 0.0, 0, p(out), p(scratch)) == 0
    for b in range(n):
        np.testing.assert_array_equal(out[b].astype(np.float64), flt.gradient_image(imgs[b]))
    for sigma in (1.0, 2.0, 3.0):
        assert emu.lib.radb_emu_filter(p(imgs), n, H, W, 6, sigma, 0, p(out), p(scratch)) == 0
        for b in range(n):
            np.testing.assert_array_equal(out[b].astype(np.float64), flt.log_image(imgs[b], sigma))
    w4 = np.zeros((n, 4, H, W))
    assert emu.lib.radb_emu_filter(p(imgs), n, H, W, 7, 0.0, 0, p(w4), p(scratch)) == 0
    w2 = np.zeros((n, 2, H, W))
    assert emu.lib.radb_emu_filter(p(imgs), n, H, W, 7, 0.0, 1, p(w2), p(scratch)) == 0
    for b in range(n):
        for k, (name, arr) in enumerate(flt.wavelet_images(imgs[b]).items()):
            np.testing.assert_array_equal(w4[b, k], arr, err_msg=name)
        for k, (name, arr) in enumerate(flt.wavelet_images(imgs[b], force2D=True).items()):
            np.testing.assert_array_equal(w2[b, k], arr, err_msg=name)


def test_emu_first_order_only_stage(emu):
    """Only ``firstorder`` enabled: the build kernel stops after the histogram stage (no level image, no texture
    matrices); the 18 features must still match the oracle, edge cases included."""
    imgs, masks = synth.make_patches(3, 64, seed=8)
    e, em_ = edge_case_batch(H=64, W=64)
    imgs, masks = np.concatenate([imgs, e]), np.concatenate([masks, em_])
    r = emu.run(imgs, masks, 25, 255, INPLANE, classes=("firstorder",), matrices=False)  # no debug dumps: the short path
    s = dict(label=255, binWidth=25, force2D=False)
    assert r["features"].shape[1] == 18 and not r["levels"].any()
    n = 0
    for b in range(len(imgs)):
        try:
            ref = orc.execute(imgs[b], masks[b], s, classes=("firstorder",))
        except ValueError:
            assert np.isnan(r["features"][b]).all() and r["status"][b] != 0
            continue
        n += 1
        np.testing.assert_allclose(r["features"][b], list(ref.values()), rtol=1e-6, atol=1e-9)
    assert n >= 8


@pytest.mark.parametrize("bw", [25, 10])
def test_emu_word_pass_stress_patterns(emu, bw):
    """The 4-pixels-per-thread neighbourhood pass (and its request-queue overflow path) on adversarial patterns."""
    imgs, masks = word_pass_stress_batch(32, 32)
    r = emu.run(imgs, masks, bw, 255, INPLANE)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=bw, force2D=False)) == len(imgs)


def test_emu_one_angle_mid_levels_packs_four_patches_per_mcc_warp(emu):
    """Literal force2D (one angle) at binWidth 10 (15-40 gray levels): radb_mcc_g8_kernel packs four (patch, angle)
    tasks per warp; 6 patches cross a warp boundary."""
    imgs, masks = synth.make_patches(6, 32, seed=5)
    r = emu.run(imgs, masks, 10, 255, LITERAL)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=10, force2D=True)) == 6
