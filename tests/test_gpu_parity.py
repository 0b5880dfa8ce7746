"""GPU parity tests proper: every call goes through the C-ABI library (libradb_b200.so) and is
compared with the oracle on the same seeded inputs -- bit-exact for the discretised image and
every integer matrix, rtol 1e-6 / atol 1e-9 (BASELINE.json north_star) for each feature."""
import os

import numpy as np
import pytest
import torch

from oracle import cmatrices, radiomics_oracle as orc
from tests.emu_runner import ATOL, RTOL, compare_with_oracle, edge_case_batch, word_pass_stress_batch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
INPLANE = orc.angles(2)[0]
LITERAL = orc.angles(2, force2D=True)[0]


def _engine(pkg, bw=10, angles=INPLANE, sym=True, alpha=0, classes=None):
    return pkg.Engine(bw, 255, angles, sym, alpha, 0.0, classes or pkg.CLASS_ORDER)


def _dbg(eng, imgs, masks):
    return eng.debug_matrices(torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda())


def test_native_library_is_loaded(gpu_pkg):
    eng = _engine(gpu_pkg)
    assert eng.F == 93 and eng.names == orc.feature_names()
    maps = open("/proc/self/maps").read()
    assert "libradb_b200.so" in maps


def test_synthetic_inplane_bw10(gpu_pkg):
    imgs, masks = gpu_pkg.synth.make_patches(24, 64, seed=0)
    eng = _engine(gpu_pkg)
    l0 = eng.launches
    r = _dbg(eng, imgs, masks)
    assert eng.launches == l0 + 5  # build, MCC (8 lanes per angle), angle + misc (thread-level), misc residual: one chunk
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=10, force2D=False)) == 24


def test_literal_force2d_bw25(gpu_pkg):
    imgs, masks = gpu_pkg.synth.make_patches(12, 64, seed=1)
    r = _dbg(_engine(gpu_pkg, 25, LITERAL), imgs, masks)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=25, force2D=True)) == 12


def test_literal_force2d_bw10_mid_levels(gpu_pkg):
    """One angle at 15-40 gray levels: the MCC kernel packs four (patch, angle) tasks per warp (37 patches: ragged tail)."""
    imgs, masks = gpu_pkg.synth.make_patches(37, 64, seed=3)
    r = _dbg(_engine(gpu_pkg, 10, LITERAL), imgs, masks)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=10, force2D=True)) == 37


def test_nonsymmetric_glcm_alpha1(gpu_pkg):
    imgs, masks = gpu_pkg.synth.make_patches(8, 32, seed=2)
    r = _dbg(_engine(gpu_pkg, 16, INPLANE, False, 1), imgs, masks)
    s = dict(label=255, binWidth=16, force2D=False, symmetricalGLCM=False, gldm_a=1)
    assert compare_with_oracle(r, imgs, masks, s) == 8


def test_edge_cases(gpu_pkg):
    imgs, masks = edge_case_batch()
    r = _dbg(_engine(gpu_pkg), imgs, masks)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=10, force2D=False)) == 6
    assert list(r["status"][:3]) == [1, 2, 3]


@pytest.mark.parametrize("hw", [(7, 13), (33, 47), (5, 96), (128, 128)])
def test_ragged_sizes(gpu_pkg, hw):
    H, W = hw
    imgs, masks = gpu_pkg.synth.make_patches(4, H, W, seed=4)
    r = _dbg(_engine(gpu_pkg), imgs, masks)
    compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=10, force2D=False))


def test_force2d_dimension1_column_only(gpu_pkg):
    imgs, masks = gpu_pkg.synth.make_patches(6, 48, 40, seed=3)
    ang = orc.angles(2, force2D=True, force2Ddimension=1)[0]
    r = _dbg(_engine(gpu_pkg, 10, ang), imgs, masks)
    s = dict(label=255, binWidth=10, force2D=True, force2Ddimension=1)
    assert compare_with_oracle(r, imgs, masks, s) == 6


@pytest.mark.parametrize("hw,n", [((224, 224), 3), ((450, 600), 1), ((300, 200), 2)])
def test_wide_mode_large_images(gpu_pkg, hw, n):
    """Images that do not fit shared memory (224x224 patches of BASELINE.json configs[3]; whole
    600x450 dermoscopy images as the reference feeds them, RadiomicExtractor.py:29-38)."""
    H, W = hw
    imgs, masks = gpu_pkg.synth.make_patches(n, H, W, seed=12)
    r = _dbg(_engine(gpu_pkg), imgs, masks)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=10, force2D=False)) == n


def test_wide_mode_literal_whole_image(gpu_pkg):
    imgs, masks = gpu_pkg.synth.make_patches(2, 450, 600, seed=13)
    r = _dbg(_engine(gpu_pkg, 10, LITERAL), imgs, masks)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=10, force2D=True)) == 2


ALL_CLASSES = ("shape2D",) + tuple(orc.CLASS_ORDER)


@pytest.mark.parametrize("hw", [(64, 64), (37, 53), (300, 260)])
def test_shape2d_102_features(gpu_pkg, hw):
    """9 shape2D + 93 = 102 features per execute (dataset.py:42), shape keys first."""
    imgs, masks = gpu_pkg.synth.make_patches(4, hw[0], hw[1], seed=14)
    masks[1, 10:14, 10:15] = 0  # a hole
    eng = _engine(gpu_pkg, 10, LITERAL, classes=ALL_CLASSES)
    assert eng.F == 102 and eng.names == orc.feature_names(ALL_CLASSES)
    r = _dbg(eng, imgs, masks)
    s = dict(label=255, binWidth=10, force2D=True)
    assert compare_with_oracle(r, imgs, masks, s, classes=ALL_CLASSES) == 4


def test_shape2d_edge_cases(gpu_pkg):
    imgs, masks = edge_case_batch()
    r = _dbg(_engine(gpu_pkg, classes=ALL_CLASSES), imgs, masks)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=10, force2D=False), classes=ALL_CLASSES) == 6


def test_random_noise_many_levels(gpu_pkg):
    rng = np.random.default_rng(7)
    imgs = rng.integers(0, 256, (6, 40, 40)).astype(np.uint8)
    masks = np.where(rng.random((6, 40, 40)) < 0.8, 255, 0).astype(np.uint8)
    r = _dbg(_engine(gpu_pkg, 4), imgs, masks)  # 64 gray levels
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=4, force2D=False)) == 6


@pytest.mark.parametrize("bw,hw", [(25, 64), (10, 64), (25, 32)])
def test_word_pass_stress_patterns(gpu_pkg, bw, hw):
    """The 4-pixels-per-thread neighbourhood pass (and its request-queue overflow path) on adversarial patterns:
    checkerboards, stripes, noise masks, bounding boxes at every column offset inside a word."""
    imgs, masks = word_pass_stress_batch(hw, hw)
    eng = _engine(gpu_pkg, bw)
    r = _dbg(eng, imgs, masks)
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=bw, force2D=False)) == len(imgs)
    # the production call (no debug buffers) runs the compile-time specialised build kernel where the configuration
    # allows it (radb_kernels.cuh: FAST); the debug call above always runs the generic instance: identical rows
    out, st = eng.extract_device(torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda())
    assert np.array_equal(out.cpu().numpy().view(np.int64), r["features"].view(np.int64))
    assert not st.cpu().numpy().any()


@pytest.mark.parametrize("bw", [25, 10, 7.5])
def test_specialised_build_kernel_equals_generic_instance(gpu_pkg, bw):
    """2 048 synthetic patches + the edge cases: production rows (FAST instance when eligible: integer binWidth) equal the
    rows of the generic instance (debug call) bit for bit; 64 of them are checked against the oracle."""
    imgs, masks = gpu_pkg.synth.make_patches(2048, 64, seed=21)
    eimg, emsk = edge_case_batch()
    if eimg.shape[1:] == imgs.shape[1:]:
        imgs, masks = np.concatenate([imgs, eimg]), np.concatenate([masks, emsk])
    eng = _engine(gpu_pkg, bw)
    d_img, d_msk = torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda()
    out, st = eng.extract_device(d_img, d_msk)
    r = eng.debug_matrices(d_img, d_msk)
    got, ref = out.cpu().numpy(), r["features"]
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(np.nan_to_num(got).view(np.int64), np.nan_to_num(ref).view(np.int64))
    assert np.array_equal(st.cpu().numpy(), r["status"])
    sub = np.linspace(0, len(imgs) - 1, 64).astype(int)
    rs = {k: v[sub] for k, v in r.items()}
    compare_with_oracle(rs, imgs[sub], masks[sub], dict(label=255, binWidth=bw, force2D=False))


def test_smooth_image_large_zones(gpu_pkg):
    H = W = 64
    yy, xx = np.mgrid[:H, :W]
    img = np.clip(100 + 40 * np.sin(xx / 9.0) + 30 * np.cos(yy / 7.0), 0, 255).astype(np.uint8)[None]
    mask = np.full((1, H, W), 255, np.uint8)
    r = _dbg(_engine(gpu_pkg), img, mask)
    assert compare_with_oracle(r, img, mask, dict(label=255, binWidth=10, force2D=False)) == 1


def test_golden_fixture(gpu_pkg):
    z = np.load(os.path.join(GOLD, "oracle_features_seed0.npz"))
    for name, bw, ang in (("inplane_bw10", 10, INPLANE), ("literal_bw10", 10, LITERAL), ("inplane_bw25", 25, INPLANE)):
        eng = _engine(gpu_pkg, bw, ang)
        out, st = eng.extract_device(torch.as_tensor(z["images"]).cuda(), torch.as_tensor(z["masks"]).cuda())
        np.testing.assert_allclose(out.cpu().numpy(), z[name], rtol=RTOL, atol=ATOL)


def test_class_subset_columns(gpu_pkg):
    imgs, masks = gpu_pkg.synth.make_patches(6, 64, seed=5)
    d_i, d_m = torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda()
    full, _ = _engine(gpu_pkg, 25).extract_device(d_i, d_m)
    sub_eng = _engine(gpu_pkg, 25, classes=("firstorder", "glcm"))  # BASELINE.json configs[0]
    sub, _ = sub_eng.extract_device(d_i, d_m)
    assert sub_eng.F == 42
    assert torch.equal(sub, full[:, :42])


def test_host_pipeline_matches_device_path(gpu_pkg):
    imgs, masks = gpu_pkg.synth.make_patches(13, 32, seed=6)
    ex = gpu_pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": 10}}, chunk=5)
    dev, st = ex.extract_batch(torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda())
    host, st_h = ex.extract_batch(imgs, masks)
    np.testing.assert_array_equal(host, dev.cpu().numpy())
    assert ex.feature_names == orc.feature_names()
    with pytest.raises(ValueError, match="not present"):
        ex.extract_batch(imgs, np.zeros_like(masks), strict=True)


def test_host_pipeline_overlapping_streams(gpu_pkg):
    # two pipeline slots = two CUDA streams in flight on one handle: each stream owns its records
    B = 40000
    imgs, masks = gpu_pkg.synth.make_patches_torch(B, 64, seed=7, device="cuda")
    ex = gpu_pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": 25}}, chunk=4096)
    dev, _ = ex.extract_batch(imgs, masks)
    host, st = ex.extract_batch(imgs.cpu().numpy(), masks.cpu().numpy())
    assert not st.any()
    np.testing.assert_array_equal(host, dev.cpu().numpy())
    ex.pipeline.ramp = True  # graded chunk sizes (HostPipeline.chunk_schedule): same rows
    assert len(set(ex.pipeline.chunk_schedule(B, 4096, True))) > 3
    host2, st = ex.extract_batch(imgs.cpu().numpy(), masks.cpu().numpy())
    np.testing.assert_array_equal(host2, dev.cpu().numpy())


def test_record_path_matches_reference_call_pattern(gpu_pkg, tmp_path):
    """RadiomicExtractor.py:23-55 end to end: PNG files -> cv2 -> gray/R/G/B executes."""
    import cv2

    rng = np.random.default_rng(9)
    recs = []
    for k in range(3):
        g, m = gpu_pkg.synth.make_patches(1, *((48, 64) if k < 2 else (300, 400)), seed=20 + k)  # last: whole-image size
        bgr = np.stack([np.clip(g[0].astype(int) + rng.integers(-20, 20, g[0].shape), 0, 255) for _ in range(3)],
                       -1).astype(np.uint8)
        ip, sp = str(tmp_path / ("img%d.png" % k)), str(tmp_path / ("seg%d.png" % k))
        cv2.imwrite(ip, bgr)
        cv2.imwrite(sp, m[0] if k != 1 else cv2.resize(m[0], (32, 24), interpolation=cv2.INTER_NEAREST))
        recs.append({"image_path": ip, "segmentation_path": sp})
    params = {"setting": {"label": 255, "binWidth": 10, "force2D": True, "symmetricalGLCM": True,
                          "additionalInfo": False},
              "imageType": {"Original": {}},
              "featureClass": {c: [] for c in ("firstorder", "shape2D", "glcm", "gldm", "glrlm", "glszm", "ngtdm")}}
    ex = gpu_pkg.RadiomicsExtractor(params)
    res = ex.parallell_extraction(recs)
    ser = ex.serial_extraction(recs)
    assert len(res) == 3 and list(res[0].keys()) == ["grayscale", "red", "green", "blue"]
    for k, rec in enumerate(recs):
        im = cv2.imread(rec["image_path"], cv2.IMREAD_COLOR)
        sg = cv2.imread(rec["segmentation_path"], cv2.IMREAD_GRAYSCALE)
        if im.shape[:2] != sg.shape[:2]:
            sg = cv2.resize(sg, (im.shape[1], im.shape[0]), interpolation=cv2.INTER_NEAREST)
        planes = {"grayscale": cv2.cvtColor(im, cv2.COLOR_BGR2GRAY), "red": im[:, :, 2], "green": im[:, :, 1],
                  "blue": im[:, :, 0]}
        for ch, arr in planes.items():
            ref = orc.execute(arr, sg, params["setting"], classes=ALL_CLASSES, matrix_backend=cmatrices)
            assert list(res[k][ch].keys()) == list(ref.keys()) and len(ref) == 102
            np.testing.assert_allclose(list(res[k][ch].values()), list(ref.values()), rtol=RTOL, atol=ATOL)
            assert res[k][ch] == ser[k][ch]
    df = gpu_pkg.features_to_dataframe(res)
    assert df.shape == (3, 4 * 102)
    # streamed in windows (decode of window k+1 under the GPU work of window k): same rows, same order
    win = ex.parallell_extraction(recs * 3, n_processes=2, window=2)
    assert len(win) == 9 and all(win[i] == res[i % 3] for i in range(9))


def test_mask_resize_on_device_matches_cv2(gpu_pkg):
    """RadiomicExtractor.py:34-35 (cv2.resize INTER_NEAREST of the mask) on the device: bit-exact."""
    import cv2

    rng = np.random.default_rng(5)
    eng = gpu_pkg.Engine(25, 255, INPLANE)
    for (sh, sw), (dh, dw) in [((450, 600), (768, 1024)), ((24, 32), (48, 64)), ((333, 500), (450, 600)), ((1000, 1500), (450, 600)),
                               ((1, 1), (5, 7)), ((97, 13), (31, 211))]:
        src = rng.integers(0, 256, (3, sh, sw)).astype(np.uint8)
        got = eng.resize_mask(torch.as_tensor(src).cuda(), (dh, dw)).cpu().numpy()
        for i in range(3):
            np.testing.assert_array_equal(got[i], cv2.resize(src[i], (dw, dh), interpolation=cv2.INTER_NEAREST))


def test_full_size_properties(gpu_pkg):
    """BASELINE.json configs[1] size (100k 64x64 patches): size-independent properties."""
    B = 100000
    imgs, masks = gpu_pkg.synth.make_patches_torch(B, 64, seed=1234, device="cuda")
    ex = gpu_pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": 25}})
    out, st = ex.extract_batch(imgs, masks)
    out2, _ = ex.extract_batch(imgs, masks)
    torch.cuda.synchronize()
    assert int(st.sum()) == 0 and not torch.isnan(out).any()
    assert torch.equal(out, out2)                                   # deterministic (idempotent re-run)
    perm = torch.randperm(B, device="cuda", generator=torch.Generator("cuda").manual_seed(1))[:20000]
    outp, _ = ex.extract_batch(imgs[perm].contiguous(), masks[perm].contiguous())
    assert torch.equal(outp, out[perm])                             # a patch's row does not depend on its slot
    n = {k: i for i, k in enumerate(ex.feature_names)}
    o = out.cpu().numpy()
    npx = masks.view(B, -1).ne(0).sum(1).cpu().numpy()
    fo = lambda k: o[:, n["original_firstorder_" + k]]
    assert (fo("Minimum") <= fo("10Percentile")).all() and (fo("10Percentile") <= fo("Median")).all()
    assert (fo("Median") <= fo("90Percentile")).all() and (fo("90Percentile") <= fo("Maximum")).all()
    np.testing.assert_allclose(fo("RootMeanSquared") ** 2 * npx, fo("Energy"), rtol=1e-12)
    np.testing.assert_allclose(fo("Variance") + fo("Mean") ** 2, fo("RootMeanSquared") ** 2, rtol=1e-9)
    # GLDM and first-order share the level histogram: sum_i p_i^2 = Uniformity = GLDM GLN / Np
    np.testing.assert_allclose(o[:, n["original_gldm_GrayLevelNonUniformity"]] / npx, fo("Uniformity"), rtol=1e-12)
    assert (o[:, n["original_glrlm_RunPercentage"]] <= 1 + 1e-12).all()
    assert (o[:, n["original_glszm_ZonePercentage"]] <= 1 + 1e-12).all()
    mcc = o[:, n["original_glcm_MCC"]]
    assert (mcc >= 0).all() and (mcc <= 1 + 1e-9).all()
    np.testing.assert_allclose(o[:, n["original_glcm_SumAverage"]], 2 * o[:, n["original_glcm_JointAverage"]], rtol=1e-12)
    # spot-check 48 rows of the big batch against the oracle
    idx = np.random.default_rng(0).choice(B, 48, replace=False)
    hi, hm = imgs[idx].cpu().numpy(), masks[idx].cpu().numpy()
    for k, b in enumerate(idx):
        ref = list(orc.execute(hi[k], hm[k], dict(label=255, binWidth=25), matrix_backend=cmatrices).values())
        np.testing.assert_allclose(o[b], ref, rtol=RTOL, atol=ATOL)


def test_dataframe_feeds_the_feature_selection_stage(gpu_pkg, tmp_path):
    """Contract of the live consumer, /root/reference/reduce_dim.py:81-128: the pickled frame must be
    all-numeric, have 4 equal channel blocks with _gs/_red/_green/_blue suffixes, and survive the
    variance filter -> z-score -> L1-logistic CV selection -> |corr| > 0.95 drop sequence."""
    import pandas as pd
    from sklearn.feature_selection import VarianceThreshold
    from sklearn.linear_model import LogisticRegressionCV
    from sklearn.model_selection import StratifiedKFold
    from sklearn.preprocessing import StandardScaler

    n = 80
    ex = gpu_pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": 10, "force2D": True},
                                     "featureClass": {c: [] for c in ("firstorder", "shape2D", "glcm", "gldm", "glrlm",
                                                                      "glszm", "ngtdm")}})
    rng = np.random.default_rng(0)
    g, m = gpu_pkg.synth.make_patches(n, 64, seed=30)
    y = (m.reshape(n, -1).mean(1) > np.median(m.reshape(n, -1).mean(1))).astype(int)  # a label the features can learn
    results = []
    planes = [np.clip(g.astype(int) + rng.integers(-25, 25, g.shape), 0, 255).astype(np.uint8) for _ in range(4)]
    feats = [ex.extract_batch(pl, m, strict=True)[0] for pl in planes]
    for i in range(n):
        results.append({ch: dict(zip(ex.feature_names, feats[c][i])) for c, ch in
                        enumerate(("grayscale", "red", "green", "blue"))})
    df = gpu_pkg.features_to_dataframe(results)
    path = tmp_path / "radiomics.pkl"
    df.to_pickle(path)
    df = pd.read_pickle(path)
    assert len(df.columns) % 4 == 0 and len(df.columns) // 4 == 102
    for sfx in ("_gs", "_red", "_green", "_blue"):
        assert sum(sfx in c for c in df.columns) == 102
    assert np.isfinite(df.to_numpy()).all()
    tr, te = df.iloc[:60], df.iloc[60:]
    sel = VarianceThreshold(1e-3).fit(tr)
    tr = pd.DataFrame(sel.transform(tr), columns=df.columns[sel.get_support()])
    te = pd.DataFrame(sel.transform(te), columns=tr.columns)
    sc = StandardScaler().fit(tr)
    tr = pd.DataFrame(sc.transform(tr), columns=tr.columns)
    model = LogisticRegressionCV(Cs=np.logspace(-2, 1, 5), cv=StratifiedKFold(5, shuffle=True, random_state=42),
                                 penalty="l1", solver="liblinear", class_weight="balanced", scoring="f1",
                                 max_iter=2000).fit(tr, y[:60])
    # (SelectFromModel(model, prefit=True) trips over l1_ratio_=None in this image's sklearn; its rule for
    # L1 models is |coef| > 1e-5)
    keep = tr.columns[np.abs(model.coef_).max(0) > 1e-5]
    assert len(keep) >= 1
    corr = tr[keep].corr().abs()
    upper = corr.where(np.triu(np.ones(corr.shape), k=1).astype(bool))
    dropped = [c for c in upper.columns if (upper[c] > 0.95).any()]
    assert len(keep) - len(dropped) >= 1


@pytest.mark.parametrize("dt", [np.uint16, np.float32, np.float64])
def test_other_pixel_types(gpu_pkg, dt):
    rng = np.random.default_rng(4)
    g, masks = gpu_pkg.synth.make_patches(6, 64, seed=6)
    if dt == np.uint16:
        imgs = (g.astype(np.uint16) * 7 + rng.integers(0, 7, g.shape).astype(np.uint16))
        bw = 64
    else:
        imgs = (np.sqrt(g.astype(np.float64)) * 11.3 - 40.0 + rng.normal(0, 0.3, g.shape)).astype(dt)
        bw = 7.5
    imgs[2] = 37.25 if dt != np.uint16 else 640
    eng = gpu_pkg.Engine(bw, 255, INPLANE, max_ng=40)
    tt = torch.as_tensor(imgs.view(np.int16) if dt == np.uint16 else imgs).cuda()
    if dt == np.uint16:
        tt = tt.view(torch.uint16)
    r = eng.debug_matrices(tt, torch.as_tensor(masks).cuda())
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=bw, force2D=False)) == 6
    # the drop-in class sizes an engine from the data range
    ex = gpu_pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": bw}})
    out, st = ex.extract_batch(tt, torch.as_tensor(masks).cuda())
    # the two engines are sized differently (max_ng 40 vs data range) and may take different reduction
    # kernels (warp-per-angle vs thread-per-angle): same features well inside the 1e-6 contract
    np.testing.assert_allclose(out.cpu().numpy(), r["features"], rtol=1e-8, atol=1e-12)


def test_float_wide_mode(gpu_pkg):
    g, masks = gpu_pkg.synth.make_patches(1, 300, 280, seed=8)
    imgs = np.log1p(g.astype(np.float64)) * 40.0
    eng = gpu_pkg.Engine(5.0, 255, LITERAL, max_ng=64)
    r = eng.debug_matrices(torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda())
    assert compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=5.0, force2D=True)) == 1


def test_derived_image_types_batch_and_record_path(gpu_pkg, tmp_path):
    """params.yml:137-145 with the point-wise image types: shape once, then a 93-feature block per type
    (original_, square_, squareroot_, logarithm_, exponential_), on tensors and on decoded records."""
    import cv2

    params = {"setting": {"label": 255, "binWidth": 10, "force2D": True},
              "imageType": {"Original": {}, "Square": {}, "SquareRoot": {}, "Logarithm": {}, "Exponential": {}},
              "featureClass": {c: [] for c in ("firstorder", "shape2D", "glcm", "gldm", "glrlm", "glszm", "ngtdm")}}
    types = list(params["imageType"])
    ex = gpu_pkg.RadiomicsExtractor(params)
    assert len(ex.feature_names) == 102 + 4 * 93
    imgs, masks = gpu_pkg.synth.make_patches(3, 56, 48, seed=15)
    out, st = ex.extract_batch(torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda())
    out_h, _ = ex.extract_batch(imgs, masks)
    np.testing.assert_array_equal(out.cpu().numpy(), out_h)
    for b in range(3):
        ref = orc.execute_image_types(imgs[b], masks[b], params["setting"], classes=ALL_CLASSES, image_types=types,
                                      matrix_backend=cmatrices)
        assert list(ref.keys()) == ex.feature_names
        np.testing.assert_allclose(out_h[b], list(ref.values()), rtol=RTOL, atol=ATOL)
    # record path (cv2 files -> BGR front-end -> every image type of every plane)
    rng = np.random.default_rng(1)
    bgr = np.stack([np.clip(imgs[0].astype(int) + rng.integers(-20, 20, imgs[0].shape), 0, 255) for _ in range(3)],
                   -1).astype(np.uint8)
    ip, sp = str(tmp_path / "i.png"), str(tmp_path / "s.png")
    cv2.imwrite(ip, bgr)
    cv2.imwrite(sp, masks[0])
    res = ex.extract_radiomics({"image_path": ip, "segmentation_path": sp})
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    for ch, arr in (("grayscale", gray), ("blue", bgr[:, :, 0])):
        ref = orc.execute_image_types(arr, masks[0], params["setting"], classes=ALL_CLASSES, image_types=types,
                                      matrix_backend=cmatrices)
        assert list(res[ch].keys()) == list(ref.keys())
        np.testing.assert_allclose(list(res[ch].values()), list(ref.values()), rtol=RTOL, atol=ATOL)


REFERENCE_IMAGE_TYPES = {"Original": {}, "Wavelet": {}, "LoG": {"sigma": [1.0, 2.0, 3.0]}, "Square": {}, "SquareRoot": {},
                         "Logarithm": {}, "Exponential": {}, "Gradient": {}}  # /root/reference/params.yml:137-145


@pytest.mark.parametrize("force2d", [True, False])
def test_reference_image_type_list_incl_wavelet_log_gradient(gpu_pkg, tmp_path, force2d):
    """Every image type the reference's params.yml enables (Original, Wavelet, LoG sigma 1/2/3, Square, SquareRoot,
    Logarithm, Exponential, Gradient): 9 + 11 x 93 columns per channel under the literal force2D reading (wavelet-H,
    wavelet-L), 9 + 13 x 93 with the in-plane reading (wavelet-LH/HL/HH/LL); batched tensors and the record path
    against the oracle (oracle/image_filters.py restates ITK / PyWavelets; parity unpinned)."""
    import cv2

    params = {"setting": {"label": 255, "binWidth": 10, "force2D": force2d, "symmetricalGLCM": True, "additionalInfo": False},
              "imageType": REFERENCE_IMAGE_TYPES,
              "featureClass": {c: [] for c in ("firstorder", "shape2D", "glcm", "gldm", "glrlm", "glszm", "ngtdm")}}
    ex = gpu_pkg.RadiomicsExtractor(params)
    assert len(ex.feature_names) == 9 + (11 if force2d else 13) * 93 and not ex.skipped_image_types
    imgs, masks = gpu_pkg.synth.make_patches(2, 40, 52, seed=25)
    out, st = ex.extract_batch(torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda())
    out = out.cpu().numpy()
    assert not st.cpu().numpy().any()
    for b in range(2):
        ref = orc.execute_image_types(imgs[b], masks[b], params["setting"], classes=ALL_CLASSES,
                                      image_types=REFERENCE_IMAGE_TYPES, matrix_backend=cmatrices)
        assert list(ref.keys()) == ex.feature_names
        bad = [(k, r, g) for k, r, g in zip(ref, ref.values(), out[b]) if not np.isclose(g, r, rtol=RTOL, atol=ATOL, equal_nan=True)]
        assert not bad, bad[:5]
    if force2d:  # the reference's own call pattern: a record on disk, 4 channels
        rng = np.random.default_rng(2)
        bgr = np.stack([np.clip(imgs[0].astype(int) + rng.integers(-20, 20, imgs[0].shape), 0, 255) for _ in range(3)], -1).astype(np.uint8)
        ip, sp = str(tmp_path / "i.png"), str(tmp_path / "s.png")
        cv2.imwrite(ip, bgr)
        cv2.imwrite(sp, masks[0])
        res = ex.parallell_extraction([{"image_path": ip, "segmentation_path": sp}])[0]
        ref = orc.execute_image_types(bgr[:, :, 2], masks[0], params["setting"], classes=ALL_CLASSES,
                                      image_types=REFERENCE_IMAGE_TYPES, matrix_backend=cmatrices)
        assert list(res["red"].keys()) == list(ref.keys())
        np.testing.assert_allclose(list(res["red"].values()), list(ref.values()), rtol=RTOL, atol=ATOL)
        df = gpu_pkg.features_to_dataframe([res])
        assert df.shape == (1, 4 * (9 + 11 * 93)) and df.columns[9 + 93] == "wavelet-H_firstorder_10Percentile_gs"


def _fuzz_patch(rng, H, W, kind):
    yy, xx = np.mgrid[:H, :W]
    if kind == 0:      # white noise
        img = rng.integers(0, 256, (H, W))
    elif kind == 1:    # few levels, big zones
        img = (rng.integers(0, 4, (H // 4 + 1, W // 4 + 1)).repeat(4, 0).repeat(4, 1)[:H, :W]) * 60 + 20
    elif kind == 2:    # smooth gradient + noise
        img = 30 + 180 * (xx / max(W - 1, 1)) * (yy / max(H - 1, 1)) + rng.normal(0, 6, (H, W))
    elif kind == 3:    # stripes (runs along one direction only)
        img = np.where((xx // 3) % 2 == 0, 70, 150) + rng.integers(0, 3, (H, W))
    else:              # checkerboard of blocks
        img = np.where(((xx // 5) + (yy // 5)) % 2 == 0, 40, 210)
    img = np.clip(img, 0, 255).astype(np.uint8)
    mk = rng.integers(0, 4)
    if mk == 0:
        mask = np.full((H, W), 255)
    elif mk == 1:
        mask = np.where(rng.random((H, W)) < rng.uniform(0.2, 0.9), 255, 0)
    elif mk == 2:
        mask = np.where((yy - H / 2) ** 2 / (H / 2.2) ** 2 + (xx - W / 2) ** 2 / (W / 2.5) ** 2 < 1, 255, 0)
    else:
        mask = np.zeros((H, W), int)
        y0, x0 = rng.integers(0, H // 2), rng.integers(0, W // 2)
        mask[y0:y0 + rng.integers(2, H // 2 + 2), x0:x0 + rng.integers(2, W // 2 + 2)] = 255
    return img, mask.astype(np.uint8)


def test_fuzz_random_shapes_patterns_masks(gpu_pkg):
    """Randomised shapes / textures / masks / bin widths / angle sets: integer matrices bit-exact and every
    feature (102 per patch, shape2D included) within tolerance of the oracle."""
    # RADB_FUZZ_TRIALS / RADB_FUZZ_SEED: longer one-off campaigns (e.g. 300 trials) outside the regular suite
    trials = int(os.environ.get("RADB_FUZZ_TRIALS", "36"))
    rng = np.random.default_rng(int(os.environ.get("RADB_FUZZ_SEED", "2026")))
    total = 0
    for trial in range(trials):
        H, W = int(rng.integers(6, 90)), int(rng.integers(6, 90))
        bw = [4, 10, 25, 7.5][trial % 4]
        literal = trial % 3 == 0
        ang = LITERAL if literal else INPLANE
        pats = [_fuzz_patch(rng, H, W, int(rng.integers(0, 5))) for _ in range(5)]
        imgs = np.stack([p[0] for p in pats])
        masks = np.stack([p[1] for p in pats])
        if trial % 7 == 6:  # binCount binning instead of a fixed width
            n = int(rng.integers(1, 65))
            eng = gpu_pkg.Engine(25, 255, ang, True, 0, 0.0, ALL_CLASSES, bin_count=n)
            r = _dbg(eng, imgs, masks)
            total += compare_with_oracle(r, imgs, masks, dict(label=255, binCount=n, force2D=literal), classes=ALL_CLASSES)
            continue
        r = _dbg(_engine(gpu_pkg, bw, ang, classes=ALL_CLASSES), imgs, masks)
        total += compare_with_oracle(r, imgs, masks, dict(label=255, binWidth=bw, force2D=literal), classes=ALL_CLASSES)
    assert total > 3 * trials


def test_ragged_mixed_sizes_and_coverage(gpu_pkg):
    """BASELINE.json configs[3]: patch sizes 32 / 64 / 224 mixed 1:1:1 with mask coverage 2-100 % in ONE
    radb_extract_ragged call (narrow and wide kernels, TMA and non-TMA staging): rows in input order,
    equal to the dense entry point size class by size class, and equal to the oracle."""
    rng = np.random.default_rng(9)
    images, masks = [], []
    for i in range(18):
        H = (32, 64, 224)[i % 3]
        g, m = gpu_pkg.synth.make_patches(1, H, seed=100 + i)
        if i % 5 == 4:  # variable coverage: a random sub-mask of the lesion, down to a few percent
            m = np.where(rng.random(m.shape) < rng.uniform(0.02, 1.0), m, 0).astype(np.uint8)
        images.append(g[0])
        masks.append(m[0])
    images.append(rng.integers(0, 256, (37, 29)).astype(np.uint8))  # odd size: no TMA path
    masks.append(np.full((37, 29), 255, np.uint8))
    images.append(images[0].copy())
    masks.append(np.zeros((32, 32), np.uint8))                      # label absent
    ex = gpu_pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": 25}})
    out, status = ex.extract_list(images, masks)
    assert status[-1] == 1 and np.isnan(out[-1]).all() and not status[:-1].any()
    for H in (32, 64, 224):
        idx = [i for i in range(18) if images[i].shape[0] == H]
        d, st = ex.extract_batch(torch.as_tensor(np.stack([images[i] for i in idx])).cuda(),
                                 torch.as_tensor(np.stack([masks[i] for i in idx])).cuda())
        assert np.array_equal(d.cpu().numpy(), out[idx])
    s = orc.resolve_settings(dict(label=255, binWidth=25, force2D=False))
    names = orc.feature_names(orc.CLASS_ORDER)
    for i in (0, 1, 4, 9, 18):  # one per size class incl. a thinned mask and the odd size (the 224s are slow on the CPU)
        f = orc.execute(images[i], masks[i], s, matrix_backend=cmatrices)
        np.testing.assert_allclose(out[i], [f[k] for k in names], rtol=RTOL, atol=ATOL)
    with pytest.raises(ValueError):
        ex.extract_list(images, masks, strict=True)


@pytest.mark.parametrize("bw", [8, 16, 32, 64])
def test_binwidth_sweep_uint16_up_to_256_levels(gpu_pkg, bw):
    """BASELINE.json configs[4]: binWidth 8..64 on uint16 intensities in [0, 2048) -> 256..32 gray levels.
    64 levels still fit shared memory; 128 and 256 run in big mode (GLCM + MCC workspace in global memory,
    u16 level image for 256)."""
    g, masks = gpu_pkg.synth.make_patches(3, 64, seed=41, dtype=np.uint16, vmax=2047)
    ng_cap = 2048 // bw
    eng = gpu_pkg.Engine(bw, 255, INPLANE, max_ng=ng_cap)
    tt = torch.as_tensor(g.view(np.int16)).cuda().view(torch.uint16)
    r = eng.debug_matrices(tt, torch.as_tensor(masks).cuda())
    assert compare_with_oracle(r, g, masks, dict(label=255, binWidth=bw, force2D=False)) == 3
    assert r["ng"].max() > ng_cap // 2
    ex = gpu_pkg.RadiomicsExtractor({"setting": {"label": 255, "binWidth": bw}})
    out, st = ex.extract_batch(tt, torch.as_tensor(masks).cuda())
    np.testing.assert_allclose(out.cpu().numpy(), r["features"], rtol=1e-8, atol=1e-12)


def test_multi_chunk_two_stream_pipeline_is_bit_identical(gpu_pkg):
    """Batches larger than one chunk are cut into equal chunks and pipelined over two streams (build of chunk
    i+1 overlaps the reductions of chunk i, two workspace slots).  Same rows, bit for bit, as one chunk --
    dense, ragged, with shape2D, and with invalid ROIs in the batch."""
    imgs, masks = gpu_pkg.synth.make_patches(61, 32, seed=77)
    masks[7] = 0
    masks[40] = 0
    masks[40, 5, 5] = 255
    ti, tm = torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda()
    eng = gpu_pkg.Engine(25, 255, INPLANE, classes=gpu_pkg.CLASS_ORDER + ("shape2D",))
    ref, rst = eng.extract_device(ti, tm)
    torch.cuda.synchronize()
    l0 = eng.launches
    for chunk in (4, 8, 20, 60):
        eng.set_chunk(chunk)
        for _ in range(2):
            out, st = eng.extract_device(ti, tm)
            assert torch.equal(st, rst)
            assert np.array_equal(out.cpu().numpy(), ref.cpu().numpy(), equal_nan=True)
    assert eng.launches >= l0 + 2 * 5 * (16 + 8 + 4 + 2)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):  # a caller stream other than the default one
        out, st = eng.extract_device(ti, tm, stream=side)
    side.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref.cpu().numpy(), equal_nan=True)
    lst_i = [imgs[i] for i in range(61)] + [np.ascontiguousarray(imgs[3][:24, :20])] * 9
    lst_m = [masks[i] for i in range(61)] + [np.ascontiguousarray(masks[3][:24, :20])] * 9
    ip, mp, io, mo, hw = gpu_pkg.pack_ragged(lst_i, lst_m)
    eng.set_chunk(0)
    r0, s0 = eng.extract_ragged(torch.as_tensor(ip).cuda(), torch.as_tensor(mp).cuda(), io, mo, hw)
    eng.set_chunk(8)
    r1, s1 = eng.extract_ragged(torch.as_tensor(ip).cuda(), torch.as_tensor(mp).cuda(), io, mo, hw)
    assert torch.equal(s0, s1) and np.array_equal(r0.cpu().numpy(), r1.cpu().numpy(), equal_nan=True)
    assert np.array_equal(r0[:61].cpu().numpy(), ref.cpu().numpy(), equal_nan=True)


def test_packed_mask_transfer_path(gpu_pkg):
    """Host pipeline with masks packed to 1 bit per pixel on the host and expanded on the device: same rows, bit
    for bit, as the plain uint8 transfer; masks with other labels and odd sizes included."""
    imgs, masks = gpu_pkg.synth.make_patches(37, 33, 29, seed=5)   # 957 pixels per patch: not a multiple of 8 / 16
    masks[3][masks[3] == 255] = 7          # another label only: ROI absent
    masks[4, :5] = 128                      # foreign label next to the ROI
    eng = gpu_pkg.Engine(25, 255, INPLANE)
    plain = gpu_pkg.HostPipeline(eng, chunk=10, pack_masks=False)
    packed = gpu_pkg.HostPipeline(eng, chunk=10, pack_masks=True, pack_threads=3)
    o0, s0 = plain.run(imgs, masks)
    o1, s1 = packed.run(imgs, masks)
    assert torch.equal(s0, s1) and s0[3] == 1
    assert np.array_equal(o0.numpy(), o1.numpy(), equal_nan=True)
    assert packed.h2d_bytes < plain.h2d_bytes * 0.6
    d = torch.empty((37, 33, 29), dtype=torch.uint8, device="cuda")
    pk = torch.as_tensor(np.packbits(masks.reshape(-1) == 255, bitorder="little")).cuda()
    eng.unpack_mask(pk, d)
    assert torch.equal(d.cpu() == 255, torch.as_tensor(masks == 255)) and set(d.unique().tolist()) <= {0, 255}


@pytest.mark.parametrize("dt", [np.uint8, np.float32])
def test_bin_count_binning(gpu_pkg, dt):
    """binCount (pyradiomics imageoperations.getBinEdges): n bins over the ROI range, numpy.histogram edges."""
    g, masks = gpu_pkg.synth.make_patches(6, 40, 36, seed=21)
    g[1] = (g[1] // 16) * 16
    if dt == np.uint8:
        g[2][masks[2] == 255] = 90
    imgs = g if dt == np.uint8 else (np.sqrt(g.astype(np.float64)) * 3.0 - 7.0).astype(np.float32)
    for n in (1, 16, 100):
        eng = gpu_pkg.Engine(25, 255, INPLANE, bin_count=n)
        r = eng.debug_matrices(torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda())
        assert compare_with_oracle(r, imgs, masks, dict(label=255, binCount=n, force2D=False)) == 6
    ex = gpu_pkg.RadiomicsExtractor({"setting": {"label": 255, "binCount": 16}})
    out, st = ex.extract_batch(torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda())
    eng = gpu_pkg.Engine(25, 255, INPLANE, bin_count=16)
    r = eng.debug_matrices(torch.as_tensor(imgs).cuda(), torch.as_tensor(masks).cuda())
    np.testing.assert_allclose(out.cpu().numpy(), r["features"], rtol=1e-8, atol=1e-12)
