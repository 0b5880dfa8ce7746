"""Host logic + the C-ABI library surface (no compute calls: no GPU needed)."""
import ctypes
import os
import re
import warnings

import numpy as np
import pytest
import torch
import yaml

import multimodal_isic_b200 as pkg
from multimodal_isic_b200 import _abi
from oracle import radiomics_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REFERENCE_LIKE_PARAMS = """
setting:
  additionalInfo: False
  label: 255
  binWidth: 10
  force2D: True
  symmetricalGLCM: True
imageType:
  Original: {}
  Wavelet: {}
  LoG:
    sigma: [1.0, 2.0, 3.0]
featureClass:
  firstorder: []
  shape2D: []
  glcm: []
  gldm: []
  glrlm: []
  glszm: []
  ngtdm: []
"""


def test_library_exports_every_declared_symbol(built):
    built.build_cuda()
    header = open(os.path.join(ROOT, "include", "radb.h")).read()
    declared = set(re.findall(r"\b(radb_[a-z_]+)\s*\(", header))
    assert declared == set(pkg.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    lib.radb_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.radb_version()


def test_settings_struct_layout_matches_header():
    # field order/types of radb_settings in include/radb.h <-> ctypes mirror
    header = open(os.path.join(ROOT, "include", "radb.h")).read()
    body = re.search(r"typedef struct radb_settings \{(.*?)\} radb_settings;", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:double|int32_t|uint32_t|int8_t)\s+([a-z_]+)", body)
    assert fields == [f[0] for f in _abi.RadbSettings._fields_]
    assert ctypes.sizeof(_abi.RadbSettings) == 72  # static_assert in csrc/radb_host.h


def test_no_cpu_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.RadbError, match="no CPU fallback"):
        pkg.RadiomicsExtractor({"setting": {"label": 255}})
    with pytest.raises(RuntimeError, match="missing"):
        _abi.load_library("/nonexistent/libradb_b200.so")


def test_settings_reference_like_file(tmp_path):
    f = tmp_path / "params.yml"
    f.write_text(REFERENCE_LIKE_PARAMS)
    s = pkg.Settings(str(f))  # strict (default): every enabled image type is implemented
    assert s.label == 255 and s.bin_width == 10.0 and not s.skipped_image_types
    assert s.angles() == orc.angles(2, force2D=True)[0] == [(0, 1)]          # literal force2D on 2-D input
    assert list(s.enabledImagetypes) == ["Original", "Wavelet", "LoG"]
    assert list(s.enabledFeatures) == ["firstorder", "shape2D", "glcm", "gldm", "glrlm", "glszm", "ngtdm"]
    # 102 = 9 shape2D + 93 (dataset.py:42), shape keys first; then one 93-block per filtered image
    n = s.feature_names()
    assert n[:102] == orc.feature_names(("shape2D",) + orc.CLASS_ORDER)
    prefixes = [b[0] for b in s.blocks]
    assert prefixes[:3] == ["original", "wavelet-H", "wavelet-L"] and all(p.startswith("log-sigma-") for p in prefixes[3:])
    assert len(n) == 9 + 93 * len(prefixes)


def test_settings_class_order_and_feature_subset():
    s = pkg.Settings({"setting": {"label": 1}, "featureClass": {"glcm": ["Contrast", "Idm"], "firstorder": None}})
    assert s.feature_names()[:2] == ["original_glcm_Contrast", "original_glcm_Idm"]
    eng_classes, perm = s.engine_columns()
    assert eng_classes == ["firstorder", "glcm"]
    eng = ["original_%s_%s" % (c, f) for c in eng_classes for f in pkg.FEATURE_NAMES[c]]
    assert [eng[i] for i in perm] == s.feature_names()
    assert pkg.Settings({"binWidth": 25, "force2D": False}).angles() == orc.angles(2)[0]
    with pytest.raises(ValueError):
        pkg.Settings({"featureClass": {"glcm": ["Homogeneity1"]}})
    assert pkg.Settings({"setting": {"binCount": 16}}).bin_count == 16 and pkg.Settings({"setting": {}}).bin_count == 0
    with pytest.raises(NotImplementedError):
        pkg.Settings({"setting": {"binCount": 1000}})
    with pytest.raises(NotImplementedError):
        pkg.Settings({"setting": {"weightingNorm": "euclidean"}})


def test_feature_name_tables_agree():
    assert {k: v for k, v in pkg.FEATURE_NAMES.items() if k != "shape2D"} == dict(orc.FEATURE_NAMES)
    assert pkg.FEATURE_NAMES["shape2D"] == orc.SHAPE2D_NAMES


def test_shape2d_on_2d_input_does_not_need_force2d():
    # pyradiomics featureextractor.computeShape: the force2D rule belongs to 3-D input; 2-D input (what the
    # reference passes, RadiomicExtractor.py:31,36) computes shape2D regardless
    s = pkg.Settings({"setting": {"force2D": False}, "featureClass": {"shape2D": [], "glcm": []}})
    assert s.classes == ["shape2D", "glcm"]
    s = pkg.Settings({"setting": {"force2D": True}, "featureClass": {"glcm": [], "shape2D": ["Perimeter"]}})
    assert s.feature_names()[0] == "original_shape2D_Perimeter" and s.engine_columns()[0] == ["shape2D", "glcm"]


def test_result_affecting_settings_are_rejected_not_ignored():
    for bad in ({"minimumROIDimensions": 1}, {"preCrop": True}, {"additionalInfo": True}, {"normalize": True}):
        with pytest.raises(NotImplementedError):
            pkg.Settings({"setting": bad})
    pkg.Settings({"setting": {"minimumROIDimensions": 2, "preCrop": False, "additionalInfo": False}})


def test_unimplemented_image_types_raise_by_default():
    cfg = {"imageType": {"Original": {}, "LBP2D": {}}, "setting": {"label": 255}}
    with pytest.raises(NotImplementedError, match="LBP2D"):
        pkg.Settings(cfg)
    with pytest.warns(RuntimeWarning, match="LBP2D"):
        s = pkg.Settings(cfg, strict=False)
    assert s.skipped_image_types == ["LBP2D"] and s.image_types == ["Original"]


def test_shard_bounds():
    assert pkg.shard_bounds([1] * 8, 4) == [0, 2, 4, 6, 8]
    b = pkg.shard_bounds([100, 1, 1, 1, 1, 100], 2)
    assert b[0] == 0 and b[-1] == 6 and 1 <= b[1] <= 5
    assert pkg.shard_bounds([], 3) == [0, 0, 0, 0]
    rng = np.random.default_rng(0)
    c = rng.uniform(1, 50, 1000)
    b = pkg.shard_bounds(c, 8)
    loads = [c[b[i]:b[i + 1]].sum() for i in range(8)]
    assert max(loads) / (sum(loads) / 8) < 1.05


def test_dataframe_contract():
    # extract_radiomics.py:54-71 + reduce_dim.py:88,97-100
    names = orc.feature_names()
    rec = {ch: dict(zip(names, np.arange(len(names), dtype=float))) for ch in ("grayscale", "red", "green", "blue")}
    df = pkg.features_to_dataframe([rec, rec, rec])
    assert df.shape == (3, 4 * 93) and len(df.columns) % 4 == 0
    assert df.columns[0] == names[0] + "_gs" and df.columns[93] == names[0] + "_red"
    assert df.columns[-1] == names[-1] + "_blue"
    assert all(str(t) == "float64" for t in df.dtypes)


def test_settings_derived_image_types():
    p = yaml.safe_load(REFERENCE_LIKE_PARAMS)
    p["imageType"] = {"Original": {}, "Square": {}, "SquareRoot": {}, "Logarithm": {}, "Exponential": {}, "Gradient": {}}
    s = pkg.Settings(p)
    assert s.image_types == ["Original", "Square", "SquareRoot", "Logarithm", "Exponential", "Gradient"]
    names = s.feature_names()
    assert len(names) == 102 + 5 * 93                      # shape once, one 93-block per image type
    assert names[:9] == ["original_shape2D_%s" % f for f in orc.SHAPE2D_NAMES]
    assert names[102] == "square_firstorder_10Percentile" and names[-1] == "gradient_ngtdm_Strength"
    assert s.engine_columns()[1] == list(range(len(names)))
    # wavelet band names follow pyradiomics' axis removal: force2D -> 1-D transform along x
    p["imageType"] = {"Wavelet": {}}
    assert [b[0] for b in pkg.Settings(p).blocks] == ["wavelet-H", "wavelet-L"]
    p["setting"]["force2D"] = False
    assert [b[0] for b in pkg.Settings(p).blocks] == ["wavelet-LH", "wavelet-HL", "wavelet-HH", "wavelet-LL"]
    p["imageType"] = {"Wavelet": {"wavelet": "db2"}}
    with pytest.raises(NotImplementedError):
        pkg.Settings(p)
    p["imageType"] = {"LoG": {}}
    with pytest.raises(ValueError):
        pkg.Settings(p)


def test_host_pipeline_chunk_schedule():
    """Graded chunk sizes of the host-to-host pipeline: every patch exactly once, in order, no chunk above the slot."""
    from multimodal_isic_b200.engine import HostPipeline

    for B in (0, 1, 100, 8191, 8192, 8193, 29695, 29696, 50000, 100000, 65543, 1000003):
        for chunk in (1, 63, 64, 1000, 8192):
            for ramp in (False, True):
                s = HostPipeline.chunk_schedule(B, chunk, ramp)
                assert sum(s) == B and all(0 < x <= chunk for x in s), (B, chunk, ramp)
    s = HostPipeline.chunk_schedule(100000, 8192, True)
    assert s[:3] == [2048, 4096, 8192] and s[-3:] == [4096, 2048, 1024]
    assert HostPipeline.chunk_schedule(100000, 8192, ramp=False) == [8192] * 12 + [1696]
    assert HostPipeline.chunk_schedule(100000, 8192, "4/4") == [2048] + [8192] * 11 + [5792, 2048]


def test_host_pipeline_packing_policy(monkeypatch):
    """Default hand-over policy per cores-per-rank (measured: profiles/r2_e2e_modes_{2,8}gpu_pool.jsonl): pack from 3 cores
    per rank, per-chunk adaptive below 6 on several ranks; explicit arguments win."""
    import os

    from multimodal_isic_b200.engine import HostPipeline

    def make(cpus, ranks, **kw):
        monkeypatch.setattr(os, "cpu_count", lambda: cpus)
        monkeypatch.setenv("LOCAL_WORLD_SIZE", str(ranks))
        return HostPipeline(None, **kw)

    p = make(16, 1)
    assert (p.pack_masks, p.adaptive, p.pack_threads, p.chunk, p.ramp) == (True, False, 12, 4096, False)
    p = make(24, 2)
    assert (p.pack_masks, p.adaptive, p.pack_threads) == (True, False, 9)
    p = make(32, 8)
    assert (p.pack_masks, p.adaptive, p.pack_threads) == (True, True, 4)
    p = make(8, 8)
    assert (p.pack_masks, p.pack_threads) == (False, 1)
    p = make(32, 8, pack_masks=False, adaptive=False, pack_threads=2, ramp="4/4")
    assert (p.pack_masks, p.adaptive, p.pack_threads, p.ramp) == (False, False, 2, "4/4")


def test_pack_ragged_layout():
    """Host packing for radb_extract_ragged: 16-byte aligned patch starts, (H, W) table, lossless pools."""
    from multimodal_isic_b200 import pack_ragged

    rng = np.random.default_rng(0)
    shapes = [(5, 7), (16, 16), (3, 3), (16, 16)]
    imgs = [rng.integers(0, 65535, s).astype(np.uint16) for s in shapes]
    msks = [(rng.random(s) < 0.5).astype(np.uint8) * 255 for s in shapes]
    ip, mp, io, mo, hw = pack_ragged(imgs, msks)
    assert ip.dtype == np.uint16 and mp.dtype == np.uint8 and hw.tolist() == [list(s) for s in shapes]
    assert (io % 16 == 0).all() and (mo % 16 == 0).all()
    for i, s in enumerate(shapes):
        k = s[0] * s[1]
        assert np.array_equal(ip[io[i] // 2: io[i] // 2 + k].reshape(s), imgs[i])
        assert np.array_equal(mp[mo[i]: mo[i] + k].reshape(s), msks[i])
    with pytest.raises(ValueError):
        pack_ragged(imgs, msks[:-1])
    with pytest.raises(ValueError):
        pack_ragged([imgs[0]], [msks[1]])


def test_pack_mask_host_matches_packbits():
    """radb_pack_mask_host (host half of the packed-mask transfer path): bit i <-> mask[i] == label, any length,
    any thread count, labels outside uint8 give an empty ROI.  Pure host code: runs without a GPU."""
    import ctypes

    lib = pkg.load_library()
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 31, 32, 33, 100003, 3_000_001):
        m = rng.choice(np.array([0, 255, 128, 1], np.uint8), size=n, p=[0.5, 0.4, 0.05, 0.05])
        for label in (255, 1, 0):
            want = np.packbits(m == label, bitorder="little")
            for th in (1, 3, 8):
                out = np.full((n + 7) // 8, 0xAA, np.uint8)
                assert lib.radb_pack_mask_host(m.ctypes.data, n, label, out.ctypes.data, th) == 0
                assert np.array_equal(out, want), (n, label, th)
        out = np.full((n + 7) // 8, 0xAA, np.uint8)
        assert lib.radb_pack_mask_host(m.ctypes.data, n, 256, out.ctypes.data, 2) == 0 and not out.any()
    assert lib.radb_pack_mask_host(None, 8, 255, None, 1) != 0


def test_record_decode_pool_keeps_order(tmp_path):
    """parallell_extraction's host side (RadiomicExtractor.py:29-36,58-65): cv2 decode of image + mask on a thread
    pool, input order kept, masks as stored (the nearest-resize of :34-35 runs on the device), missing files raise."""
    import cv2

    from multimodal_isic_b200.extractor import RadiomicsExtractor

    rng = np.random.default_rng(1)
    recs, want = [], []
    for k in range(9):
        H, W = 20 + k, 31 - k
        bgr = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
        m = (rng.random((H, W)) < 0.5).astype(np.uint8) * 255
        ip, sp = str(tmp_path / ("i%d.png" % k)), str(tmp_path / ("s%d.png" % k))
        cv2.imwrite(ip, bgr)
        cv2.imwrite(sp, m if k != 4 else cv2.resize(m, (W // 2, H // 2), interpolation=cv2.INTER_NEAREST))
        recs.append({"image_path": ip, "segmentation_path": sp, "other": k})
        want.append((bgr, m))
    for workers in (None, 1, 4):
        got = RadiomicsExtractor._load_records(recs, workers)
        assert len(got) == 9
        for k, ((im, sg), (bgr, m)) in enumerate(zip(got, want)):
            assert np.array_equal(im, bgr)
            if k != 4:
                assert np.array_equal(sg, m)
            else:
                assert sg.shape == (m.shape[0] // 2, m.shape[1] // 2)
    with pytest.raises(FileNotFoundError):
        RadiomicsExtractor._load_records(recs + [{"image_path": str(tmp_path / "nope.png"), "segmentation_path": recs[0]["segmentation_path"]}], 3)


def test_settings_parse_the_reference_parameter_file():
    """The reference's own params.yml (RadiomicExtractor.py:15), when the checkout is present (it is not on the GPU
    box): label 255, binWidth 10, force2D -> one along-row angle; every enabled image type is implemented (strict
    parsing succeeds): 9 shape2D + 93 x 11 filtered images (Original, wavelet-H/L, LoG sigma 1/2/3, Square, SquareRoot,
    Logarithm, Exponential, Gradient) per channel."""
    path = "/root/reference/params.yml"
    if not os.path.exists(path):
        pytest.skip("reference checkout not present")
    s = pkg.Settings(path)
    assert s.label == 255 and s.bin_width == 10.0 and s.bin_count == 0 and s.angles() == [(0, 1)]
    assert list(s.enabledImagetypes) == ["Original", "Wavelet", "LoG", "Square", "SquareRoot", "Logarithm", "Exponential", "Gradient"]
    assert not s.skipped_image_types and s.image_types == list(s.enabledImagetypes)
    assert [b[0] for b in s.blocks] == ["original", "wavelet-H", "wavelet-L", "log-sigma-1-0-mm-3D", "log-sigma-2-0-mm-3D",
                                        "log-sigma-3-0-mm-3D", "square", "squareroot", "logarithm", "exponential", "gradient"]
    names = s.feature_names()
    assert len(names) == 9 + 93 * 11 and names[0].startswith("original_shape2D_") and names[9] == "original_firstorder_10Percentile"
    assert sum(n.startswith("exponential_") for n in names) == 93
    assert not any(n.startswith("diagnostics_") for n in names)  # additionalInfo: False (params.yml:62)
