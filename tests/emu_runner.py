"""Runs the emulated kernel (tests/emu) and compares any engine result with the oracle.
Shared by the CPU (emulation) and GPU parity tests so that both read the same way."""
from __future__ import annotations

import ctypes

import numpy as np

from multimodal_isic_b200 import _abi
from oracle import cmatrices, radiomics_oracle as orc


class EmuRunner:
    def __init__(self, so_path):
        self.lib = ctypes.CDLL(so_path)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        self.lib.radb_emu_extract.argtypes = [vp, vp, i32, vp, i64, i32, i32, i64, i64] + [vp] * 10
        self.lib.radb_emu_last_error.restype = ctypes.c_char_p

    def run_ragged(self, images, masks, bin_width=10, label=255, angles=((0, 1),), classes=_abi.CLASS_ORDER, max_ng=0):
        """Host twin of radb_extract_ragged on a list of variable-size (image, mask) pairs."""
        from multimodal_isic_b200.engine import pack_ragged

        ip, mp, io, mo, hw = pack_ragged(images, masks)
        dtype = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2, np.dtype(np.float64): 3}[ip.dtype]
        s = _abi.make_settings(bin_width, label, angles, classes=classes, max_ng=max_ng)
        F = self.lib.radb_emu_feature_count(ctypes.byref(s))
        n = len(hw)
        out, status = np.zeros((n, F)), np.zeros(n, np.int32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self.lib.radb_emu_extract_ragged.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                     ctypes.c_int64] + [ctypes.c_void_p] * 5
        rc = self.lib.radb_emu_extract_ragged(ctypes.byref(s), p(ip), dtype, p(mp), n, p(io), p(mo), p(hw), p(out), p(status))
        assert rc == 0, self.lib.radb_emu_last_error()
        return out, status

    def is_wide(self, H, W, bin_width=10, angles=((0, 1),)):
        s = _abi.make_settings(bin_width, 255, angles)
        return self.lib.radb_emu_is_wide(ctypes.byref(s), H, W)

    def run(self, imgs, masks, bin_width=10, label=255, angles=((0, 1),), symmetrical=True, alpha=0,
            classes=_abi.CLASS_ORDER, max_ng=0, bin_count=0, packed=False, matrices=True):
        """``packed``: hand the masks over bit-packed (radb_extract_packed's layout) instead of as bytes."""
        imgs = np.ascontiguousarray(imgs)
        dtype = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2, np.dtype(np.float64): 3}[imgs.dtype]
        masks = np.ascontiguousarray(masks, dtype=np.uint8)
        B, H, W = imgs.shape
        s = _abi.make_settings(bin_width, label, angles, symmetrical, alpha, classes=classes, max_ng=max_ng, bin_count=bin_count)
        F = self.lib.radb_emu_feature_count(ctypes.byref(s))
        ng = self.lib.radb_emu_max_ng(ctypes.byref(s))
        assert F > 0 and ng > 0, self.lib.radb_emu_last_error()
        na, nr = len(angles), max(H, W)
        r = dict(features=np.zeros((B, F)), status=np.zeros(B, np.int32), levels=np.zeros((B, H, W), np.int32),
                 glcm=np.zeros((B, na, ng, ng), np.int32), glrlm=np.zeros((B, na, ng, nr), np.int32),
                 glszm=np.zeros((B, ng, H * W), np.int32), gldm=np.zeros((B, ng, 2 * na + 1), np.int32),
                 ngtdm_n=np.zeros((B, ng), np.int32), ngtdm_s=np.zeros((B, ng)), ng=np.zeros(B, np.int32))
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        mstride = H * W
        if packed:
            masks = np.ascontiguousarray(np.packbits((masks == label).reshape(B, -1), axis=1, bitorder="little"))
            mstride = masks.shape[1]
        self.lib.radb_emu_set_mask_bits(int(packed))
        rc = self.lib.radb_emu_extract(ctypes.byref(s), p(imgs), dtype, p(masks), B, H, W, H * W * imgs.itemsize, mstride,
                                       *[p(r[k]) if (matrices or k in ("features", "status")) else None
                                         for k in ("features", "status", "levels", "glcm", "glrlm", "glszm",
                                                   "gldm", "ngtdm_n", "ngtdm_s", "ng")])
        self.lib.radb_emu_set_mask_bits(0)
        assert rc == 0, self.lib.radb_emu_last_error()
        return r


RTOL, ATOL = 1e-6, 1e-9  # BASELINE.json north_star: every floating-point feature


def compare_with_oracle(r, imgs, masks, settings, check_matrices=True, classes=orc.CLASS_ORDER):
    """Bit-exact discretised image + integer matrices, features within rtol 1e-6 / atol 1e-9.
    ``r``: dict from EmuRunner.run or Engine.debug_matrices.  Returns the number of valid patches."""
    s = orc.resolve_settings(settings)
    names = orc.feature_names(classes)
    nvalid = 0
    for b in range(len(imgs)):
        try:
            m = orc.matrices(imgs[b], masks[b], s, matrix_backend=cmatrices)
        except ValueError as e:
            msg = str(e)
            want = 1 if "not present" in msg else 2 if "1 segmented voxel" in msg else 3
            assert r["status"][b] == want, (b, msg, r["status"][b])
            assert np.isnan(r["features"][b]).all()
            continue
        nvalid += 1
        assert r["status"][b] == 0, (b, r["status"][b])
        Ng = m["Ng"]
        if check_matrices:
            assert r["ng"][b] == Ng
            np.testing.assert_array_equal(r["levels"][b], m["levels"])
            np.testing.assert_array_equal(r["glcm"][b][:, :Ng, :Ng].transpose(1, 2, 0), m["glcm"])
            assert r["glcm"][b].sum() == m["glcm"].sum()
            np.testing.assert_array_equal(r["glrlm"][b][:, :Ng, :].transpose(1, 2, 0), m["glrlm"])
            Ns = m["glszm"].shape[1]
            np.testing.assert_array_equal(r["glszm"][b][:Ng, :Ns], m["glszm"])
            assert r["glszm"][b].sum() == m["glszm"].sum()
            np.testing.assert_array_equal(r["gldm"][b][:Ng], m["gldm"])
            np.testing.assert_array_equal(r["ngtdm_n"][b][:Ng], m["ngtdm_n"])
            np.testing.assert_allclose(r["ngtdm_s"][b][:Ng], m["ngtdm_s"], rtol=1e-12, atol=1e-12)
        f = orc.execute(imgs[b], masks[b], s, classes=classes, matrix_backend=cmatrices)
        ref = np.array([f[k] for k in names])
        got = r["features"][b]
        bad = [(k, a, g) for k, a, g in zip(names, ref, got) if not np.isclose(g, a, rtol=RTOL, atol=ATOL, equal_nan=True)]
        assert not bad, (b, bad[:5])
    return nvalid


def edge_case_batch(H=16, W=12, seed=3):
    """Edge cases the reference's engine distinguishes: absent label, single voxel, 1-D ROI,
    flat ROI, full-patch ROI, two-level checkerboard, ROI touching every border, sparse ROI."""
    rng = np.random.default_rng(seed)
    n = 9
    imgs = rng.integers(0, 256, (n, H, W)).astype(np.uint8)
    masks = np.zeros((n, H, W), np.uint8)
    masks[0] = 0                                   # label absent
    masks[1, H // 2, W // 2] = 255                 # single voxel
    masks[2, 3, 2:W - 2] = 255                     # one row: too few dimensions
    masks[3, 2:H - 2, 2:W - 2] = 255
    imgs[3] = 130                                  # flat ROI (one gray level)
    masks[4] = 255                                 # whole patch
    yy, xx = np.mgrid[:H, :W]
    imgs[5] = np.where((yy + xx) % 2 == 0, 40, 200)
    masks[5, 1:H - 1, 1:W - 1] = 255               # checkerboard: lambda = -1 in the GLCM spectrum
    masks[6] = 255
    masks[6, 4:8, 3:7] = 0                         # hole + touches every border
    masks[7] = np.where(rng.random((H, W)) < 0.15, 255, 0)  # sparse, many isolated voxels
    masks[8, 0:2, 0:2] = 255                       # 2x2 ROI in a corner
    masks[7, 0, 0] = 128                           # other labels are not ROI
    return imgs, masks


def word_pass_stress_batch(H=64, W=64):
    """Patterns that stress the 4-pixel-word neighbourhood pass of the build kernel: checkerboards (two union requests per
    pixel: the request queue's overflow path), diagonal stripes, one-pixel runs, bounding boxes that start at every
    column offset inside a word, ragged masks, isolated pixels.  uint8, two to eleven levels at binWidth 25."""
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[:H, :W]
    imgs, masks = [], []

    def add(img, mask):
        imgs.append(np.asarray(img, np.uint8))
        masks.append(np.where(mask, 255, 0).astype(np.uint8))

    full = np.ones((H, W), bool)
    add(40 + 100 * ((xx + yy) & 1), full)                       # checkerboard: NW / NE unions everywhere
    add(40 + 60 * ((xx + yy) % 3), full)                        # diagonal stripes
    add(40 + 60 * ((xx - yy) % 3), full)                        # anti-diagonal stripes
    add(40 + 100 * (xx & 1), full)                              # vertical one-pixel stripes
    add(40 + 100 * (yy & 1), full)                              # horizontal stripes
    add(rng.integers(0, 256, (H, W)), rng.random((H, W)) < 0.5)  # noise image, noise mask: isolated pixels
    for off in range(4):                                        # bbox starting at x = 5 + off, width not a multiple of 4
        m = np.zeros((H, W), bool)
        m[3:H - 5, 5 + off:W - 6 - 2 * off] = True
        add(40 + 100 * ((xx + yy) & 1), m)
        add(rng.integers(0, 256, (H, W)), m & (rng.random((H, W)) < 0.9))
    m = np.zeros((H, W), bool)
    m[10:12, 7:9] = True                                        # 2 x 2 ROI inside one word pair
    add(rng.integers(0, 256, (H, W)), m)
    add(np.full((H, W), 77), full)                              # flat: one zone of H * W pixels
    return np.stack(imgs), np.stack(masks)
