"""Drop-in for ``/root/reference/RadiomicExtractor.py``: same class name, constructor argument
and method names (including the ``parallell_extraction`` spelling), same per-record result
(``{"grayscale","red","green","blue"}`` -> ordered mapping of pyradiomics feature names), so
``/root/reference/extract_radiomics.py:48-71`` runs unchanged on top of it.  The arithmetic
runs in the sm_100a CUDA engine (``csrc/``) through the C-ABI in ``include/radb.h``."""
from __future__ import annotations

import logging
import time
from collections import OrderedDict

import numpy as np
import torch

from . import _abi
from .engine import Engine, HostPipeline
from .settings import IMAGE_TYPE_CODES, Settings

logger = logging.getLogger(__name__)

CHANNELS = ("grayscale", "red", "green", "blue")


def _status_error(code, label):
    msg = _abi.STATUS_MESSAGES.get(int(code), "extraction failed (status %d)" % code)
    if code == 1:
        msg = msg % (label,)
    elif code == 3:
        msg = msg % (1, 2)
    return ValueError(msg)


class RadiomicsExtractor:
    """``RadiomicsExtractor(param_file)`` -- mirrors ``RadiomicExtractor.py:12-94``.

    ``param_file`` is the pyradiomics YAML path the reference passes (``'params.yml'``) or the
    equivalent dict.  Extra keyword-only arguments are B200-side knobs: ``device`` (CUDA
    ordinal), ``strict`` (default True: enabled-but-unimplemented image types / classes raise
    ``NotImplementedError``; False: they are logged at ERROR level, skipped and listed in
    ``skipped_image_types``), ``chunk`` (patches per pipelined H2D chunk), ``max_ng``."""

    def __init__(self, param_file, *, device=0, strict=True, chunk=4096, max_ng=0, **setting_overrides):
        self.params = Settings(param_file, strict=strict, **setting_overrides)
        self.skipped_image_types = list(self.params.skipped_image_types)
        self.device = int(device)
        # U1 (oracle/U1_ANGLES.md): say which angle reading is in effect -- the literal one for the reference's
        # own file (2-D arrays + force2D: True -> one along-row offset) is the unverified one
        ang = self.params.angles(2)
        self.angle_reading = ("literal force2D on a 2-D array: %d offset(s) %s" % (len(ang), ang)
                              if self.params.settings["force2D"] else "in-plane: %d offsets %s" % (len(ang), ang))
        logger.warning("radb angle semantics: %s (see oracle/U1_ANGLES.md)", self.angle_reading)
        eng_classes, self._perm = self.params.engine_columns()
        s = self.params.settings
        bc = self.params.bin_count
        self.engine = Engine(self.params.bin_width, self.params.label, self.params.angles(2),
                             bool(s["symmetricalGLCM"]), float(s["gldm_a"]), float(s["voxelArrayShift"]),
                             eng_classes, max_ng, self.device, bc)
        self._engine_args = (self.params.bin_width, self.params.label, self.params.angles(2),
                             bool(s["symmetricalGLCM"]), float(s["gldm_a"]), float(s["voxelArrayShift"]), eng_classes)
        self._engines_ng = {}  # engines for non-uint8 pixels, keyed by their gray-level bound
        # derived image types (params.yml:141-144) are float64 images in the original intensity range
        self.derived_types = [t for t in self.params.image_types if t != "Original"]
        self._has_original = "Original" in self.params.image_types
        self._derived_engine = None
        if any(t in IMAGE_TYPE_CODES for t in self.derived_types):
            tex_classes = [c for c in eng_classes if c != "shape2D"]
            self._derived_engine = Engine(self.params.bin_width, self.params.label, self.params.angles(2),
                                          bool(s["symmetricalGLCM"]), float(s["gldm_a"]), float(s["voxelArrayShift"]),
                                          tex_classes, min(255, int(255.0 // self.params.bin_width) + 3), self.device, bc)
        self.feature_names = self.params.feature_names()
        self._perm_t = None
        self._identity = self._perm == list(range(self.engine.F))
        self.pipeline = HostPipeline(self.engine, chunk)

    # ---- reference API -------------------------------------------------------------------
    def get_enabled_image_types(self):  # RadiomicExtractor.py:17-18
        return list(self.params.enabledImagetypes.keys())

    def get_enabled_features(self):  # RadiomicExtractor.py:20-21
        return list(self.params.enabledFeatures.keys())

    @staticmethod
    def _load_record(record):
        """cv2 decode exactly as RadiomicExtractor.py:29,33 (host side): interleaved BGR image and the mask as
        stored.  The nearest-neighbour resize of a mask whose size differs from the image's (:34-35) and the gray /
        R / G / B planes of :30,41-47 are produced on the GPU (``radb_resize_mask``, the front-end kernel: both
        bit-exact with cv2)."""
        import cv2

        im = cv2.imread(record["image_path"], cv2.IMREAD_COLOR)
        if im is None:
            raise FileNotFoundError(record["image_path"])
        sg = cv2.imread(record["segmentation_path"], cv2.IMREAD_GRAYSCALE)
        if sg is None:
            raise FileNotFoundError(record["segmentation_path"])
        return np.ascontiguousarray(im), np.ascontiguousarray(sg)

    def _extract_records(self, loaded, max_bytes=256 << 20):
        """loaded: list of (bgr [H,W,3], mask [H,W]) -> list of per-record channel dicts (input order)."""
        results = [None] * len(loaded)
        groups = {}
        for i, (im, sg) in enumerate(loaded):  # one launch per (image size, stored mask size)
            groups.setdefault(im.shape[:2] + sg.shape[:2], []).append(i)
        dev = torch.device("cuda", self.device)
        for (H, W, mh, mw), idxs in groups.items():
            per = max(1, max_bytes // (H * W * 3))
            for s0 in range(0, len(idxs), per):
                part = idxs[s0:s0 + per]
                bgr = torch.as_tensor(np.stack([loaded[i][0] for i in part])).to(dev, non_blocking=True)
                msk = torch.as_tensor(np.stack([loaded[i][1] for i in part])).to(dev, non_blocking=True)
                if (mh, mw) != (H, W):  # RadiomicExtractor.py:34-35
                    msk = self.engine.resize_mask(msk, (H, W))
                if self.derived_types or not self._has_original:
                    out, status, planes = self.engine.extract_bgr(bgr, msk, return_planes=True)
                    out, status = self._device_blocks(planes.view(-1, H, W), msk.repeat_interleave(4, dim=0),
                                                      first=(out, status))
                else:
                    out, status = self.engine.extract_bgr(bgr, msk)
                feats = self._permute(out).cpu().numpy()
                status = status.cpu().numpy()
                if status.any():  # the reference has no try/except: pyradiomics' ValueError aborts the run
                    raise _status_error(int(status[np.nonzero(status)[0][0]]), self.params.label)
                for k, i in enumerate(part):
                    results[i] = {ch: OrderedDict(zip(self.feature_names, feats[4 * k + c].tolist()))
                                  for c, ch in enumerate(CHANNELS)}
        return results

    def extract_radiomics(self, list_of_dicts):  # RadiomicExtractor.py:23-55 (one record)
        return self._extract_records([self._load_record(list_of_dicts)])[0]

    @staticmethod
    def _load_records(list_of_dicts, n_workers=None):
        """Decode all records, in input order.  The reference fans whole records over ``cpu_count() - 1``
        processes (RadiomicExtractor.py:60-65); here only the JPEG/PNG decode is host work, and cv2 releases
        the GIL, so a thread pool of the same width does it."""
        import os
        from concurrent.futures import ThreadPoolExecutor

        if n_workers is None:
            n_workers = max(1, (os.cpu_count() or 2) - 1)
        if n_workers <= 1 or len(list_of_dicts) <= 1:
            return [RadiomicsExtractor._load_record(r) for r in list_of_dicts]
        with ThreadPoolExecutor(int(n_workers)) as pool:
            return list(pool.map(RadiomicsExtractor._load_record, list_of_dicts))  # map() keeps the input order

    def parallell_extraction(self, list_of_dicts, n_processes=None, window=None):  # RadiomicExtractor.py:58-71
        """Order-preserving extraction of all records.  ``n_processes`` (default ``cpu_count() - 1``, as in the
        reference) is the width of the host decode pool; the feature fan-out is over GPU CTAs.  Records are streamed
        in windows of ``window`` records (default 8 per decode thread, at least 32): window k+1 is decoded on the
        pool while window k is on the GPU, so host memory holds two windows, not the whole data set (the reference
        streams records through ``pool.imap``).  Records of equal image size inside a window share a launch."""
        import os
        from concurrent.futures import ThreadPoolExecutor

        logger.info("Extraction mode: parallel")
        t0 = time.time()
        n_workers = int(n_processes) if n_processes else max(1, (os.cpu_count() or 2) - 1)
        window = int(window) if window else max(32, 8 * n_workers)
        records = list(list_of_dicts)
        results = []
        if len(records) <= window or n_workers <= 1:
            results = self._extract_records(self._load_records(records, n_workers))
        else:
            with ThreadPoolExecutor(n_workers) as pool:
                def submit(lo):
                    return [pool.submit(self._load_record, r) for r in records[lo:lo + window]]

                pending = submit(0)
                for lo in range(0, len(records), window):
                    loaded = [f.result() for f in pending]  # input order; a decode error propagates as in the reference
                    pending = submit(lo + window) if lo + window < len(records) else []
                    results.extend(self._extract_records(loaded))
        h, m, s = self._convert_time(t0, time.time())
        logger.info(f" Time taken: {h}h:{m}m:{s}s")
        return results

    def serial_extraction(self, list_of_dicts):  # RadiomicExtractor.py:74-85
        logger.info("Extraction mode: serial")
        t0 = time.time()
        all_results = [self.extract_radiomics(r) for r in list_of_dicts]
        h, m, s = self._convert_time(t0, time.time())
        logger.info(f" Time taken: {h}h:{m}m:{s}s")
        return all_results

    def _convert_time(self, start_time, end_time):  # RadiomicExtractor.py:88-94
        dt = end_time - start_time
        return int(dt // 3600), int((dt % 3600) // 60), int(dt % 60)

    # ---- batched entry points --------------------------------------------------------------
    def _permute(self, out):
        if self._identity:
            return out
        if self._perm_t is None or self._perm_t.device != out.device:
            self._perm_t = torch.as_tensor(self._perm, device=out.device)
        return out.index_select(1, self._perm_t)

    def _engine_for(self, images, masks, texture_only=False):
        """uint8 pixels: the engine sized from binWidth.  Other pixel types: the gray-level count depends
        on the data, so an engine sized for this batch's largest ROI range (rounded up to a multiple of
        8 levels) is created on demand and cached.  ``texture_only``: without shape2D (filtered image types)."""
        t = torch.as_tensor(images)
        if (t.dtype == torch.uint8 or self.params.bin_count) and not texture_only:  # binCount: Ng = binCount whatever the pixel type
            return self.engine, self.pipeline
        if self.params.bin_count:
            ng = int(self.params.bin_count)
        else:
            m = torch.as_tensor(masks) == self.params.label
            f = t.to(torch.float64) if t.dtype != torch.uint16 else t.to(torch.int32).to(torch.float64)
            big = torch.finfo(torch.float64).max
            lo = torch.where(m, f, torch.full_like(f, big)).flatten(1).amin(1)
            hi = torch.where(m, f, torch.full_like(f, -big)).flatten(1).amax(1)
            ok = hi >= lo
            span = float(((hi - lo)[ok] / self.params.bin_width).max().item()) if bool(ok.any()) else 0.0
            ng = min(256, (int(span) + 3 + 7) // 8 * 8)
        key = (ng, texture_only)
        if key not in self._engines_ng:
            args = list(self._engine_args)
            if texture_only:
                args[6] = [c for c in args[6] if c != "shape2D"]
            eng = Engine(*args, max_ng=ng, device=self.device, bin_count=self.params.bin_count)
            self._engines_ng[key] = (eng, HostPipeline(eng, self.pipeline.chunk))
        return self._engines_ng[key]

    def extract_batch(self, images, masks, strict=False):
        """``images`` ``[B, H, W]`` uint8 (the reference's cv2 planes), uint16, float32 or float64; ``masks``
        ``[B, H, W]`` uint8.  CUDA tensors -> CUDA tensors
        ``(features [B, F] float64, status [B] int32)``, asynchronous on the current stream;
        host arrays -> NumPy arrays through the pinned, chunk-pipelined path.
        ``strict=True`` raises the ValueError pyradiomics would raise for an invalid ROI
        (the reference has no try/except, RadiomicExtractor.py:23-55); otherwise such rows are NaN."""
        if self.derived_types or not self._has_original:
            return self._extract_multi_type(images, masks, strict)
        engine, pipeline = self._engine_for(images, masks)
        if isinstance(images, torch.Tensor) and images.is_cuda:
            out, status = engine.extract_device(images, masks)
            out = self._permute(out)
            if strict:
                bad = torch.nonzero(status)
                if bad.numel():
                    raise _status_error(int(status[bad[0, 0]]), self.params.label)
            return out, status
        out, status = pipeline.run(images, masks)
        out = self._permute(out).numpy()
        status = status.numpy()
        if strict and status.any():
            raise _status_error(int(status[np.nonzero(status)[0][0]]), self.params.label)
        return out, status


    def extract_list(self, images, masks, strict=False):
        """Variable-size batch: ``images`` / ``masks`` are sequences of 2-D arrays (one dtype; each mask has
        its image's shape) -- the shape the reference's records have after decoding (RadiomicExtractor.py:
        29-36: whole images of differing sizes).  One call packs them into two pools, copies them to the
        device and runs ``radb_extract_ragged`` (patches of equal size share a launch).  Returns NumPy
        ``(features [n, F], status [n])`` in input order."""
        if self.derived_types or not self._has_original:
            raise NotImplementedError("extract_list handles the Original image type")
        from .engine import pack_ragged

        img_pool, mask_pool, img_off, mask_off, hw = pack_ragged(images, masks)
        engine = self.engine
        if img_pool.dtype != np.uint8 and not self.params.bin_count:
            engine, _ = self._engine_for_pool(images, masks)
        dev = torch.device("cuda", self.device)
        ip = torch.as_tensor(img_pool.view(np.int16) if img_pool.dtype == np.uint16 else img_pool).to(dev)
        if img_pool.dtype == np.uint16:
            ip = ip.view(torch.uint16)
        out, status = engine.extract_ragged(ip, torch.as_tensor(mask_pool).to(dev), img_off, mask_off, hw)
        out = self._permute(out).cpu().numpy()
        status = status.cpu().numpy()
        if strict and status.any():
            raise _status_error(int(status[np.nonzero(status)[0][0]]), self.params.label)
        return out, status

    def _engine_for_pool(self, images, masks):
        """Non-uint8 ragged batches: an engine sized for the largest ROI range of the batch (see _engine_for)."""
        span = 0.0
        for im, mk in zip(images, masks):
            roi = np.asarray(im)[np.asarray(mk) == self.params.label]
            if roi.size:
                span = max(span, float(roi.max()) - float(roi.min()))
        ng = min(256, (int(span / self.params.bin_width) + 3 + 7) // 8 * 8)
        key = (ng, False)
        if key not in self._engines_ng:
            eng = Engine(*self._engine_args, max_ng=ng, device=self.device)
            self._engines_ng[key] = (eng, HostPipeline(eng, self.pipeline.chunk))
        return self._engines_ng[key]

    # ---- several image types ------------------------------------------------------------------
    def _device_blocks(self, images, masks, first=None):
        """[shape | one block per filtered image, in pyradiomics' order] for uint8 device images; ``first`` =
        precomputed (out, status) of the Original/shape engine."""
        if images.dtype != torch.uint8:
            raise NotImplementedError("derived image types are implemented for uint8 input images")
        out0, status = first if first is not None else self.engine.extract_device(images, masks)
        nshape = 9 if "shape2D" in self.params.classes else 0
        blocks = [out0[:, :nshape]] if nshape else []  # shape descriptors come first, whatever the image-type order
        wave = None
        for name, t, arg in self.params.blocks:
            if t == "Original":
                blocks.append(out0[:, nshape:])
                continue
            if t in IMAGE_TYPE_CODES:  # point-wise types: float64 images in the original intensity range
                o, st = self._derived_engine.extract_device(self.engine.derive_image(images, IMAGE_TYPE_CODES[t]), masks)
            else:
                if t == "Wavelet":
                    if wave is None:  # one transform yields every band
                        wave = self.engine.filter_image(images, "Wavelet", x_only=bool(self.params.settings["force2D"]))
                    img = wave[:, arg].contiguous()
                elif t == "LoG":
                    img = self.engine.filter_image(images, "LoG", sigma=arg)
                else:
                    img = self.engine.filter_image(images, "Gradient")
                # the gray-level range of a filtered image depends on the data: an engine sized for this batch
                eng, _ = self._engine_for(img, masks, texture_only=True)
                o, st = eng.extract_device(img, masks)
            blocks.append(o)
            status = torch.maximum(status, st)
        return torch.cat(blocks, dim=1), status

    def _extract_multi_type(self, images, masks, strict):
        on_device = isinstance(images, torch.Tensor) and images.is_cuda
        dev = torch.device("cuda", self.device)
        if on_device:
            out, status = self._device_blocks(images, masks)
            out = self._permute(out)
        else:
            images, masks = torch.as_tensor(images), torch.as_tensor(masks)
            outs, sts = [], []
            for s0 in range(0, len(images), self.pipeline.chunk):
                o, st = self._device_blocks(images[s0:s0 + self.pipeline.chunk].to(dev),
                                            masks[s0:s0 + self.pipeline.chunk].to(dev))
                outs.append(self._permute(o).cpu())
                sts.append(st.cpu())
            out, status = torch.cat(outs), torch.cat(sts)
        if strict and bool((status != 0).any()):
            raise _status_error(int(status[status != 0][0]), self.params.label)
        return (out, status) if on_device else (out.numpy(), status.numpy())


def features_to_dataframe(results, suffixes=("gs", "red", "green", "blue")):
    """``extract_radiomics.py:54-71``: 4 per-channel frames concatenated on axis 1 with
    ``_gs/_red/_green/_blue`` suffixes, as plain float64 columns."""
    import pandas as pd

    frames = [pd.DataFrame([item[ch] for item in results]) for ch in CHANNELS]
    df = pd.concat(frames, axis=1)
    n = len(df.columns) // 4
    df.columns = [f"{col}_{sfx}" for col, sfx in zip(df.columns, sum(([s] * n for s in suffixes), []))]
    return df.astype("float64")
