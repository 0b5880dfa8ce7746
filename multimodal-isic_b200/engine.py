"""Thin host wrapper over the C-ABI (``include/radb.h``): one ``Engine`` = one ``radb_handle``
on one device.  PyTorch is used for device/pinned buffers and streams only; every number is
computed by the sm_100a kernels in ``csrc/``.  No CPU fallback."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _abi


class RadbError(RuntimeError):
    pass


class Engine:
    def __init__(self, bin_width, label, angles, symmetrical_glcm=True, gldm_alpha=0.0, voxel_array_shift=0.0,
                 classes=_abi.CLASS_ORDER, max_ng=0, device=0, bin_count=0):
        if not torch.cuda.is_available():
            raise RadbError("radb: no CUDA device visible -- the radiomic engine has no CPU fallback")
        self.lib = _abi.load_library()
        self.device = int(device)
        self.label = int(label)
        self.n_angles = len(angles)
        self._settings = _abi.make_settings(bin_width, label, angles, symmetrical_glcm, gldm_alpha,
                                            voxel_array_shift, classes, max_ng, device, bin_count)
        h = ctypes.c_void_p()
        rc = self.lib.radb_create(ctypes.byref(self._settings), ctypes.byref(h))
        if rc != 0:
            raise RadbError("radb_create failed (%d): %s" % (rc, self.lib.radb_last_error().decode()))
        self._h = h
        self.F = self.lib.radb_feature_count(self._h)
        self.names = [self.lib.radb_feature_name(self._h, i).decode() for i in range(self.F)]
        self.max_ng = self.lib.radb_max_ng(self._h)
        self.has_packed = hasattr(self.lib, "radb_extract_packed")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.radb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(self.lib.radb_launch_count(self._h))

    def set_chunk(self, patches):
        """Patches per chunk of the device pipeline (0 = default); see radb_set_chunk."""
        rc = self.lib.radb_set_chunk(self._h, int(patches))
        if rc != 0:
            raise RadbError("radb_set_chunk failed (%d): %s" % (rc, self.lib.radb_last_error().decode()))

    def chunk_rows(self, B, H, W, dtype=torch.uint8):
        """Rows per chunk a dense call with ``B`` patches will use (radb_chunk_rows)."""
        n = int(self.lib.radb_chunk_rows(self._h, int(H), int(W), self.DTYPES[dtype], int(B)))
        if n < 0:
            raise RadbError("radb_chunk_rows failed (%d): %s" % (n, self.lib.radb_last_error().decode()))
        return n

    def set_chunk_events(self, events):
        """``events``: torch.cuda.Event list; event c is recorded when the rows of chunk c of the NEXT dense extraction
        call are final (radb_set_chunk_events).  The events must exist on the device already (record them once)."""
        import ctypes

        arr = (ctypes.c_void_p * len(events))(*[int(e.cuda_event) for e in events])
        rc = self.lib.radb_set_chunk_events(self._h, arr, len(events))
        if rc != 0:
            raise RadbError("radb_set_chunk_events failed (%d): %s" % (rc, self.lib.radb_last_error().decode()))

    def set_profiling(self, on):
        self.lib.radb_set_profiling(self._h, int(bool(on)))

    def kernel_ms(self):
        """Accumulated device milliseconds of the (build, angle, misc) kernels since profiling was switched on."""
        ms = (ctypes.c_double * 3)()
        rc = self.lib.radb_kernel_ms(self._h, ctypes.byref(ms))
        if rc != 0:
            raise RadbError("radb_kernel_ms failed (%d): %s" % (rc, self.lib.radb_last_error().decode()))
        return dict(build=ms[0], angle=ms[1], misc=ms[2])

    def smem_bytes(self, H, W, dtype=_abi.DTYPE_U8):
        return int(self.lib.radb_smem_bytes(self._h, H, W, dtype))

    DTYPES = {torch.uint8: _abi.DTYPE_U8, torch.uint16: _abi.DTYPE_U16, torch.float32: _abi.DTYPE_F32,
              torch.float64: _abi.DTYPE_F64}

    def _check(self, images, masks):
        if images.dtype not in self.DTYPES:
            raise NotImplementedError("pixel dtype %s is not implemented (uint8, uint16, float32, float64)" % images.dtype)
        if images.dtype != torch.uint8 and self._settings.max_ng <= 0 and self._settings.bin_count <= 0:
            raise RadbError("non-uint8 pixels need an engine created with max_ng (gray-level bound) or bin_count")
        if masks.dtype != torch.uint8:
            raise TypeError("masks must be uint8")
        if images.dim() != 3 or images.shape != masks.shape:
            raise ValueError("images and masks must both be [B, H, W]")
        if not (images.is_cuda and masks.is_cuda and images.device.index == self.device):
            raise ValueError("images / masks must live on cuda:%d" % self.device)
        if not (images.is_contiguous() and masks.is_contiguous()):
            raise ValueError("images / masks must be contiguous")

    def extract_device(self, images, masks, out=None, status=None, stream=None):
        """Device tensors in, device tensors out; asynchronous on ``stream`` (default: current)."""
        self._check(images, masks)
        B, H, W = images.shape
        dev = images.device
        if out is None:
            out = torch.empty((B, self.F), dtype=torch.float64, device=dev)
        if status is None:
            status = torch.empty((B,), dtype=torch.int32, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = self.lib.radb_extract(self._h, images.data_ptr(), self.DTYPES[images.dtype], masks.data_ptr(), B, H, W,
                                   H * W * images.element_size(), H * W, out.data_ptr(), status.data_ptr(),
                                   st.cuda_stream)
        if rc != 0:
            raise RadbError("radb_extract failed (%d): %s" % (rc, self.lib.radb_last_error().decode()))
        return out, status

    def extract_packed(self, images, mask_bits, out=None, status=None, stream=None, stride_b=0):
        """``images`` [B, H, W] device tensor, ``mask_bits`` device uint8 tensor holding the bit-packed masks
        (radb_extract_packed: bit i of patch b at byte ``b * stride_b + i // 8``, LSB first, set = ROI;
        ``stride_b`` defaults to ``packed_stride(H, W)``, the layout ``pack_masks_host`` writes)."""
        if images.dtype not in self.DTYPES or mask_bits.dtype != torch.uint8:
            raise TypeError("images must be uint8/uint16/float32/float64 and mask_bits uint8")
        if images.dim() != 3 or not (images.is_cuda and mask_bits.is_cuda and images.is_contiguous() and mask_bits.is_contiguous()):
            raise ValueError("images [B, H, W] and mask_bits must be contiguous CUDA tensors")
        B, H, W = images.shape
        stride = int(stride_b) if stride_b else self.packed_stride(H, W)
        if stride < (H * W + 7) // 8 or mask_bits.numel() < B * stride:
            raise ValueError("mask_bits holds fewer than B * stride bytes (stride >= ceil(H*W/8))")
        dev = images.device
        if out is None:
            out = torch.empty((B, self.F), dtype=torch.float64, device=dev)
        if status is None:
            status = torch.empty((B,), dtype=torch.int32, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = self.lib.radb_extract_packed(self._h, images.data_ptr(), self.DTYPES[images.dtype], mask_bits.data_ptr(), B, H, W,
                                          H * W * images.element_size(), stride, out.data_ptr(), status.data_ptr(),
                                          st.cuda_stream)
        if rc != 0:
            raise RadbError("radb_extract_packed failed (%d): %s" % (rc, self.lib.radb_last_error().decode()))
        return out, status

    def extract_ragged(self, img_pool, mask_pool, img_off, mask_off, hw, out=None, status=None, stream=None):
        """Variable-size batch (radb_extract_ragged): ``img_pool`` / ``mask_pool`` are 1-D device tensors
        holding the patches back to back, ``img_off`` / ``mask_off`` (int64, BYTES) and ``hw`` ([n, 2] int32)
        are host arrays.  Row i of the result belongs to patch i."""
        if not (img_pool.is_cuda and mask_pool.is_cuda and img_pool.is_contiguous() and mask_pool.is_contiguous()):
            raise ValueError("pools must be contiguous CUDA tensors")
        if img_pool.dtype not in self.DTYPES or mask_pool.dtype != torch.uint8:
            raise TypeError("img_pool must be uint8/uint16/float32/float64 and mask_pool uint8")
        import numpy as np

        img_off = np.ascontiguousarray(img_off, dtype=np.int64)
        mask_off = np.ascontiguousarray(mask_off, dtype=np.int64)
        hw = np.ascontiguousarray(hw, dtype=np.int32).reshape(-1, 2)
        n = len(hw)
        if len(img_off) != n or len(mask_off) != n:
            raise ValueError("img_off, mask_off and hw must have one entry per patch")
        es = img_pool.element_size()
        if n:
            px = hw[:, 0].astype(np.int64) * hw[:, 1]
            if (img_off < 0).any() or (mask_off < 0).any() or (img_off + px * es > img_pool.numel() * es).any() or \
                    (mask_off + px > mask_pool.numel()).any():
                raise ValueError("patch extends beyond its pool")
        dev = img_pool.device
        if out is None:
            out = torch.empty((n, self.F), dtype=torch.float64, device=dev)
        if status is None:
            status = torch.empty((n,), dtype=torch.int32, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = self.lib.radb_extract_ragged(self._h, img_pool.data_ptr(), self.DTYPES[img_pool.dtype], mask_pool.data_ptr(),
                                          n, img_off.ctypes.data, mask_off.ctypes.data, hw.ctypes.data, out.data_ptr(),
                                          status.data_ptr(), st.cuda_stream)
        if rc != 0:
            raise RadbError("radb_extract_ragged failed (%d): %s" % (rc, self.lib.radb_last_error().decode()))
        return out, status

    def extract_bgr(self, bgr, masks, stream=None, return_planes=False):
        """Decoded records in, four feature rows per record out (gray, R, G, B): ``bgr`` [n, H, W, 3]
        uint8 (cv2.imread layout), ``masks`` [n, H, W] uint8, both on the device.  The gray/R/G/B planes
        are produced by the front-end kernel; the four executes share the mask."""
        if bgr.dtype != torch.uint8 or masks.dtype != torch.uint8:
            raise TypeError("bgr and masks must be uint8")
        if bgr.dim() != 4 or bgr.shape[3] != 3 or tuple(bgr.shape[:3]) != tuple(masks.shape):
            raise ValueError("bgr must be [n, H, W, 3] and masks [n, H, W]")
        if not (bgr.is_cuda and masks.is_cuda and bgr.is_contiguous() and masks.is_contiguous()):
            raise ValueError("bgr / masks must be contiguous CUDA tensors")
        n, H, W, _ = bgr.shape
        dev = bgr.device
        planes = torch.empty((n, 4, H, W), dtype=torch.uint8, device=dev)
        out = torch.empty((n * 4, self.F), dtype=torch.float64, device=dev)
        status = torch.empty((n * 4,), dtype=torch.int32, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = self.lib.radb_extract_bgr(self._h, bgr.data_ptr(), masks.data_ptr(), n, H, W, planes.data_ptr(),
                                       out.data_ptr(), status.data_ptr(), st.cuda_stream)
        if rc != 0:
            raise RadbError("radb_extract_bgr failed (%d): %s" % (rc, self.lib.radb_last_error().decode()))
        if return_planes:
            return out, status, planes
        return out, status

    def pack_mask_host(self, masks, packed, threads):
        """Host half of the packed-mask transfer (radb_pack_mask_host): contiguous uint8 host tensor ``masks``
        -> 1 bit per pixel in the host tensor ``packed`` (numel >= ceil(masks.numel() / 8))."""
        if masks.is_cuda or packed.is_cuda or masks.dtype != torch.uint8 or packed.dtype != torch.uint8:
            raise TypeError("pack_mask_host takes uint8 host tensors")
        if not (masks.is_contiguous() and packed.is_contiguous()) or packed.numel() * 8 < masks.numel():
            raise ValueError("masks / packed must be contiguous and packed large enough")
        rc = self.lib.radb_pack_mask_host(masks.data_ptr(), masks.numel(), self.label, packed.data_ptr(), int(threads))
        if rc != 0:
            raise RadbError("radb_pack_mask_host failed (%d)" % rc)

    @staticmethod
    def packed_stride(H, W):
        """Bytes between the bit streams of consecutive patches: ceil(H*W/8) rounded up to 16 (TMA staging)."""
        return ((H * W + 7) // 8 + 15) // 16 * 16

    def pack_masks_host(self, masks, packed, threads):
        """Per-patch packing for ``extract_packed`` (radb_pack_masks_host): ``masks`` [n, H, W] uint8 host tensor ->
        ``packed`` host bytes, patch b at ``b * packed_stride(H, W)``."""
        if masks.is_cuda or packed.is_cuda or masks.dtype != torch.uint8 or packed.dtype != torch.uint8:
            raise TypeError("pack_masks_host takes uint8 host tensors")
        n, H, W = masks.shape
        stride = self.packed_stride(H, W)
        if not (masks.is_contiguous() and packed.is_contiguous()) or packed.numel() < n * stride:
            raise ValueError("masks / packed must be contiguous and packed large enough")
        rc = self.lib.radb_pack_masks_host(masks.data_ptr(), n, H * W, self.label, packed.data_ptr(), stride, int(threads))
        if rc != 0:
            raise RadbError("radb_pack_masks_host failed (%d)" % rc)

    def unpack_mask(self, packed, masks, stream=None):
        """Device half (radb_unpack_mask): ``packed`` device bytes -> ``masks`` uint8 device tensor (label where set)."""
        if not (packed.is_cuda and masks.is_cuda) or masks.dtype != torch.uint8 or not masks.is_contiguous():
            raise ValueError("unpack_mask takes contiguous CUDA tensors")
        st = stream if stream is not None else torch.cuda.current_stream(masks.device)
        rc = self.lib.radb_unpack_mask(self._h, packed.data_ptr(), masks.numel(), masks.data_ptr(), st.cuda_stream)
        if rc != 0:
            raise RadbError("radb_unpack_mask failed (%d): %s" % (rc, self.lib.radb_last_error().decode()))

    def resize_mask(self, masks, size, stream=None):
        """``cv2.resize(mask, (W, H), interpolation=cv2.INTER_NEAREST)`` (RadiomicExtractor.py:34-35) on the device, bit-exact
        with cv2 4.x: ``masks`` uint8 CUDA [n, h, w] -> uint8 CUDA [n, H, W] for ``size = (H, W)``."""
        if masks.dtype != torch.uint8 or not masks.is_cuda or not masks.is_contiguous() or masks.dim() != 3:
            raise ValueError("resize_mask takes a contiguous uint8 CUDA tensor [n, h, w]")
        n, sh, sw = masks.shape
        H, W = int(size[0]), int(size[1])
        out = torch.empty((n, H, W), dtype=torch.uint8, device=masks.device)
        st = stream if stream is not None else torch.cuda.current_stream(masks.device)
        rc = self.lib.radb_resize_mask(self._h, masks.data_ptr(), n, sh, sw, out.data_ptr(), H, W, st.cuda_stream)
        if rc != 0:
            raise RadbError("radb_resize_mask failed (%d): %s" % (rc, self.lib.radb_last_error().decode()))
        return out

    def derive_image(self, images, type_code, stream=None):
        """Point-wise derived image type (1 Square, 2 SquareRoot, 3 Logarithm, 4 Exponential) of uint8
        device images [B, H, W] -> float64 [B, H, W] (pyradiomics imageoperations.get*Image)."""
        if images.dtype != torch.uint8 or not images.is_cuda or not images.is_contiguous():
            raise ValueError("derive_image takes contiguous uint8 CUDA images")
        B, H, W = images.shape
        out = torch.empty((B, H, W), dtype=torch.float64, device=images.device)
        mx = torch.empty((B,), dtype=torch.int32, device=images.device)
        st = stream if stream is not None else torch.cuda.current_stream(images.device)
        rc = self.lib.radb_derive_image(self._h, images.data_ptr(), B, H * W, int(type_code), out.data_ptr(),
                                        mx.data_ptr(), st.cuda_stream)
        if rc != 0:
            raise RadbError("radb_derive_image failed (%d): %s" % (rc, self.lib.radb_last_error().decode()))
        return out

    def filter_image(self, images, kind, sigma=0.0, x_only=False, stream=None):
        """Filtered image types (radb_filter_image) of uint8 device images [B, H, W]:
        ``kind`` "Gradient" -> float32 [B, H, W]; "LoG" (``sigma``) -> float32 [B, H, W]; "Wavelet" -> float64
        [B, K, H, W] with K = 2 (wavelet-H, wavelet-L; ``x_only``: pyradiomics' force2D axis removal) or 4
        (wavelet-LH, -HL, -HH, -LL)."""
        if images.dtype != torch.uint8 or not images.is_cuda or not images.is_contiguous() or images.dim() != 3:
            raise ValueError("filter_image takes contiguous uint8 CUDA images [B, H, W]")
        B, H, W = images.shape
        dev = images.device
        code = {"Gradient": 5, "LoG": 6, "Wavelet": 7}[kind]
        scratch = None
        if kind == "Gradient":
            out = torch.empty((B, H, W), dtype=torch.float32, device=dev)
        elif kind == "LoG":
            if H < 4 or W < 4:
                raise ValueError("LoG needs at least 4 pixels per axis (ITK recursive Gaussian)")
            out = torch.empty((B, H, W), dtype=torch.float32, device=dev)
            scratch = torch.empty((B * H * W * 12 + 16,), dtype=torch.uint8, device=dev)
        else:
            out = torch.empty((B, 2 if x_only else 4, H, W), dtype=torch.float64, device=dev)
            if not x_only:
                scratch = torch.empty((B * 2 * (H + 1) * (W + 1),), dtype=torch.float64, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = self.lib.radb_filter_image(self._h, images.data_ptr(), B, H, W, code, float(sigma), 1 if x_only else 0,
                                        out.data_ptr(), scratch.data_ptr() if scratch is not None else None, st.cuda_stream)
        if rc != 0:
            raise RadbError("radb_filter_image failed (%d): %s" % (rc, self.lib.radb_last_error().decode()))
        if scratch is not None:
            scratch.record_stream(st)
        return out

    def debug_matrices(self, images, masks):
        """Features plus the integer matrices (numpy, trimmed to shapes the oracle uses)."""
        self._check(images, masks)
        B, H, W = images.shape
        dev = images.device
        ng, na, nr = self.max_ng, self.n_angles, max(H, W)
        z = lambda shape, dt=torch.int32: torch.zeros(shape, dtype=dt, device=dev)
        out = z((B, self.F), torch.float64)
        status = z((B,))
        bufs = dict(levels=z((B, H, W)), glcm=z((B, na, ng, ng)), glrlm=z((B, na, ng, nr)),
                    glszm=z((B, ng, H * W)), gldm=z((B, ng, 2 * na + 1)), ngtdm_n=z((B, ng)),
                    ngtdm_s=z((B, ng), torch.float64), ng=z((B,)))
        st = torch.cuda.current_stream(dev)
        rc = self.lib.radb_debug_matrices(self._h, images.data_ptr(), self.DTYPES[images.dtype], masks.data_ptr(), B, H, W,
                                          H * W * images.element_size(), H * W, out.data_ptr(), status.data_ptr(),
                                          *[bufs[k].data_ptr() for k in ("levels", "glcm", "glrlm", "glszm", "gldm",
                                                                          "ngtdm_n", "ngtdm_s", "ng")],
                                          st.cuda_stream)
        if rc != 0:
            raise RadbError("radb_debug_matrices failed (%d): %s" % (rc, self.lib.radb_last_error().decode()))
        torch.cuda.synchronize(dev)
        res = {k: v.cpu().numpy() for k, v in bufs.items()}
        res["features"] = out.cpu().numpy()
        res["status"] = status.cpu().numpy()
        return res


def pack_ragged(images, masks, align=16):
    """Host-side packing of variable-size (image [H, W], mask [H, W]) pairs into two flat pools with
    ``align``-byte aligned patch starts (keeps the TMA staging path).  Returns NumPy
    ``(img_pool, mask_pool, img_off, mask_off, hw)`` ready for ``Engine.extract_ragged``."""
    import numpy as np

    if len(images) != len(masks):
        raise ValueError("one mask per image")
    n = len(images)
    if n == 0:
        return np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros((0, 2), np.int32)
    dt = np.asarray(images[0]).dtype
    es = dt.itemsize
    hw = np.zeros((n, 2), np.int32)
    img_off, mask_off = np.zeros(n, np.int64), np.zeros(n, np.int64)
    io = mo = 0
    for i, (im, mk) in enumerate(zip(images, masks)):
        im, mk = np.asarray(im), np.asarray(mk)
        if im.ndim != 2 or im.shape != mk.shape or im.dtype != dt:
            raise ValueError("every image must be 2-D, of one dtype, with a mask of its own shape")
        hw[i] = im.shape
        img_off[i], mask_off[i] = io, mo
        io += (im.size * es + align - 1) // align * align
        mo += (im.size + align - 1) // align * align
    img_pool = np.zeros(io // es, dt)
    mask_pool = np.zeros(mo, np.uint8)
    for i, (im, mk) in enumerate(zip(images, masks)):
        k = hw[i, 0] * hw[i, 1]
        img_pool[img_off[i] // es: img_off[i] // es + k] = np.asarray(im).reshape(-1)
        mask_pool[mask_off[i]: mask_off[i] + k] = np.asarray(mk, dtype=np.uint8).reshape(-1)
    return img_pool, mask_pool, img_off, mask_off, hw


class HostPipeline:
    """Host-buffer entry point: streams host (NumPy) patches through the engine in chunks with
    double-buffered pinned staging so H2D copies, kernels and D2H copies overlap.

    The path is bound by the host-to-device link (pixels + masks, ~50 GB/s), and of a mask only
    ``mask == label`` matters, so with ``pack_masks`` the masks cross the link at 1 bit per pixel: packed on
    the host by ``pack_threads`` persistent threads (AVX2, memory bound: ~100 GB/s on 16 cores) while the device works on
    the previous chunks (a helper thread runs one chunk ahead of the enqueue loop, ``slots`` chunks are in
    flight) and read packed by the kernels (``radb_extract_packed``): 15.9 -> 9.7-10.1 ms per 100 k 64x64 patches
    (the raw link needs 8.5 ms for the same bytes; caller-packed masks: 9.4 ms).  ``pack_masks=None`` (default)
    enables it when this process has at least 3 host cores to itself (``os.cpu_count() // LOCAL_WORLD_SIZE``), and
    ``adaptive=None`` makes it per-chunk adaptive below 6 cores per rank (eight ranks on a 32-core host are bound by
    the host's memory system, not the link: see the policy note in ``__init__``)."""

    def __init__(self, engine, chunk=4096, pack_masks=None, pack_threads=None, slots=6, slot_bytes=128 << 20, adaptive=None,
                 ramp=None):
        import os

        self.engine = engine
        self.chunk = int(chunk)
        # graded chunk sizes (see chunk_schedule): small chunks first and last, so the pipeline's fill (pack + H2D of the
        # first chunk before any kernel runs) and drain (kernels + D2H of the last chunk after the link went idle) shrink.
        # Measured (scripts/e2e_ramp_ab.py, profiles/r2_e2e_chunking.jsonl): worth ~3 % with 8192-patch chunks, nothing
        # with the 4096-patch chunks that are the default now (uniformly smaller chunks shorten both ends as well) -> off
        # True / False, or "4,2/2,4,8": the divisors of ``chunk`` at the head / tail (tuning knob, env RADB_E2E_RAMP)
        env = os.environ.get("RADB_E2E_RAMP", "0")
        self.ramp = ramp if ramp is not None else (env if "/" in env else env == "1")
        self.slot_bytes = int(slot_bytes)  # pinned / device bytes of one image buffer of one slot, at most
        self.slots = max(2, int(slots))  # chunks in flight: the host packs / enqueues ahead of the device
        self._bufs = None
        self._key = None
        self._pool = None
        self._n = 0  # patches per slot of the current buffers
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        cores = max(1, (os.cpu_count() or 1) // ranks)
        # Packing threads: all of a rank's cores up to 4, then 3/4 of them, 12 at most (measured on 16 cores, one rank:
        # 6 / 8 / 12 / 16 threads -> 10.3 / 10.5 / 9.95 / 10.0 ms per step)
        self.pack_threads = int(pack_threads) if pack_threads else (cores if cores <= 4 else max(4, min(cores * 3 // 4, 12)))
        # Default policy (measured with the persistent packing threads, scripts/e2e_modes.py ->
        # profiles/r2_e2e_modes_{2,8}gpu_pool.jsonl).  Packing trades host memory traffic (the packer reads the 4 KB the
        # DMA engine would have read) for link bytes, so it pays while the LINK is the bound: one rank on 16 cores 6.3 ->
        # 10.3 M patches/s, two ranks on 24 cores 12.5 -> 17.1 M (8 threads each; break-even at 3 threads per rank).  Eight
        # ranks on 32 cores are bound by the host's memory system either way (19.9 M raw, 19.7 M packed): there the
        # adaptive mode -- pack a chunk only while the host keeps ahead of the link -- is best (20.5 M), and it is the
        # default whenever a rank has fewer than 6 cores to itself.
        self.pack_masks = bool(pack_masks) if pack_masks is not None else cores >= 3
        self.h2d_bytes = 0  # bytes copied host -> device by the last run()
        self.timeline = None  # set to [] to record per-chunk CUDA-event marks of the next run() (scripts/e2e_timeline.py)
        # pack a chunk's masks only while the host keeps ahead of the link (see run())
        self.adaptive = bool(adaptive) if adaptive is not None else (ranks > 1 and cores < 6)
        self._host_ahead = True
        self.packed_chunks = self.total_chunks = 0

    def slot_patches(self, B, H, W, itemsize):
        """Patches per slot: the chunk size, capped by the batch and by ``slot_bytes`` per image buffer, so that a
        ten-image call at the reference's 600x450 size pins ~10 images, not 8192 of them."""
        n = min(self.chunk, max(1, int(B)), max(1, self.slot_bytes // (H * W * itemsize)))
        return max(1, n)

    @staticmethod
    def chunk_schedule(B, chunk, ramp=True):
        """Chunk sizes of one run() over B patches, in order; they sum to B and none exceeds ``chunk``.  With ``ramp``
        and enough patches the run starts with chunk/4 and chunk/2 and ends with chunk/2, chunk/4, chunk/8 (the
        remainder sorted in among them): the link is the bottleneck of the host-to-host path, and what it cannot hide
        is the time before the first kernel starts and after the last copy ends -- both proportional to the size of
        the chunk at that end."""
        B, chunk = int(B), max(1, int(chunk))
        if B <= 0:
            return []
        head = [chunk // 4, chunk // 2]
        tail = [chunk // 2, chunk // 4, chunk // 8]
        if isinstance(ramp, str):  # "h1,h2/t1,t2,t3": divisors of chunk
            hs, ts = ramp.split("/")
            head = [max(1, chunk // int(d)) for d in hs.split(",") if d]
            tail = [max(1, chunk // int(d)) for d in ts.split(",") if d]
        if not ramp or chunk < 64 or B < sum(head) + sum(tail) + 2 * chunk:
            return [min(chunk, B - s) for s in range(0, B, chunk)]
        rest = B - sum(head) - sum(tail)
        full, rem = divmod(rest, chunk)
        last = sorted(tail + ([rem] if rem else []), reverse=True)
        return head + [chunk] * full + last

    def _ensure(self, n, H, W, dtype=torch.uint8):
        key = (H, W, dtype)
        if self._key == key and self._n >= n:
            return
        dev = torch.device("cuda", self.engine.device)
        F = self.engine.F
        self._bufs = None  # release the previous buffers first
        bufs = []
        for _ in range(self.slots):
            bufs.append(dict(
                h_img=torch.empty((n, H, W), dtype=dtype).pin_memory(),
                h_msk=None,
                d_img=torch.empty((n, H, W), dtype=dtype, device=dev),
                h_pk=torch.empty((n * Engine.packed_stride(H, W),), dtype=torch.uint8).pin_memory(),
                d_pk=torch.empty((n * Engine.packed_stride(H, W),), dtype=torch.uint8, device=dev),
                d_msk=None,
                d_out=torch.empty((n, F), dtype=torch.float64, device=dev),
                d_st=torch.empty((n,), dtype=torch.int32, device=dev),
                stream=torch.cuda.Stream(dev), done=torch.cuda.Event()))
        self._bufs = bufs
        self._key = key
        self._n = n

    def run(self, images, masks, out=None, status=None, device_out=None, masks_packed=False):
        """``images``/``masks``: host arrays [B, H, W] (images uint8/uint16/float32/float64, masks uint8; NumPy or CPU tensors; pinned tensors
        skip the staging copy).  Returns host ``(features [B, F] float64, status [B] int32)``.  ``device_out``
        (optional ``[B, F]`` float64 CUDA tensor) also keeps the rows on the device -- the multi-GPU driver
        all-gathers them from there.  ``masks_packed``: ``masks`` already holds the bit-packed streams
        (``[B, packed_stride(H, W)]`` uint8, the layout of ``pack_masks_host``): no host packing at all."""
        images = torch.as_tensor(images)
        masks = torch.as_tensor(masks)
        if images.is_cuda:
            raise ValueError("HostPipeline takes host buffers; use Engine.extract_device for device tensors")
        B, H, W = images.shape
        F = self.engine.F
        chunk = self.slot_patches(B, H, W, images.element_size())
        self._ensure(chunk, H, W, images.dtype)
        if out is None:
            out = torch.empty((B, F), dtype=torch.float64).pin_memory()
        if status is None:
            status = torch.empty((B,), dtype=torch.int32).pin_memory()
        pinned_in = images.is_pinned() and masks.is_pinned()
        # masks cross the link at 1 bit per pixel and are consumed packed by the kernels (radb_extract_packed);
        # patches whose pixel count is not a multiple of 128 bits keep the byte masks (16-byte aligned bit rows)
        pstride = self.engine.packed_stride(H, W)
        if masks_packed:
            if masks.dtype != torch.uint8 or masks.dim() != 2 or masks.shape[0] != B or masks.shape[1] != pstride or not masks.is_contiguous():
                raise ValueError("masks_packed: masks must be a contiguous uint8 [B, %d] tensor" % pstride)
        pack = (not masks_packed) and self.pack_masks and masks.dtype == torch.uint8 and masks.is_contiguous() and self.engine.has_packed
        self.h2d_bytes = 0
        self.packed_chunks = 0
        sizes = self.chunk_schedule(B, chunk, self.ramp)
        starts = [0] * len(sizes)
        for k in range(1, len(sizes)):
            starts[k] = starts[k - 1] + sizes[k - 1]
        self.total_chunks = len(starts)
        dev = torch.device("cuda", self.engine.device)

        import time

        def prepare(k):
            """Host work of chunk k, one chunk ahead of the enqueue loop on a helper thread: wait until the
            slot has drained, then pack the masks (the C call releases the GIL and fans out to pack_threads).
            Adaptive: packing trades host work for link bytes, so it only pays while the host is AHEAD of the link.
            If the slot had to be waited for, the link / device is the bottleneck -> pack; if it was already free the
            host is the bottleneck (few cores per rank) -> hand this chunk's masks over as bytes.  Returns whether
            the chunk was packed."""
            s0, n0 = starts[k], sizes[k]
            bk = self._bufs[k % self.slots]
            t0 = time.perf_counter()
            bk["done"].synchronize()
            t1 = time.perf_counter()
            waited = t1 - t0 > 30e-6
            self.stats["slot_wait_s"] += t1 - t0
            if not pack:
                return False
            do = (not self.adaptive) or waited or (k < self.slots and self._host_ahead)
            if k >= self.slots:
                self._host_ahead = waited
            if do:
                self.engine.pack_masks_host(masks[s0:s0 + n0], bk["h_pk"], self.pack_threads)
                self.packed_chunks += 1
                self.stats["pack_s"] += time.perf_counter() - t1
            return do

        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor

            self._pool = ThreadPoolExecutor(1)
        # where the host side of the last run() spent its time: helper thread waiting for a slot to drain (the link /
        # device is behind) and packing; enqueue loop waiting for the helper (the host is behind)
        self.stats = {"slot_wait_s": 0.0, "pack_s": 0.0, "result_wait_s": 0.0}
        fut = self._pool.submit(prepare, 0) if starts else None
        for k, s in enumerate(starts):
            n = sizes[k]
            b = self._bufs[k % self.slots]
            t_w = time.perf_counter()
            packed = fut.result()
            self.stats["result_wait_s"] += time.perf_counter() - t_w
            if k + 1 < len(starts):
                fut = self._pool.submit(prepare, k + 1)
            nb = n * pstride
            d_out = device_out[s:s + n] if device_out is not None else b["d_out"][:n]
            tl = None
            if self.timeline is not None:
                tl = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                if not self.timeline:
                    self._tl0 = torch.cuda.Event(enable_timing=True)
                    self._tl0.record(b["stream"])
                self.timeline.append((k, n, tl))
                tl[0].record(b["stream"])
            with torch.cuda.stream(b["stream"]):
                if pinned_in:
                    b["d_img"][:n].copy_(images[s:s + n], non_blocking=True)
                else:
                    b["h_img"][:n].copy_(images[s:s + n])
                    b["d_img"][:n].copy_(b["h_img"][:n], non_blocking=True)
                if masks_packed:
                    src = masks[s:s + n].reshape(-1)
                    if not pinned_in:
                        b["h_pk"][:nb].copy_(src)
                        src = b["h_pk"][:nb]
                    b["d_pk"][:nb].copy_(src, non_blocking=True)
                    self.h2d_bytes += n * H * W * images.element_size() + nb
                    if tl is not None:
                        tl[1].record(b["stream"])
                    self.engine.extract_packed(b["d_img"][:n], b["d_pk"], d_out, b["d_st"][:n], stream=b["stream"])
                elif packed:
                    b["d_pk"][:nb].copy_(b["h_pk"][:nb], non_blocking=True)
                    self.h2d_bytes += n * H * W * images.element_size() + nb
                    if tl is not None:
                        tl[1].record(b["stream"])
                    self.engine.extract_packed(b["d_img"][:n], b["d_pk"], d_out, b["d_st"][:n], stream=b["stream"])
                else:
                    if b["d_msk"] is None or b["d_msk"].shape[0] < n:
                        b["d_msk"] = torch.empty((self._n, H, W), dtype=torch.uint8, device=dev)
                    if pinned_in:
                        b["d_msk"][:n].copy_(masks[s:s + n], non_blocking=True)
                    else:
                        if b["h_msk"] is None or b["h_msk"].shape[0] < n:
                            b["h_msk"] = torch.empty((self._n, H, W), dtype=torch.uint8).pin_memory()
                        b["h_msk"][:n].copy_(masks[s:s + n])
                        b["d_msk"][:n].copy_(b["h_msk"][:n], non_blocking=True)
                    self.h2d_bytes += n * H * W * images.element_size() + n * H * W
                    if tl is not None:
                        tl[1].record(b["stream"])
                    self.engine.extract_device(b["d_img"][:n], b["d_msk"][:n], d_out, b["d_st"][:n], stream=b["stream"])
                if tl is not None:
                    tl[2].record(b["stream"])
                out[s:s + n].copy_(d_out, non_blocking=True)
                status[s:s + n].copy_(b["d_st"][:n], non_blocking=True)
                if tl is not None:
                    tl[3].record(b["stream"])
                b["done"].record(b["stream"])
        for b in self._bufs:
            b["done"].synchronize()
        return out, status

    def timeline_ms(self):
        """Per-chunk marks of the run recorded with ``timeline = []``: (chunk, patches, H2D start, H2D end = kernels start,
        kernels end, D2H end), milliseconds since the first chunk's H2D start."""
        torch.cuda.synchronize()
        rows = [(k, n) + tuple(self._tl0.elapsed_time(e) for e in ev) for k, n, ev in self.timeline]
        self.timeline = None
        return rows
