"""ctypes mirror of ``include/radb.h`` (the C-ABI drop-in boundary) and the loader of
``libradb_b200.so``.  There is no CPU fallback: if the CUDA library is missing or no GPU
is visible, construction fails loudly."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libradb_b200.so")

CLASS_BITS = {"firstorder": 1, "glcm": 2, "gldm": 4, "glrlm": 8, "glszm": 16, "ngtdm": 32, "shape2D": 64}
CLASS_ORDER = ("firstorder", "glcm", "gldm", "glrlm", "glszm", "ngtdm")
DTYPE_U8, DTYPE_U16, DTYPE_F32, DTYPE_F64 = 0, 1, 2, 3

STATUS_MESSAGES = {
    1: "Label (%s) not present in mask",
    2: "mask only contains 1 segmented voxel! Cannot extract features for a single voxel.",
    3: "mask has too few dimensions (number of dimensions %d, minimum required %d)",
    4: "ROI has more gray levels than the extractor was sized for (raise max_ng)",
}


class RadbSettings(ctypes.Structure):
    _fields_ = [
        ("bin_width", ctypes.c_double),
        ("bin_count", ctypes.c_int32),
        ("label", ctypes.c_int32),
        ("n_angles", ctypes.c_int32),
        ("angles", (ctypes.c_int8 * 2) * 8),
        ("symmetrical_glcm", ctypes.c_int32),
        ("gldm_alpha", ctypes.c_double),
        ("voxel_array_shift", ctypes.c_double),
        ("class_mask", ctypes.c_uint32),
        ("max_ng", ctypes.c_int32),
        ("device", ctypes.c_int32),
    ]


def make_settings(bin_width, label, angles, symmetrical_glcm=True, gldm_alpha=0.0, voxel_array_shift=0.0,
                  classes=CLASS_ORDER, max_ng=0, device=0, bin_count=0):
    s = RadbSettings()
    s.bin_width = float(bin_width)
    s.bin_count = int(bin_count or 0)
    s.label = int(label)
    s.n_angles = len(angles)
    for a, (dy, dx) in enumerate(angles):
        s.angles[a][0] = int(dy)
        s.angles[a][1] = int(dx)
    s.symmetrical_glcm = int(bool(symmetrical_glcm))
    s.gldm_alpha = float(gldm_alpha)
    s.voxel_array_shift = float(voxel_array_shift)
    mask = 0
    for c in classes:
        mask |= CLASS_BITS[c]
    s.class_mask = mask
    s.max_ng = int(max_ng)
    s.device = int(device)
    return s


_lib = None


def load_library(path=None):
    """Load ``libradb_b200.so`` and declare every prototype of ``include/radb.h``."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("RADB_LIB") or LIB_PATH  # RADB_LIB: A/B builds of the same CUDA library
    if not os.path.exists(path):
        raise RuntimeError(
            "radb: CUDA library %s is missing -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.radb_create.argtypes = [ctypes.POINTER(RadbSettings), ctypes.POINTER(vp)]
    lib.radb_create.restype = i32
    lib.radb_destroy.argtypes = [vp]
    lib.radb_destroy.restype = None
    lib.radb_feature_count.argtypes = [vp]
    lib.radb_feature_count.restype = i32
    lib.radb_feature_name.argtypes = [vp, i32]
    lib.radb_feature_name.restype = ctypes.c_char_p
    lib.radb_reserve.argtypes = [vp, i32, i32, i32, i64, vp]
    lib.radb_reserve.restype = i32
    lib.radb_smem_bytes.argtypes = [vp, i32, i32, i32]
    lib.radb_smem_bytes.restype = i32
    lib.radb_extract.argtypes = [vp, vp, i32, vp, i64, i32, i32, i64, i64, vp, vp, vp]
    lib.radb_extract.restype = i32
    lib.radb_extract_packed.argtypes = [vp, vp, i32, vp, i64, i32, i32, i64, i64, vp, vp, vp]
    lib.radb_extract_packed.restype = i32
    lib.radb_extract_ragged.argtypes = [vp, vp, i32, vp, i64, vp, vp, vp, vp, vp, vp]
    lib.radb_extract_ragged.restype = i32
    lib.radb_extract_bgr.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp, vp, vp]
    lib.radb_extract_bgr.restype = i32
    lib.radb_pack_mask_host.argtypes = [vp, i64, i32, vp, i32]
    lib.radb_pack_mask_host.restype = i32
    lib.radb_pack_masks_host.argtypes = [vp, i64, i64, i32, vp, i64, i32]
    lib.radb_pack_masks_host.restype = i32
    lib.radb_resize_mask.argtypes = [vp, vp, i64, i32, i32, vp, i32, i32, vp]
    lib.radb_resize_mask.restype = i32
    lib.radb_unpack_mask.argtypes = [vp, vp, i64, vp, vp]
    lib.radb_unpack_mask.restype = i32
    lib.radb_derive_image.argtypes = [vp, vp, i64, i64, i32, vp, vp, vp]
    lib.radb_derive_image.restype = i32
    lib.radb_filter_image.argtypes = [vp, vp, i64, i32, i32, i32, ctypes.c_double, i32, vp, vp, vp]
    lib.radb_filter_image.restype = i32
    lib.radb_debug_matrices.argtypes = [vp, vp, i32, vp, i64, i32, i32, i64, i64, vp, vp] + [vp] * 8 + [vp]
    lib.radb_debug_matrices.restype = i32
    lib.radb_max_ng.argtypes = [vp]
    lib.radb_max_ng.restype = i32
    lib.radb_launch_count.argtypes = [vp]
    lib.radb_launch_count.restype = i64
    lib.radb_set_chunk.argtypes = [vp, i64]
    lib.radb_set_chunk.restype = i32
    lib.radb_chunk_rows.argtypes = [vp, i32, i32, i32, i64]
    lib.radb_chunk_rows.restype = i64
    lib.radb_set_chunk_events.argtypes = [vp, vp, i32]
    lib.radb_set_chunk_events.restype = i32
    lib.radb_set_profiling.argtypes = [vp, i32]
    lib.radb_set_profiling.restype = i32
    lib.radb_kernel_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_double * 3)]
    lib.radb_kernel_ms.restype = i32
    lib.radb_last_error.argtypes = []
    lib.radb_last_error.restype = ctypes.c_char_p
    lib.radb_version.argtypes = []
    lib.radb_version.restype = ctypes.c_char_p
    if path == LIB_PATH:
        _lib = lib
    return lib


EXPORTED_SYMBOLS = (
    "radb_create", "radb_destroy", "radb_feature_count", "radb_feature_name", "radb_reserve", "radb_smem_bytes",
    "radb_extract", "radb_extract_packed", "radb_extract_ragged", "radb_extract_bgr", "radb_pack_mask_host", "radb_pack_masks_host", "radb_unpack_mask", "radb_resize_mask", "radb_derive_image", "radb_filter_image", "radb_debug_matrices", "radb_max_ng", "radb_launch_count", "radb_set_chunk", "radb_chunk_rows", "radb_set_chunk_events", "radb_set_profiling",
    "radb_kernel_ms", "radb_last_error",
    "radb_version",
)
