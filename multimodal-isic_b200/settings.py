"""Extraction settings: parse the pyradiomics parameter file the reference passes to
``RadiomicsFeatureExtractor(param_file)`` (``/root/reference/RadiomicExtractor.py:15``,
``/root/reference/params.yml``) or an equivalent dict, and resolve them into what the
CUDA engine needs."""
from __future__ import annotations

import logging
import warnings
from collections import OrderedDict

import yaml

from ._abi import CLASS_ORDER

logger = logging.getLogger(__name__)

# pyradiomics defaults (featureextractor._getDefaultSettings / per-class kwargs.get defaults)
DEFAULTS = OrderedDict(
    minimumROIDimensions=2, minimumROISize=None, normalize=False, normalizeScale=1, removeOutliers=None,
    resampledPixelSpacing=None, interpolator="sitkBSpline", preCrop=False, padDistance=5, distances=[1],
    force2D=False, force2Ddimension=0, resegmentRange=None, label=1, additionalInfo=True,
    binWidth=25, binCount=None, symmetricalGLCM=True, weightingNorm=None, gldm_a=0, voxelArrayShift=0,
)

FEATURE_NAMES = {
    "firstorder": ["10Percentile", "90Percentile", "Energy", "Entropy", "InterquartileRange", "Kurtosis",
                   "Maximum", "MeanAbsoluteDeviation", "Mean", "Median", "Minimum", "Range",
                   "RobustMeanAbsoluteDeviation", "RootMeanSquared", "Skewness", "TotalEnergy", "Uniformity",
                   "Variance"],
    "glcm": ["Autocorrelation", "ClusterProminence", "ClusterShade", "ClusterTendency", "Contrast",
             "Correlation", "DifferenceAverage", "DifferenceEntropy", "DifferenceVariance", "Id", "Idm", "Idmn",
             "Idn", "Imc1", "Imc2", "InverseVariance", "JointAverage", "JointEnergy", "JointEntropy", "MCC",
             "MaximumProbability", "SumAverage", "SumEntropy", "SumSquares"],
    "gldm": ["DependenceEntropy", "DependenceNonUniformity", "DependenceNonUniformityNormalized",
             "DependenceVariance", "GrayLevelNonUniformity", "GrayLevelVariance", "HighGrayLevelEmphasis",
             "LargeDependenceEmphasis", "LargeDependenceHighGrayLevelEmphasis",
             "LargeDependenceLowGrayLevelEmphasis", "LowGrayLevelEmphasis", "SmallDependenceEmphasis",
             "SmallDependenceHighGrayLevelEmphasis", "SmallDependenceLowGrayLevelEmphasis"],
    "glrlm": ["GrayLevelNonUniformity", "GrayLevelNonUniformityNormalized", "GrayLevelVariance",
              "HighGrayLevelRunEmphasis", "LongRunEmphasis", "LongRunHighGrayLevelEmphasis",
              "LongRunLowGrayLevelEmphasis", "LowGrayLevelRunEmphasis", "RunEntropy", "RunLengthNonUniformity",
              "RunLengthNonUniformityNormalized", "RunPercentage", "RunVariance", "ShortRunEmphasis",
              "ShortRunHighGrayLevelEmphasis", "ShortRunLowGrayLevelEmphasis"],
    "glszm": ["GrayLevelNonUniformity", "GrayLevelNonUniformityNormalized", "GrayLevelVariance",
              "HighGrayLevelZoneEmphasis", "LargeAreaEmphasis", "LargeAreaHighGrayLevelEmphasis",
              "LargeAreaLowGrayLevelEmphasis", "LowGrayLevelZoneEmphasis", "SizeZoneNonUniformity",
              "SizeZoneNonUniformityNormalized", "SmallAreaEmphasis", "SmallAreaHighGrayLevelEmphasis",
              "SmallAreaLowGrayLevelEmphasis", "ZoneEntropy", "ZonePercentage", "ZoneVariance"],
    "ngtdm": ["Busyness", "Coarseness", "Complexity", "Contrast", "Strength"],
    "shape2D": ["Elongation", "MajorAxisLength", "MaximumDiameter", "MeshSurface", "MinorAxisLength", "Perimeter",
                "PerimeterSurfaceRatio", "PixelSurface", "Sphericity"],
}
SUPPORTED_CLASSES = set(CLASS_ORDER) | {"shape2D"}
SUPPORTED_IMAGE_TYPES = {"Original", "Square", "SquareRoot", "Logarithm", "Exponential", "Gradient", "LoG", "Wavelet"}
IMAGE_TYPE_CODES = {"Square": 1, "SquareRoot": 2, "Logarithm": 3, "Exponential": 4}  # radb_derive_image
# settings whose non-default value would change results and that the engine does not implement
_UNSUPPORTED_IF_SET = ("normalize", "removeOutliers", "resampledPixelSpacing", "resegmentRange", "weightingNorm",
                       "minimumROISize", "preCrop")


def in_plane_angles(ndim=2, distances=(1,), force2D=False, force2Ddimension=0):
    """pyradiomics ``cmatrices.c:build_angles`` for an ``ndim``-D array: the
    unidirectional offsets, in the C generator's order.  For the reference's literal call --
    a 2-D SimpleITK image (``RadiomicExtractor.py:31``) with ``force2D: True``
    (``params.yml:100``, ``force2Ddimension`` 0) -- axis 0 is removed and a single along-row
    offset remains; ``force2D: False`` on a 2-D image gives the 4 in-plane offsets."""
    if list(distances) != [1]:
        raise NotImplementedError("distances other than [1] are not implemented")
    if ndim != 2:
        raise NotImplementedError("only 2-D images (the reference's input) are implemented")
    fd = force2Ddimension if force2D else -1
    D, stride = 1, 3
    n_all = 1
    for d in range(ndim):
        if d != fd:
            n_all *= stride
    n_all -= 1
    out = []
    for a_idx in range(n_all // 2):
        a_off, ang = 1, []
        for d in range(ndim):
            if d == fd:
                ang.append(0)
            else:
                ang.append(D - (a_idx // a_off) % stride)
                a_off *= stride
        out.append(tuple(ang))
    return out


class Settings:
    """Resolved settings + enabled image types / feature classes (file order preserved)."""

    def __init__(self, params=None, strict=True, **overrides):
        """``strict`` (default): an enabled image type or feature class this engine does not implement raises
        ``NotImplementedError`` naming it -- a drop-in must not silently emit a narrower column set (the
        reference's driver silences warnings, extract_radiomics.py:13-19).  ``strict=False`` logs at ERROR level,
        warns, skips them and lists them in ``skipped_image_types`` / ``skipped_classes``."""
        if params is None:
            params = {}
        if isinstance(params, (str, bytes)) or hasattr(params, "__fspath__"):
            with open(params) as fh:
                params = yaml.safe_load(fh) or {}
        if not isinstance(params, dict):
            raise TypeError("param_file must be a path or a dict")
        if not any(k in params for k in ("setting", "imageType", "featureClass")):
            params = {"setting": dict(params)}  # a bare settings dict
        given = dict(params.get("setting") or {})
        given.update(overrides)
        self.settings = OrderedDict(DEFAULTS)
        self.settings.update(given)
        self._given = given
        # pyradiomics: no imageType section -> Original only; no featureClass section -> all classes
        image_types = params.get("imageType")
        self.enabledImagetypes = OrderedDict((k, v or {}) for k, v in (image_types or {"Original": {}}).items())
        feature_class = params.get("featureClass")
        if feature_class is None:
            feature_class = OrderedDict((c, []) for c in ("firstorder", "glcm", "gldm", "glrlm", "glszm", "ngtdm"))
        self.enabledFeatures = OrderedDict((k, list(v) if v else []) for k, v in feature_class.items())
        self.strict = strict
        self.skipped_image_types = []
        self.skipped_classes = []
        self._validate()

    def _complain(self, msg):
        if self.strict:
            raise NotImplementedError(msg + " (pass strict=False to skip them explicitly)")
        logger.error(msg)
        warnings.warn(msg, RuntimeWarning, stacklevel=4)

    def _validate(self):
        s = self.settings
        for k in _UNSUPPORTED_IF_SET:
            if s.get(k) not in (None, False):
                raise NotImplementedError("setting %r=%r is not implemented by the B200 engine" % (k, s[k]))
        if int(s.get("minimumROIDimensions", 2)) != 2:
            # the kernels hard-code pyradiomics' default (ROI must span both axes: status 3 otherwise)
            raise NotImplementedError("minimumROIDimensions=%r is not implemented (only the default 2)" % s["minimumROIDimensions"])
        if self._given.get("additionalInfo") is True:
            # pyradiomics would add non-numeric diagnostics_* keys; the reference switches them off (params.yml:62)
            raise NotImplementedError("additionalInfo: True (diagnostics_* keys) is not implemented; the reference sets it False")
        if not float(s["binWidth"]) > 0:
            raise ValueError("binWidth must be > 0")
        if s.get("binCount") is not None and not (1 <= int(s["binCount"]) <= 256):
            raise NotImplementedError("binCount must be 1..256")
        skipped_types = [t for t in self.enabledImagetypes if t not in SUPPORTED_IMAGE_TYPES]
        if skipped_types:
            self._complain("image types %s are enabled but not implemented by the B200 engine" % skipped_types)
        self.skipped_image_types = skipped_types
        self.image_types = [t for t in self.enabledImagetypes if t in SUPPORTED_IMAGE_TYPES]
        if not self.image_types:
            raise NotImplementedError("no implemented image type is enabled")
        self.blocks = self._image_blocks()
        skipped_cls = [c for c in self.enabledFeatures if c not in SUPPORTED_CLASSES]
        if skipped_cls:
            self._complain("feature classes %s are enabled but not implemented by the B200 engine" % skipped_cls)
        self.skipped_classes = skipped_cls
        self.classes = [c for c in self.enabledFeatures if c in SUPPORTED_CLASSES]
        # pyradiomics featureextractor.computeShape: the "force2D must be True" rule belongs to 3-D input; the
        # reference passes 2-D images (RadiomicExtractor.py:31,36), for which shape2D is computed regardless
        if not self.classes:
            raise ValueError("no implemented feature class is enabled")
        for c in self.classes:
            unknown = [f for f in self.enabledFeatures[c] if f not in FEATURE_NAMES[c]]
            if unknown:
                raise ValueError("unknown / deprecated features for class %s: %s" % (c, unknown))

    def _image_blocks(self):
        """One entry per filtered image pyradiomics would yield, in its order (imageType file order; the images of one
        type in the order of imageoperations.get*Image): ``(feature-name prefix, image type, argument)``."""
        s = self.settings
        blocks = []
        for t in self.image_types:
            opt = self.enabledImagetypes.get(t) or {}
            if t == "Wavelet":
                if opt.get("wavelet", "coif1") != "coif1" or int(opt.get("level", 1)) != 1 or int(opt.get("start_level", 0)) != 0:
                    raise NotImplementedError("Wavelet: only the pyradiomics defaults (coif1, level 1, start_level 0) are implemented")
                if s.get("force2D", False):
                    # getWaveletImage: axes = [1, 0] minus force2Ddimension (the same axis removal as the texture
                    # angles, oracle/U1_ANGLES.md) -> a 1-D transform
                    if int(s.get("force2Ddimension", 0)) != 0:
                        raise NotImplementedError("Wavelet with force2Ddimension != 0 is not implemented")
                    blocks += [("wavelet-H", t, 0), ("wavelet-L", t, 1)]
                else:
                    blocks += [("wavelet-%s" % b, t, k) for k, b in enumerate(("LH", "HL", "HH", "LL"))]
            elif t == "LoG":
                sig = opt.get("sigma")
                if not sig:
                    raise ValueError("LoG: no sigma values given (pyradiomics yields no image without them)")
                for v in sig:
                    if not float(v) > 0:
                        raise ValueError("LoG: sigma must be > 0")
                    blocks.append(("log-sigma-%s-mm-3D" % str(v).replace(".", "-"), t, float(v)))
            elif t == "Gradient":
                if opt.get("gradientUseSpacing", s.get("gradientUseSpacing", True)) is not True:
                    pass  # spacing is (1, 1) for GetImageFromArray images: same result either way
                blocks.append(("gradient", t, None))
            else:
                blocks.append((t.lower(), t, None))
        return blocks

    # ---- resolved views
    @property
    def label(self):
        return int(self.settings["label"])

    @property
    def bin_width(self):
        return float(self.settings["binWidth"])

    @property
    def bin_count(self):
        """pyradiomics: when binCount is set it takes precedence over binWidth (imageoperations.getBinEdges)."""
        return int(self.settings["binCount"] or 0)

    def angles(self, ndim=2):
        s = self.settings
        return in_plane_angles(ndim, s["distances"], bool(s["force2D"]), int(s["force2Ddimension"]))

    def _class_features(self, c):
        sel = self.enabledFeatures[c]
        return [f for f in FEATURE_NAMES[c] if not sel or f in sel]

    def feature_names(self):
        """Output keys in pyradiomics order: shape descriptors first (A.1 step 3, always ``original_``),
        then for every enabled image type (file order) every class (file order) in A.2 order."""
        names = []
        if "shape2D" in self.classes:
            names += ["original_shape2D_%s" % f for f in self._class_features("shape2D")]
        for name, _, _ in self.blocks:
            for c in self.classes:
                if c != "shape2D":
                    names += ["%s_%s_%s" % (name, c, f) for f in self._class_features(c)]
        return names

    def engine_columns(self):
        """(engine classes, column permutation): the engine emits, per image type, the classes in its fixed
        order with all their features; the extractor concatenates [shape | block(type 1) | block(type 2) ...].
        ``perm`` maps that super-row to ``feature_names()``."""
        eng_classes = [c for c in ("shape2D",) + tuple(CLASS_ORDER) if c in self.classes]
        eng_names = ["original_shape2D_%s" % f for f in FEATURE_NAMES["shape2D"]] if "shape2D" in eng_classes else []
        for name, _, _ in self.blocks:
            eng_names += ["%s_%s_%s" % (name, c, f) for c in eng_classes if c != "shape2D" for f in FEATURE_NAMES[c]]
        pos = {n: i for i, n in enumerate(eng_names)}
        return eng_classes, [pos[n] for n in self.feature_names()]
