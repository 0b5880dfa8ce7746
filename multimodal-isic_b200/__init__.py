"""multimodal-isic_b200 -- B200-native radiomic feature engine, a drop-in for the
``RadiomicExtractor.py`` / ``extract_radiomics.py`` path of rbuler/multimodal-isic.
Import it as ``multimodal_isic_b200`` (the hyphenated directory name is not an identifier)."""
from ._abi import CLASS_ORDER, EXPORTED_SYMBOLS, LIB_PATH, load_library  # noqa: F401
from .settings import FEATURE_NAMES, Settings, in_plane_angles  # noqa: F401
from .engine import Engine, HostPipeline, RadbError, pack_ragged  # noqa: F401
from .extractor import RadiomicsExtractor, features_to_dataframe  # noqa: F401
from .sharded import OverlappedGather, all_gather_rows, shard_bounds, sharded_extract  # noqa: F401
from . import numa, synth  # noqa: F401
