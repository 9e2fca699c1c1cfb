"""gadm_b200 -- B200 (sm_100a) implementation of the scene-to-model dense correspondence path of
Ray0089/geometric-aware-dense-matching: the geoMatch matching head (fused tcgen05/TMA similarity + argmax /
softmax / soft coordinates) and the geometric kNN that builds the network's neighbourhoods.

The directory is named `geometric-aware-dense-matching_b200/`; `import gadm_b200` (repo-root shim) loads it.
All compute goes through libgadm.so (C ABI: include/gadm.h).  No CPU fallback exists."""
from ._lib import GadmError, LIB_PATH, load as load_library  # noqa: F401

__all__ = ["GadmError", "LIB_PATH", "load_library", "matching", "knn", "dgcnn", "pointops", "sharding", "synth",
           "ops", "pipeline", "randla"]


def __getattr__(name):  # lazy: importing the package must not need CUDA (CPU-side tests import synth/sharding)
    if name in ("matching", "knn", "dgcnn", "pointops", "sharding", "synth", "ops", "pipeline", "randla"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
