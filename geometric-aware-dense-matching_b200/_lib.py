"""ctypes binding of libgadm.so (C ABI in include/gadm.h).

There is NO CPU fallback: if the shared library is missing or the device is not a B200 the product
path raises.  This is the same binding a maintainer of the reference would add."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgadm.so")

c_void_p, c_int, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t


class KnnJob(ctypes.Structure):
    """gadm_knn_job (include/gadm.h)."""
    _fields_ = [("support_off", ctypes.c_int64), ("query_off", ctypes.c_int64), ("out_off", ctypes.c_int64),
                ("support_bstride", ctypes.c_int64), ("query_bstride", ctypes.c_int64),
                ("out_bstride", ctypes.c_int64),
                ("n_support", ctypes.c_int32), ("n_query", ctypes.c_int32), ("k", ctypes.c_int32),
                ("batch", ctypes.c_int32)]


# every symbol include/gadm.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gadm_strerror": (ctypes.c_char_p, [c_int]),
    "gadm_abi_version": (c_int, []),
    "gadm_init": (c_int, [c_int]),
    "gadm_config_set": (c_int, [ctypes.c_char_p, c_int]),
    "gadm_last_cuda_error": (ctypes.c_char_p, []),
    "gadm_operand_k": (c_int, [c_int, c_int]),
    "gadm_aux_floats": (c_size_t, [c_int, c_int]),
    "gadm_prep_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gadm_prep_rows_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gadm_compact_rows": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gadm_prep_rows_sel": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "gadm_match_fwd_sel": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                   c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gadm_pack_match_outputs": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_void_p, c_void_p]),
    "gadm_pack_indices_u16": (c_int, [c_void_p, ctypes.c_int64, c_void_p, c_void_p]),
    "gadm_prep_model": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "gadm_match_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gadm_match_workspace_bytes": (c_size_t, []),
    "gadm_circle_loss_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "gadm_circle_loss_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "gadm_circle_loss_bwd_split": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "gadm_circle_loss_bwd_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                           c_void_p]),
    "gadm_kabsch_moments": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                    c_void_p, c_void_p]),
    "gadm_kabsch_moments_w": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                      c_int, c_void_p, c_void_p]),
    "gadm_kabsch_poses": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "gadm_knn3d_workspace_bytes": (c_size_t, [ctypes.POINTER(KnnJob), c_int, c_int]),
    "gadm_knn3d": (c_int, [c_void_p, c_void_p, ctypes.POINTER(KnnJob), c_int, c_int, c_void_p, c_void_p,
                           c_void_p, c_size_t, c_void_p]),
    "gadm_knn_feat": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gadm_knn_feat_tc_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "gadm_knn_feat_tc": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gadm_graph_feature_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "gadm_graph_feature": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                   c_void_p]),
    "gadm_group_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gadm_group_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gadm_gather_neighbour": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gadm_gather_max": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gadm_relative_pos_encoding": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gadm_graph_feature_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gadm_gather_neighbour_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gadm_gather_max_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gadm_seg_mask": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
}

PAD_MODES = {"none": 0, "minus_one": 1, "e0": 2}
OPERAND_MODES = {"bf16": 0, "bf16x3": 1, "bf16n": 2}
MATCH_MODES = {"argmax": 0, "soft": 1, "argmax_unit": 2, "argmax_bf16n": 3}
KNN_ALGOS = {"brute": 0, "grid": 1, "auto": 2}

_lib = None
_lock = threading.Lock()
_inited = set()


class GadmError(RuntimeError):
    pass


def load():
    """dlopen libgadm.so and type every entry point.  Raises if the library is not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise GadmError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(or make -C geometric-aware-dense-matching_b200/csrc). There is no CPU fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)       # AttributeError here == header/library mismatch
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        lib = load()
        msg = lib.gadm_strerror(rc).decode()
        if rc == -5:
            msg += ": " + lib.gadm_last_cuda_error().decode()
        raise GadmError(f"{what} failed ({rc}): {msg}")


def config_set(key: str, value: int):
    """gadm_config_set: kernel-selection switches for profiling and tests (-1 = automatic)."""
    check(load().gadm_config_set(key.encode(), int(value)), f"gadm_config_set({key})")


def ensure_init(device_index):
    """gadm_init once per device per process (fails loudly on anything that is not sm_100)."""
    lib = load()
    if device_index not in _inited:
        check(lib.gadm_init(int(device_index)), "gadm_init")
        _inited.add(device_index)
    return lib
