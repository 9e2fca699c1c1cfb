"""TEST INFRASTRUCTURE ONLY -- Python face of the 3-D kNN oracle.

  knn_port(...)       our plain-C restatement (oracle/knn_oracle.c), ties by ascending index.
  knn_reference(...)  the reference's OWN code: knn_.cxx + nanoflann.hpp compiled from /root/reference
                      into oracle/_ref/libref_knn.so (oracle/Makefile), called exactly as
                      knn.pyx:71-109 calls it (omp=True -> cpp_knn_batch_omp, knn_.cxx:104-135).
  knn_search_ref(...) DataProcessing.knn_search (helper_tool.py:161-170): reference + astype(int32).
  schedule(...)       the 22-call per-sample kNN schedule of datasets/lm/linemod_pbr.py:534-569.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(_HERE, "_build", "liboracle_knn.so")
_REF = os.path.join(_HERE, "_ref", "libref_knn.so")


def build(force=False):
    """Compile the C restatement and (when /root/reference is present) the reference itself."""
    if force or not os.path.exists(_PORT) or \
            os.path.getmtime(_PORT) < os.path.getmtime(os.path.join(_HERE, "knn_oracle.c")):
        subprocess.run(["make", "-s", "-C", _HERE, "_build/liboracle_knn.so"], check=True)
    if os.path.isdir("/root/reference") and (force or not os.path.exists(_REF)):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


_port_lib = None
_ref_lib = None


def _port():
    global _port_lib
    if _port_lib is None:
        build()
        lib = ctypes.CDLL(_PORT)
        f32p, i64p, sz = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64), ctypes.c_size_t
        lib.oracle_knn_batch.argtypes = [f32p, sz, sz, f32p, sz, sz, i64p, f32p]
        lib.oracle_knn_batch.restype = ctypes.c_int
        lib.oracle_knn_dist2_of.argtypes = [f32p, sz, sz, f32p, sz, sz, i64p, f32p]
        lib.oracle_knn_dist2_of.restype = ctypes.c_int
        _port_lib = lib
    return _port_lib


def have_reference():
    if not os.path.exists(_REF) and os.path.isdir("/root/reference"):
        build()
    return os.path.exists(_REF)


def _ref():
    global _ref_lib
    if _ref_lib is None:
        if not have_reference():
            raise RuntimeError("oracle/_ref/libref_knn.so missing (build it where /root/reference exists)")
        lib = ctypes.CDLL(_REF)
        f32p, sz = ctypes.POINTER(ctypes.c_float), ctypes.c_size_t
        lib.gadm_ref_knn_batch.argtypes = [f32p, sz, sz, sz, f32p, sz, sz, ctypes.POINTER(ctypes.c_long),
                                           ctypes.c_int]
        lib.gadm_ref_knn_batch.restype = None
        _ref_lib = lib
    return _ref_lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 3 and a.shape[2] == 3, a.shape
    return a


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def knn_port(support_pts, query_pts, k, return_dist=False):
    """support [B,N1,3], query [B,N2,3] -> int64 [B,N2,k] (+ fp32 d2), (d2, idx)-lexicographic."""
    s, q = _f32(support_pts), _f32(query_pts)
    B, ns, _ = s.shape
    nq = q.shape[1]
    idx = np.zeros((B, nq, k), dtype=np.int64)
    d2 = np.zeros((B, nq, k), dtype=np.float32)
    rc = _port().oracle_knn_batch(_fp(s), B, ns, _fp(q), nq, k,
                                  idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _fp(d2))
    if rc != 0:
        raise ValueError("oracle_knn_batch: bad arguments (need 1 <= k <= min(N1, 64))")
    return (idx, d2) if return_dist else idx


def dist2_of(support_pts, query_pts, idx):
    """Reference-metric d2 of given neighbour indices [B,N2,k]."""
    s, q = _f32(support_pts), _f32(query_pts)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    B, nq, k = idx.shape
    d2 = np.zeros((B, nq, k), dtype=np.float32)
    rc = _port().oracle_knn_dist2_of(_fp(s), B, s.shape[1], _fp(q), nq, k,
                                     idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _fp(d2))
    if rc != 0:
        raise ValueError("index out of range")
    return d2


def knn_reference(support_pts, query_pts, k, omp=True):
    """The reference's knn_batch (knn.pyx:71-109): int64 [B,N2,k]."""
    s, q = _f32(support_pts), _f32(query_pts)
    B, ns, dim = s.shape
    nq = q.shape[1]
    out = np.zeros((B, nq, k), dtype=np.int64)                      # knn.pyx:93
    _ref().gadm_ref_knn_batch(_fp(s), B, ns, dim, _fp(q), nq, k,
                              out.ctypes.data_as(ctypes.POINTER(ctypes.c_long)), 1 if omp else 0)
    return out


def knn_search_ref(support_pts, query_pts, k):
    """DataProcessing.knn_search, helper_tool.py:161-170."""
    return knn_reference(support_pts, query_pts, k, omp=True).astype(np.int32)


def tie_free_rows(d2_kplus1):
    """d2 [.., k+1] ascending (k+1 nearest) -> bool[..]: no two equal adjacent distances, i.e. the
    row's index output is fully determined by the data."""
    return np.all(d2_kplus1[..., 1:] != d2_kplus1[..., :-1], axis=-1)


def schedule(cld, sr2dptxyz, k_nei=16):
    """The 22 kNN calls of one sample, in the reference's order (linemod_pbr.py:528-569).
    cld [N,3]; sr2dptxyz {1,2,4,8: [(in_size/s)^2, 3]}.  Yields (name, support, query, k)."""
    rgb_ds_sr = [4, 8, 8, 8]                                       # :528
    calls, xyz_lvl = [], []
    for i in range(4):                                             # :533
        calls.append(("cld_nei_idx%d" % i, cld, cld, k_nei))       # :534-536
        sub = cld[: cld.shape[0] // 4]                             # :537
        calls.append(("cld_interp_idx%d" % i, sub, cld, 1))        # :539-541
        xyz_lvl.append(cld)
        calls.append(("r2p_ds_nei_idx%d" % i, sr2dptxyz[rgb_ds_sr[i]], sub, k_nei))   # :546-548
        calls.append(("p2r_ds_nei_idx%d" % i, sub, sr2dptxyz[rgb_ds_sr[i]], 1))       # :550-552
        cld = sub                                                  # :554
    rgb_up_sr = [4, 2, 2]                                          # :557
    for i in range(3):
        lvl = xyz_lvl[4 - i - 1]
        calls.append(("r2p_up_nei_idx%d" % i, sr2dptxyz[rgb_up_sr[i]], lvl, k_nei))   # :559-562
        calls.append(("p2r_up_nei_idx%d" % i, lvl, sr2dptxyz[rgb_up_sr[i]], 1))       # :564-567
    return calls
