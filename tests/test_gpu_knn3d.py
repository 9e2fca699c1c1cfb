"""-m gpu: CUDA 3-D kNN (through the C ABI) vs the oracle port and, where present, the compiled reference."""
import numpy as np
import pytest
import torch

from oracle import knn_oracle as ko

pytestmark = pytest.mark.gpu


def _run(cuda, s, q, k, algo, return_dist=True):
    from gadm_b200 import ops
    B, n1, _ = s.shape
    n2 = q.shape[1]
    st, qt = torch.from_numpy(s).to(cuda), torch.from_numpy(q).to(cuda)
    jobs = ops.make_jobs([(0, 0, 0, n1, n2, n2 * k, n1, n2, k, B)])
    idx, d2 = ops.knn3d_jobs(st.view(-1, 3), qt.view(-1, 3), jobs, B * n2 * k, algo, return_dist=True)
    torch.cuda.synchronize()
    return idx.view(B, n2, k).cpu().numpy(), d2.view(B, n2, k).cpu().numpy()


def _clouds(seed, B, n1, n2, kind):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        s = rng.random((B, n1, 3), dtype=np.float32)
        q = rng.random((B, n2, 3), dtype=np.float32)
    elif kind == "surface":
        from gadm_b200 import synth
        s = np.stack([synth.depth_cloud(128, n1, seed + b)[0] for b in range(B)])
        q = s[:, :n2].copy() if n2 <= n1 else np.concatenate([s, s[:, : n2 - n1]], 1)
    elif kind == "dups":            # wrap-padded duplicates as datasets/lm/linemod_pbr.py:492 produces
        u = rng.random((B, (n1 * 3) // 4, 3), dtype=np.float32)
        s = np.concatenate([u, u[:, : n1 - u.shape[1]]], 1)
        q = s[:, :n2].copy()
    elif kind == "lattice":         # many exact distance ties
        g = np.stack(np.meshgrid(*[np.arange(12, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(-1, 3)
        s = np.stack([g[rng.permutation(len(g))[:n1]] for _ in range(B)])
        q = s[:, :n2].copy()
    return np.ascontiguousarray(s), np.ascontiguousarray(q)


@pytest.mark.parametrize("algo", ["brute", "grid"])
@pytest.mark.parametrize("kind,B,n1,n2,k", [
    ("uniform", 2, 1000, 700, 16), ("uniform", 1, 4096, 4096, 16), ("uniform", 3, 333, 1001, 1),
    ("surface", 2, 3200, 3200, 16), ("dups", 2, 2048, 2048, 16), ("lattice", 1, 1500, 1500, 8),
    ("uniform", 1, 50, 200, 16), ("uniform", 1, 33, 5, 32), ("surface", 1, 12800, 3200, 1),
])
def test_knn3d_bit_exact_vs_port(cuda, algo, kind, B, n1, n2, k):
    """Index AND distance output identical to the (d2, index)-lexicographic oracle, ties included."""
    s, q = _clouds(7 + n1, B, n1, n2, kind)
    idx, d2 = _run(cuda, s, q, k, algo)
    ref_idx, ref_d2 = ko.knn_port(s, q, k, return_dist=True)
    assert np.array_equal(d2, ref_d2)
    assert np.array_equal(idx.astype(np.int64), ref_idx)


@pytest.mark.parametrize("algo", ["brute", "grid"])
def test_knn3d_vs_compiled_reference(cuda, algo):
    """Against the reference's own nanoflann code (oracle/_ref): identical rows wherever no fp32 distance tie
    occurs among the first k+1 candidates; bit-identical sorted distance vectors on every row."""
    if not ko.have_reference():
        pytest.skip("oracle/_ref/libref_knn.so not built")
    from gadm_b200 import synth
    k = 16
    for dup in (0.0, 0.1):
        cld, _ = synth.depth_cloud(128, 12800, 1000, dup_frac=dup)
        s = cld[None]
        idx, d2 = _run(cuda, s, s, k, algo)
        ref = ko.knn_reference(s, s, k)
        ref_d2 = ko.dist2_of(s, s, ref)
        assert np.array_equal(d2, ref_d2), "sorted distance vectors must be bit-identical"
        _, d2k1 = ko.knn_port(s, s, k + 1, return_dist=True)
        free = ko.tie_free_rows(d2k1)
        assert np.array_equal(idx[free].astype(np.int64), ref[free])
        if dup == 0.0:
            assert free.mean() > 0.99


def test_knn_search_signature(cuda):
    """DataProcessing.knn_search: numpy in, np.int32 [B, N2, k] out (helper_tool.py:161-170)."""
    from gadm_b200.knn import DataProcessing
    s, q = _clouds(3, 2, 800, 300, "uniform")
    out = DataProcessing.knn_search(s, q, 16)
    assert out.dtype == np.int32 and out.shape == (2, 300, 16)
    assert np.array_equal(out.astype(np.int64), ko.knn_port(s, q, 16))


def test_knn_pyramid_22_calls(cuda):
    """The whole per-sample schedule (linemod_pbr.py:534-569) in one call == 22 oracle calls."""
    from gadm_b200 import synth
    from gadm_b200.knn import KnnPyramid
    B, N = 2, 3200
    cld, sr = synth.frame_batch(B, 64, N, seed=11)
    pyr = KnnPyramid(N, {s: sr[s].shape[1] for s in (2, 4, 8)}, B)
    out = pyr(cld.to(cuda), {s: v.to(cuda) for s, v in sr.items()})
    torch.cuda.synchronize()
    n_checked = 0
    for b in range(B):
        calls = ko.schedule(cld[b].numpy(), {s: v[b].numpy() for s, v in sr.items()})
        assert len(calls) == 22
        for name, sup, qry, k in calls:
            ref = ko.knn_port(sup[None], qry[None], k)[0]
            assert np.array_equal(out[name][b].cpu().numpy().astype(np.int64), ref), name
            n_checked += 1
    assert n_checked == 44
    assert out["cld_sub_idx0"].shape == (B, N // 4, 16)


def test_knn3d_full_size_properties(cuda):
    """BASELINE size (12800 x 12800, k=16): sortedness, self at distance 0, and grid == brute."""
    from gadm_b200 import synth
    cld, _ = synth.depth_cloud(128, 12800, 4242)
    s = cld[None]
    ib, db = _run(cuda, s, s, 16, "brute")
    ig, dg = _run(cuda, s, s, 16, "grid")
    assert np.array_equal(ib, ig) and np.array_equal(db, dg)
    assert np.all(np.diff(db, axis=-1) >= 0)
    assert np.all(db[..., 0] == 0)
    assert np.all((ib >= 0) & (ib < 12800))
    # every row's ids are distinct
    assert all(len(set(r)) == 16 for r in ib[0, ::97])


def test_knn3d_errors(cuda):
    from gadm_b200 import ops, _lib
    s = torch.rand(1, 10, 3, device=cuda)
    with pytest.raises(_lib.GadmError):
        ops.knn3d(s, s, 16, 2)          # k > n_support: the reference leaves stale ids, we refuse
    with pytest.raises(_lib.GadmError):
        ops.knn3d(torch.rand(1, 100, 3, device=cuda), s, 33, 2)   # k > 32 unsupported


@pytest.mark.parametrize("switches", [{"knn.ppc": 4}, {"knn.ppc": 40}, {"knn.grid_min": 1}, {"knn.grid_min": 100000}])
def test_knn3d_tuning_switches_do_not_change_results(cuda, switches):
    """gadm_config_set('knn.*'): cell size of the grid and the cloud size below which AUTO scans -- performance knobs
    only (the library reads no environment variable); the output stays bit-identical to the oracle."""
    from gadm_b200 import _lib
    s, q = _clouds(99, 2, 3000, 1500, "surface")
    ref_idx, ref_d2 = ko.knn_port(s, q, 16, return_dist=True)
    try:
        for k_, v in switches.items():
            _lib.config_set(k_, v)
        idx, d2 = _run(cuda, s, q, 16, "auto")
    finally:
        for k_ in switches:
            _lib.config_set(k_, -1)
    assert np.array_equal(d2, ref_d2) and np.array_equal(idx.astype(np.int64), ref_idx)
    with pytest.raises(_lib.GadmError):
        _lib.config_set("knn.ppc", 1000)
