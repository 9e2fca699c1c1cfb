"""CPU (-m "not gpu"): the oracle is pinned against fixtures generated FROM THE REFERENCE ITSELF
(tests/golden/make_golden.py) and, where oracle/_ref exists, against the compiled reference live."""
import os

import numpy as np
import pytest
import torch

from oracle import dgcnn_oracle as do
from oracle import knn_oracle as ko
from oracle import match_oracle as mo
from oracle import pointops_oracle as po

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", ["uniform", "self", "one_nn", "dups"])
def test_knn_port_vs_reference_golden(case):
    z = np.load(os.path.join(G, "knn_golden.npz"))
    sup, qry, k = z[case + "_support"], z[case + "_query"], int(z[case + "_k"])
    idx, d2 = ko.knn_port(sup, qry, k, return_dist=True)
    # sorted distance vectors are bit-identical to the reference's on every row, ties included
    assert np.array_equal(d2, z[case + "_ref_d2"])
    # index rows identical wherever no exact distance tie exists among the first k+1 candidates
    if k < sup.shape[1]:
        _, d2k1 = ko.knn_port(sup, qry, k + 1, return_dist=True)
        free = ko.tie_free_rows(d2k1)
    else:
        free = np.ones(idx.shape[:2], dtype=bool)
    assert np.array_equal(idx[free], z[case + "_ref_idx"][free])
    if case != "dups":
        assert free.all()
    else:
        assert (~free).any(), "the duplicate case must exercise ties"
        # on tied rows the index SETS still agree once the rank-k boundary tie group is removed
        ref = z[case + "_ref_idx"]
        for b, q in zip(*np.nonzero(~free)):
            dk = d2[b, q, -1]
            keep = d2[b, q] < dk
            assert set(idx[b, q][keep]) == set(ref[b, q][z[case + "_ref_d2"][b, q] < dk])


def test_knn_compiled_reference_matches_golden():
    if not ko.have_reference():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    z = np.load(os.path.join(G, "knn_golden.npz"))
    for case in ("uniform", "self", "one_nn", "dups"):
        ref = ko.knn_reference(z[case + "_support"], z[case + "_query"], int(z[case + "_k"]))
        assert np.array_equal(ref, z[case + "_ref_idx"])
    out = ko.knn_search_ref(z["uniform_support"], z["uniform_query"], 16)
    assert out.dtype == np.int32 and out.shape == (2, 120, 16)          # helper_tool.py:170


def test_knn_port_properties_and_errors():
    rng = np.random.default_rng(5)
    s = rng.random((1, 64, 3), dtype=np.float32)
    idx, d2 = ko.knn_port(s, s, 64, return_dist=True)                   # k == n_support: a permutation per row
    assert np.all(np.sort(idx, -1) == np.arange(64))
    assert np.all(np.diff(d2, axis=-1) >= 0) and np.all(idx[..., 0] == np.arange(64))
    with pytest.raises(ValueError):
        ko.knn_port(s, s, 65)                                           # k > n_support refused
    # many exact ties: a lattice; order must be (d2, index) lexicographic
    g = np.stack(np.meshgrid(*[np.arange(4, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(1, -1, 3)
    idx, d2 = ko.knn_port(g, g, 7, return_dist=True)
    for q in range(g.shape[1]):
        pairs = list(zip(d2[0, q], idx[0, q]))
        assert pairs == sorted(pairs)


def test_knn_schedule_is_22_calls():
    from gadm_b200 import synth
    cld, sr = synth.depth_cloud(64, 3200, 3)
    calls = ko.schedule(cld, sr)
    assert len(calls) == 22
    assert [c[3] for c in calls[:4]] == [16, 1, 16, 1]
    assert calls[0][1].shape == (3200, 3) and calls[4][1].shape == (800, 3) and calls[12][1].shape == (50, 3)
    assert calls[2][1].shape == (256, 3)            # sr2dptxyz[4] at in_size 64
    assert calls[-1][1].shape == (800, 3) and calls[-1][2].shape == (1024, 3)


def test_match_oracle_vs_reference_lines():
    z = np.load(os.path.join(G, "match_golden.npz"))
    rgbd, mesh, seg = (torch.from_numpy(z[k]) for k in ("rgbd", "mesh", "seg"))
    mask = mo.seg_mask(seg)
    assert np.array_equal(mask.numpy(), z["live_mask"])
    idx, mx, _ = mo.match_hard(rgbd, mesh, row_mask=mask)
    assert np.array_equal(idx.numpy(), z["live_idx"])                   # evaluator.py:93
    assert np.array_equal(mx.numpy(), z["live_max"])
    assert np.array_equal(mo.similarity(rgbd, mesh, mask).numpy(), z["live_sim"])
    pidx, _, _ = mo.match_hard(rgbd, mesh, pad_mode="minus_one")
    assert np.array_equal(pidx.numpy(), z["pad_idx"])                   # pvn3d_eval_utils_kpls.py:437-441
    # (the fixture was produced from a contiguous transposed copy: same maths, last-ulp GEMM differences)
    assert np.allclose(mo.similarity(rgbd, mesh, None, "minus_one").numpy(), z["pad_sim"], rtol=0, atol=3e-7)
    assert (z["pad_idx"] == mesh.shape[1]).sum() >= 4, "fixture must exercise the pad column"


def test_match_soft_extension_definition():
    g = torch.Generator().manual_seed(0)
    rgbd, mesh = torch.randn((32, 50), generator=g), torch.randn((32, 40), generator=g)
    xyz = torch.randn((40, 3), generator=g)
    out = mo.match_soft(rgbd, mesh, xyz, gamma=16.0)
    S = mo.similarity(rgbd, mesh)
    w = torch.softmax(16.0 * S, dim=1)
    assert torch.allclose(out["weight"], w.max(1).values) and torch.allclose(out["soft_xyz"], w @ xyz)
    assert torch.equal(out["idx"], S.argmax(1))
    o64 = mo.match_soft(rgbd, mesh, xyz, dtype=torch.float64)
    assert (o64["weight"].float() - out["weight"]).abs().max() < 1e-5
    # e0 pad (geoMatch_DGCNN.py:95-98): similarity with the pad column is the normalised first channel
    Sp = mo.similarity(rgbd, mesh, None, "e0")
    f_hat = torch.nn.functional.normalize(rgbd.t(), dim=1)
    assert torch.allclose(Sp[:, -1], f_hat[:, 0], atol=1e-6)


def test_kabsch_oracle_recovers_transform():
    g = torch.Generator().manual_seed(1)
    A = torch.randn((50, 3), generator=g)
    R = torch.linalg.qr(torch.randn((3, 3), generator=g))[0]
    if torch.linalg.det(R) < 0:
        R[:, 0] *= -1
    t = torch.tensor([0.1, 0.2, 0.3])
    T = mo.best_fit_transform(A, A @ R.T + t)
    assert torch.allclose(T[:, :3].float(), R, atol=1e-5) and torch.allclose(T[:, 3].float(), t, atol=1e-5)


def test_dgcnn_oracle_vs_reference_module():
    z = np.load(os.path.join(G, "dgcnn_golden.npz"))
    x, x9 = torch.from_numpy(z["x"]), torch.from_numpy(z["x9"])
    assert np.array_equal(do.knn(x, 20).numpy(), z["knn_idx"])                       # dgcnn.py:21-27
    assert np.array_equal(do.get_graph_feature(x, k=20).numpy(), z["graph"])         # dgcnn.py:30-56
    assert np.array_equal(do.get_graph_feature(x9, k=16, dim9=True).numpy(), z["graph9"])
    idx, gaps, vals = do.knn_with_gaps(x, 20)
    assert torch.equal(idx, do.knn(x, 20)) and gaps.shape == (2, 200) and (gaps >= 0).all()


def test_pointops_oracle_semantics():
    g = torch.Generator().manual_seed(2)
    xyz, new = torch.rand((2, 50, 3), generator=g), torch.rand((2, 20, 3), generator=g)
    idx, d = po.knnquery_naive(4, xyz, new)
    assert idx.dtype == torch.int32 and idx.shape == (2, 20, 4)
    assert np.array_equal(idx.numpy().astype(np.int64), ko.knn_port(xyz.numpy(), new.numpy(), 4))
    f = torch.randn((2, 5, 50), generator=g)
    out = po.grouping(f, idx)
    assert out.shape == (2, 5, 20, 4) and out[1, 3, 7, 2] == f[1, 3, idx[1, 7, 2]]
    go = torch.randn((2, 5, 20, 4), generator=g)
    gf = po.grouping_backward(go, idx, 50)
    f2 = f.clone().requires_grad_(True)
    po.grouping(f2, idx).backward(go)
    assert torch.allclose(gf, f2.grad)


def test_pointops_oracle_vs_reference_source():
    """oracle/pointops_oracle.py against pointops_golden.npz: the outputs of KNNQueryNaive.forward
    (lib/pointops/functions/pointops.py:396-426) and QueryAndGroup.forward (:548-585) executed from the reference text."""
    z = np.load(os.path.join(G, "pointops_golden.npz"))
    t = lambda k: torch.from_numpy(z[k])
    k = int(z["k"])
    assert float(z["knn_gap"]) > 0, "fixture must be tie-free among the first k + 1 distances"
    assert torch.equal(po.knnquery_naive(k, t("xyz"), t("new_xyz"))[0], t("knn_idx"))
    assert torch.equal(po.knnquery_naive(k, t("xyz"))[0], t("knn_self_idx"))
    nf, gx, idx = po.query_and_group(k, t("xyz"), t("new_xyz"), t("features"))
    assert torch.equal(idx, t("qg_idx")) and torch.equal(gx, t("qg_grouped_xyz")) and torch.equal(nf, t("qg_new_features"))
    assert torch.equal(po.query_and_group(k, t("xyz"), t("new_xyz"), t("features"), use_xyz=False)[0], t("qg_features_only"))
    assert torch.equal(po.query_and_group(k, t("xyz"), t("new_xyz"))[0], t("qg_xyz_only"))
    assert torch.equal(po.grouping(t("features"), t("knn_idx")), t("qg_new_features")[:, 3:])


def test_randla_oracle_vs_reference_source():
    """oracle/randla_oracle.py against the outputs of the reference's own function bodies (randla_golden.npz)."""
    from oracle import randla_oracle as ro
    z = np.load(os.path.join(G, "randla_golden.npz"))
    t = lambda k: torch.from_numpy(z[k])
    assert torch.equal(ro.random_sample(t("feature"), t("pool_idx")), t("random_sample"))
    assert torch.equal(ro.nearest_interpolation(t("feature"), t("interp_idx")), t("nearest_interpolation"))
    assert torch.equal(ro.gather_neighbour(t("xyz"), t("neigh_idx")), t("gather_neighbour"))
    assert torch.equal(ro.relative_pos_encoding(t("xyz"), t("neigh_idx")), t("relative_pos_encoding"))


def test_circle_oracle_reproduces_reference_loss():
    """oracle/circle_oracle.py against the outputs of the reference's own CircleLoss / matching_loss /
    pointwise_feature_matching (tests/golden/circle_golden.npz, generated by make_golden.py from /root/reference)."""
    from oracle import circle_oracle as co
    g = np.load(os.path.join(G, "circle_golden.npz"))
    t = lambda k: torch.from_numpy(g[k])
    r = float(g["positive_r"])
    total = co.batch_loss(t("rgbd"), t("mesh")[0], t("labels"), t("match_idx"), t("xyz"), t("vis"), r)
    assert abs(float(total) - float(g["ref_total"])) <= 1e-6 * abs(float(g["ref_total"]))
    for b in range(2):                                       # sample 2 has < 3 foreground rows: skipped (geoMatch.py:128)
        one = co.batch_loss(t("rgbd")[b:b + 1], t("mesh")[0], t("labels")[b:b + 1], t("match_idx")[b:b + 1],
                            t("xyz"), t("vis")[b:b + 1], r)
        assert abs(float(one) - float(g["ref_per_sample"][b])) <= 1e-6 * abs(float(g["ref_per_sample"][b]))
    assert float(co.batch_loss(t("rgbd")[2:], t("mesh")[0], t("labels")[2:], t("match_idx")[2:], t("xyz"),
                               t("vis")[2:], r)) == 0.0
    # the DGCNN variant (geoMatch_DGCNN.py:53-78, :80-136): e0 pad column, per-vertex radius positive_r / 1000 * depth
    rad = torch.stack([co.dgcnn_radius(t("xyz"), t("RT")[b], float(g["dgcnn_positive_r"])) for b in range(3)])
    tot2 = co.batch_loss(t("rgbd"), t("mesh")[0], t("labels"), t("match_idx"), t("xyz"), t("vis"), rad, pad="e0")
    assert abs(float(tot2) - float(g["dgcnn_ref_total"])) <= 1e-6 * abs(float(g["dgcnn_ref_total"]))
    # the symmetry-aware branch (geoMatch.py:138-141 -> matching_loss_sys :86-100)
    tot3 = co.batch_loss_sys(t("rgbd"), t("mesh"), t("labels"), t("match_idx"), t("sys_idx"))
    assert abs(float(tot3) - float(g["sys_ref_total"])) <= 1e-6 * abs(float(g["sys_ref_total"]))
    # the mask the loss sees: every on-model row has its own (visible) ground-truth vertex as a positive, an
    # off-model row has the pad column only
    mi = t("match_idx")[0].long()
    mask = co.positive_mask(mi, t("xyz"), t("vis")[0], r)
    on = mi != 64
    assert torch.all(mask[on, :64].gather(1, mi[on][:, None])) and not mask[on, 64].any()
    assert torch.all(mask[~on, 64]) and not mask[~on, :64].any()
