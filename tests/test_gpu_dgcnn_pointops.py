"""-m gpu: DGCNN feature-space kNN + graph feature, pointops knnquery/grouping, RandLA gather_neighbour."""
import numpy as np
import pytest
import torch

from oracle import dgcnn_oracle as do
from oracle import pointops_oracle as po

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("B,C,N,k,dim9", [(2, 64, 1024, 20, False), (1, 9, 777, 16, True), (2, 3, 500, 20, False),
                                          (1, 128, 300, 32, False), (2, 128, 1000, 16, False), (1, 64, 4096, 20, False),
                                          (1, 256, 520, 5, False), (1, 64, 260, 1, False), (3, 192, 300, 20, False),
                                          (1, 64, 256, 20, False)])
def test_knn_feat_vs_oracle(cuda, B, C, N, k, dim9, tc):
    """Exact where the minimum adjacent-rank gap (ranks 1..k+1) exceeds the fp32 noise of the distance form
    (SURVEY.md 7.3-5): gap > 2e-4 * max(1, |pd|max / 79)."""
    from gadm_b200 import dgcnn
    g = torch.Generator().manual_seed(5000 + N)
    x = torch.randn((B, C, N), generator=g)
    xs = x[:, :3] if dim9 else x
    ref_idx, gaps, vals = do.knn_with_gaps(xs, k)
    from gadm_b200 import ops
    # tc: shapes the tensor-core kernel supports (C % 64 == 0, k <= 20, N % 4 == 0, N >= 256) run there, the others
    # -- and everything when it is switched off -- on the fp32 SIMT kernel; same gates for both
    lib = ops._lib.load()
    on_tc = tc and lib.gadm_knn_feat_tc_workspace_bytes(B, C, N, 3 if dim9 else C, k) > 0
    if tc and not on_tc and C % 64 == 0 and k <= 20 and N >= 256 and N % 4 == 0:
        pytest.fail("the tensor-core path should have taken this shape")
    ops._KNN_FEAT_TC = tc
    try:
        idx = ops.knn_feat(x.to(cuda).contiguous(), k, 3 if dim9 else C).cpu()
    finally:
        ops._KNN_FEAT_TC = True
    thresh = 2e-4 * max(1.0, float(vals.abs().max()) / 79.0)
    ok = gaps > thresh
    assert ok.float().mean() > 0.5      # low-dimensional inputs have many near-ties; they are excluded, not failed
    assert torch.equal(idx[ok], ref_idx[ok])
    assert torch.all(idx[..., 0] == torch.arange(N)[None]), "self is always rank 0"
    # rows below the gap threshold must still be the same SET up to the near-tied members
    same_set = (torch.sort(idx, -1).values == torch.sort(ref_idx, -1).values).all(-1)
    assert same_set.float().mean() > 0.95


@pytest.mark.parametrize("tc", [True, False])
def test_knn_feat_exact_ties_go_to_the_smaller_index(cuda, tc):
    """Every point twice (bit-identical copies at scattered positions): a copy scores exactly like its original in
    every row, so the two must come out next to each other, smaller index first -- in the per-thread lists, across the
    column slices of a row and across tiles (tensor-core kernel), and in the warp lists of the SIMT kernel."""
    from gadm_b200 import ops
    g = torch.Generator().manual_seed(77)
    B, C, N, k = 2, 64, 1024, 20
    half = torch.randn((B, C, N // 2), generator=g)
    perm = torch.randperm(N, generator=g)
    x = torch.empty((B, C, N))
    x[:, :, perm[: N // 2]] = half
    x[:, :, perm[N // 2:]] = half                       # point perm[i] == point perm[i + N/2]
    twin = torch.empty(N, dtype=torch.long)
    twin[perm[: N // 2]] = perm[N // 2:]
    twin[perm[N // 2:]] = perm[: N // 2]
    ops._KNN_FEAT_TC = tc
    try:
        idx = ops.knn_feat(x.to(cuda).contiguous(), k, C).cpu()
    finally:
        ops._KNN_FEAT_TC = True
    # ranks come in (original, copy) pairs: 0/1, 2/3, ... hold twins, smaller index first
    a, b = idx[..., 0::2], idx[..., 1::2]
    assert torch.equal(twin[a], b), "a copy must follow its original"
    assert torch.all(a < b), "ties resolve to the smaller index"
    n = torch.arange(N)[None]
    assert torch.equal(idx[..., 0], torch.minimum(n, twin[n]).expand(B, N))   # the row's own pair leads


def test_get_graph_feature_vs_oracle(cuda):
    from gadm_b200 import dgcnn
    g = torch.Generator().manual_seed(1)
    B, C, N, k = 2, 64, 512, 20
    x = torch.randn((B, C, N), generator=g)
    idx = do.knn(x, k)
    out = dgcnn.get_graph_feature(x.to(cuda), k=k, idx=idx.to(cuda)).cpu()
    ref = do.get_graph_feature(x, k=k, idx=idx)
    assert out.shape == (B, 2 * C, N, k)
    assert torch.equal(out, ref), "gather + subtract is exact in fp32"
    # idx=None path (kNN inside), dim9 path
    x9 = torch.randn((1, 9, 300), generator=g)
    out9 = dgcnn.get_graph_feature(x9.to(cuda), k=16, dim9=True).cpu()
    assert out9.shape == (1, 18, 300, 16)
    assert torch.equal(out9[:, 9:, :, 0], x9)


@pytest.mark.parametrize("B,N,k", [(2, 4096, 20), (1, 777, 16), (3, 300, 5)])
def test_dim9_knn_runs_the_3d_search_and_meets_the_feature_gates(cuda, B, N, k):
    """get_graph_feature(dim9=True) ranks by the first three channels (models/dgcnn.py:38): that search runs on the
    exact 3-D grid kNN.  Same gates as the feature-space kernel against the reference's fp32 distance form."""
    from gadm_b200 import dgcnn
    g = torch.Generator().manual_seed(900 + N)
    x = torch.randn((B, 9, N), generator=g)
    ref_idx, gaps, vals = do.knn_with_gaps(x[:, :3], k)
    idx = dgcnn.knn_xyz(x.to(cuda), k).cpu()
    assert idx.dtype == torch.int64 and idx.shape == (B, N, k)
    ok = gaps > 2e-4 * max(1.0, float(vals.abs().max()) / 79.0)
    assert ok.float().mean() > 0.5
    assert torch.equal(idx[ok], ref_idx[ok])
    assert torch.all(idx[..., 0] == torch.arange(N)[None]), "self is always rank 0"
    same_set = (torch.sort(idx, -1).values == torch.sort(ref_idx, -1).values).all(-1)
    assert same_set.float().mean() > 0.95


def test_pointops_knnquery_and_grouping(cuda):
    from gadm_b200 import pointops
    g = torch.Generator().manual_seed(2)
    b, n, m, c, ns = 2, 600, 200, 16, 8
    xyz, new_xyz = torch.rand((b, n, 3), generator=g), torch.rand((b, m, 3), generator=g)
    idx = pointops.knnquery(ns, xyz.to(cuda), new_xyz.to(cuda))
    assert idx.dtype == torch.int32 and idx.shape == (b, m, ns)
    ref_idx, _ = po.knnquery_naive(ns, xyz, new_xyz)
    # the naive oracle sums (dx^2 + dy^2) + dz^2 like the kernel; ties are measure-zero on random input
    assert torch.equal(idx.cpu(), ref_idx)
    assert torch.equal(pointops.knnquery(ns, xyz.to(cuda)).cpu()[:, :, 0], torch.arange(n).int()[None].expand(b, n))

    feats = torch.randn((b, c, n), generator=g)
    f = feats.to(cuda).requires_grad_(True)
    out = pointops.grouping(f, idx)
    assert torch.equal(out.detach().cpu(), po.grouping(feats, idx.cpu()))
    go = torch.randn((b, c, m, ns), generator=g)
    out.backward(go.to(cuda))
    ref_g = po.grouping_backward(go, idx.cpu(), n)
    assert torch.allclose(f.grad.cpu(), ref_g, rtol=1e-5, atol=1e-5)

    qg = pointops.QueryAndGroup(nsample=ns)
    grouped, gxyz = qg(xyz.to(cuda), new_xyz.to(cuda), feats.to(cuda))
    assert grouped.shape == (b, 3 + c, m, ns) and gxyz.shape == (b, 3, m, ns)
    ref_new, ref_gxyz, _ = po.query_and_group(ns, xyz, new_xyz, feats)
    assert torch.equal(grouped.cpu(), ref_new) and torch.equal(gxyz.cpu(), ref_gxyz)


def test_pointops_vs_reference_golden(cuda):
    """knnquery / knnquery_heap / grouping / QueryAndGroup against pointops_golden.npz: outputs of the reference's own
    KNNQueryNaive.forward (pointops.py:396-426) and QueryAndGroup.forward (:548-585), executed from its source text."""
    import os
    from gadm_b200 import pointops
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "pointops_golden.npz"))
    t = lambda k: torch.from_numpy(z[k]).to(cuda)
    k = int(z["k"])
    assert float(z["knn_gap"]) > 0                          # tie-free: the index order is unambiguous
    assert torch.equal(pointops.knnquery(k, t("xyz"), t("new_xyz")), t("knn_idx"))
    assert torch.equal(pointops.knnquery_heap(k, t("xyz")), t("knn_self_idx"))
    assert torch.equal(pointops.grouping(t("features"), t("knn_idx")), t("qg_new_features")[:, 3:])
    nf, gx, idx = pointops.QueryAndGroup(nsample=k, return_idx=True)(t("xyz"), t("new_xyz"), t("features"))
    assert idx.dtype == torch.int64 and torch.equal(idx, t("qg_idx"))
    assert torch.equal(gx, t("qg_grouped_xyz")) and torch.equal(nf, t("qg_new_features"))
    assert torch.equal(pointops.QueryAndGroup(nsample=k, use_xyz=False)(t("xyz"), t("new_xyz"), t("features"))[0],
                       t("qg_features_only"))
    assert torch.equal(pointops.QueryAndGroup(nsample=k)(t("xyz"), t("new_xyz"))[0], t("qg_xyz_only"))
    with pytest.raises(NotImplementedError):
        pointops.QueryAndGroup(radius=0.1)


def test_gather_neighbour(cuda):
    """models/RandLA/RandLANet.py:729-738."""
    from gadm_b200 import pointops
    g = torch.Generator().manual_seed(4)
    B, N, C, K = 2, 500, 24, 16
    pc = torch.randn((B, N, C), generator=g)
    idx = torch.randint(0, N, (B, N, K), generator=g)
    out = pointops.gather_neighbour(pc.to(cuda), idx.to(cuda)).cpu()
    ref = torch.gather(pc, 1, idx.reshape(B, -1).unsqueeze(-1).repeat(1, 1, C)).reshape(B, N, K, C)
    assert torch.equal(out, ref)


def test_randla_consumers_vs_golden_and_oracle(cuda):
    """random_sample / nearest_interpolation / relative_pos_encoding (RandLANet.py:90-120, 720-727)
    against the fixture produced by executing the reference's own function bodies, and against the oracle at a
    larger ragged shape.  Gathers and maxima are exact; the distance is sqrt of an fp32 sum of squares whose
    reduction order torch does not pin, so it gets 2 ulp."""
    import os
    from gadm_b200 import randla
    from oracle import randla_oracle as ro
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "randla_golden.npz"))
    t = lambda k: torch.from_numpy(z[k])
    out = randla.random_sample(t("feature").to(cuda), t("pool_idx").to(cuda)).cpu()
    assert torch.equal(out, t("random_sample"))
    out = randla.nearest_interpolation(t("feature").to(cuda), t("interp_idx").to(cuda)).cpu()
    assert torch.equal(out, t("nearest_interpolation"))
    rpe = randla.relative_pos_encoding(t("xyz").to(cuda), t("neigh_idx").to(cuda)).cpu()
    ref = t("relative_pos_encoding")
    assert torch.equal(rpe[..., 1:], ref[..., 1:])
    assert torch.allclose(rpe[..., 0], ref[..., 0], rtol=3e-7, atol=0)

    g = torch.Generator().manual_seed(9)
    B, C, N, M, K = 3, 37, 1000, 333, 16
    feat = torch.randn((B, C, N, 1), generator=g)
    pool = torch.randint(0, N, (B, M, K), generator=g)
    assert torch.equal(randla.random_sample(feat.to(cuda), pool.to(cuda)).cpu(), ro.random_sample(feat, pool))
    up = torch.randint(0, N, (B, 2500, 1), generator=g)
    assert torch.equal(randla.nearest_interpolation(feat.to(cuda), up.to(cuda)).cpu(),
                       ro.nearest_interpolation(feat, up))
    pool20 = torch.randint(0, N, (B, M, 20), generator=g)          # k > 16 path
    assert torch.equal(randla.random_sample(feat.to(cuda), pool20.to(cuda)).cpu(), ro.random_sample(feat, pool20))
    xyz = torch.rand((B, N, 3), generator=g)
    nei = torch.randint(0, N, (B, N, K), generator=g)
    got, want = randla.relative_pos_encoding(xyz.to(cuda), nei.to(cuda)).cpu(), ro.relative_pos_encoding(xyz, nei)
    # (the CPU's own summation order of x^2 + y^2 + z^2 depends on the host's vector width: a few ulp)
    assert torch.equal(got[..., 1:], want[..., 1:]) and torch.allclose(got[..., 0], want[..., 0], rtol=1e-6, atol=0)


def test_gather_ops_gradients_vs_autograd_through_the_oracles(cuda):
    """The gathers that sit inside the trained networks are differentiable in the reference (torch.gather / max / cat:
    models/dgcnn.py:41-54, RandLANet.py:90-120, :720-738).  Gradients of the fused ops against torch autograd through the
    CPU oracles, incl. repeated indices (scatter-ADD) and ties of the max (first neighbour wins, as torch's CPU max)."""
    from gadm_b200 import dgcnn, randla
    from oracle import randla_oracle as ro
    g = torch.Generator().manual_seed(17)
    B, C, N, k = 2, 12, 300, 16
    x = torch.randn((B, C, N), generator=g)
    idx = torch.randint(0, N, (B, N, k), generator=g)
    idx[:, :, 3] = idx[:, :, 2]                                   # repeated neighbours
    go = torch.randn((B, 2 * C, N, k), generator=g)
    xg = x.to(cuda).requires_grad_(True)
    dgcnn.get_graph_feature(xg, k=k, idx=idx.to(cuda)).backward(go.to(cuda))
    xr = x.clone().requires_grad_(True)
    do.get_graph_feature(xr, k=k, idx=idx).backward(go)
    assert torch.allclose(xg.grad.cpu(), xr.grad, rtol=1e-4, atol=1e-4)

    M, K = 70, 16
    feat = torch.randn((B, C, N, 1), generator=g)
    feat[:, :, 5] = feat[:, :, 9]                                 # exact ties between two candidates of a maximum
    pool = torch.randint(0, N, (B, M, K), generator=g)
    pool[:, :, 0], pool[:, :, 1] = 9, 5
    go2 = torch.randn((B, C, M, 1), generator=g)
    fg = feat.to(cuda).requires_grad_(True)
    randla.random_sample(fg, pool.to(cuda)).backward(go2.to(cuda))
    fr = feat.clone().requires_grad_(True)
    ro.random_sample(fr, pool).backward(go2)
    assert torch.allclose(fg.grad.cpu(), fr.grad, rtol=1e-5, atol=1e-5)
    interp = torch.randint(0, N, (B, 3 * M, 1), generator=g)
    go3 = torch.randn((B, C, 3 * M, 1), generator=g)
    fg2 = feat.to(cuda).requires_grad_(True)
    randla.nearest_interpolation(fg2, interp.to(cuda)).backward(go3.to(cuda))
    fr2 = feat.clone().requires_grad_(True)
    ro.nearest_interpolation(fr2, interp).backward(go3)
    assert torch.allclose(fg2.grad.cpu(), fr2.grad, rtol=1e-5, atol=1e-5)

    pc = torch.randn((B, N, 7), generator=g)
    nidx = torch.randint(0, N, (B, N, K), generator=g)            # the reference's version needs one row per point
    go4 = torch.randn((B, N, K, 7), generator=g)
    pg = pc.to(cuda).requires_grad_(True)
    randla.gather_neighbour(pg, nidx.to(cuda)).backward(go4.to(cuda))
    pr = pc.clone().requires_grad_(True)
    ro.gather_neighbour(pr, nidx).backward(go4)
    assert torch.allclose(pg.grad.cpu(), pr.grad, rtol=1e-5, atol=1e-5)

    xyz = torch.rand((B, N, 3), generator=g)
    nn_idx = torch.randint(0, N, (B, N, K), generator=g)
    nn_idx[:, :, 0] = (torch.arange(N) + 1) % N                   # never the point itself in slot 0 ...
    same = nn_idx == torch.arange(N)[None, :, None]
    nn_idx[same] = (nn_idx[same] + 1) % N                         # ... nor anywhere (|p - q| is not differentiable at 0)
    go5 = torch.randn((B, N, K, 10), generator=g)
    zg = xyz.to(cuda).requires_grad_(True)
    randla.relative_pos_encoding(zg, nn_idx.to(cuda)).backward(go5.to(cuda))
    zr = xyz.clone().requires_grad_(True)
    ro.relative_pos_encoding(zr, nn_idx).backward(go5)
    assert torch.allclose(zg.grad.cpu(), zr.grad, rtol=1e-4, atol=1e-4)
