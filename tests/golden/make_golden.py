"""Generates the golden fixtures in this directory FROM THE REFERENCE ITSELF.  Run in the build container only
(needs /root/reference; the GPU box never runs this):   python tests/golden/make_golden.py

  knn_golden.npz     outputs of the reference's own compiled nanoflann kNN (oracle/_ref, built from
                     /root/reference/models/RandLA/utils/nearest_neighbors/knn_.cxx) on small seeded clouds
  match_golden.npz   outputs of the reference's matcher lines, EXECUTED from the reference source text
                     (evaluator.py:89-93 and utils/pvn3d_eval_utils_kpls.py:437-441) on seeded descriptors
  dgcnn_golden.npz   outputs of models/dgcnn.py knn() / get_graph_feature() imported from /root/reference
  randla_golden.npz  outputs of RandLANet.py random_sample / nearest_interpolation / relative_pos_encoding /
                     gather_neighbour, EXECUTED from the reference source text
  circle_golden.npz  the training-side matching loss: CircleLoss imported from models/loss.py, GeoMatch.matching_loss /
                     pointwise_feature_matching and pdist EXECUTED from the reference source text
  pointops_golden.npz  lib/pointops/functions/pointops.py: KNNQueryNaive.forward (:396-426, pure torch, the reference's
                     own oracle for its CUDA knnquery) and QueryAndGroup.forward (:548-585) EXECUTED from the reference
                     source text (its `grouping` CUDA call replaced by the gather its docstring :151-155 specifies)
  heads_golden.json  state_dict keys / shapes of the GeoMatch head stacks built with the reference's models/pytorch_utils.py
Nothing from the reference is copied into the repo: only inputs and numeric outputs are stored."""
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import knn_oracle as ko  # noqa: E402


def ref_lines(path, lo, hi):
    with open(os.path.join(REF, path)) as f:
        lines = f.readlines()[lo - 1: hi]
    return textwrap.dedent("".join(lines))


def make_knn():
    rng = np.random.default_rng(123)
    out = {}
    cases = {
        "uniform": (rng.random((2, 300, 3), dtype=np.float32), rng.random((2, 120, 3), dtype=np.float32), 16),
        "self": (None, None, 16),
        "one_nn": (rng.random((1, 257, 3), dtype=np.float32), rng.random((1, 333, 3), dtype=np.float32), 1),
    }
    s = rng.random((1, 400, 3), dtype=np.float32)
    cases["self"] = (s, s.copy(), 16)
    u = rng.random((1, 150, 3), dtype=np.float32)           # wrap-padded duplicates (linemod_pbr.py:492)
    d = np.concatenate([u, u[:, :50]], axis=1)
    cases["dups"] = (d, d.copy(), 8)
    for name, (sup, qry, k) in cases.items():
        ref = ko.knn_reference(sup, qry, k, omp=True)       # the reference's own code
        assert np.array_equal(ref, ko.knn_reference(sup, qry, k, omp=False))
        out[name + "_support"], out[name + "_query"] = sup, qry
        out[name + "_k"] = np.int64(k)
        out[name + "_ref_idx"] = ref
        out[name + "_ref_d2"] = ko.dist2_of(sup, qry, ref)
    np.savez_compressed(os.path.join(HERE, "knn_golden.npz"), **out)


def make_match():
    g = torch.Generator().manual_seed(7)
    d, N, M = 64, 96, 64
    bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
    rgbd = bf(torch.randn((d, N), generator=g))
    mesh = bf(torch.randn((d, M), generator=g))
    rgbd[:, :8] = -rgbd[:, :8].abs()                        # rows on which the -1 pad column wins
    seg = torch.randn((2, N), generator=g)
    out = {"rgbd": rgbd.numpy(), "mesh": mesh.numpy(), "seg": seg.numpy()}

    # evaluator.py:78-93 executed from the reference text (cuda() calls are not on these lines)
    ns = {"torch": torch, "F": F, "seg_features": seg, "rgbd_features": rgbd.clone(), "mesh_features": mesh.clone()}
    exec(ref_lines("evaluator.py", 78, 79), ns)             # seg argmax, transpose
    exec(ref_lines("evaluator.py", 82, 82), ns)             # cls_msk
    exec(ref_lines("evaluator.py", 88, 93), ns)             # select, normalize x2, matmul, torch.max
    out["live_mask"] = ns["cls_msk"].numpy()
    out["live_idx"] = ns["obj_pts_idx"].numpy()
    out["live_max"] = ns["max_th"].numpy()
    out["live_sim"] = ns["obj_pts_sim"].numpy()

    # padded variant, utils/pvn3d_eval_utils_kpls.py:437-441 (pad -1 column, normalise, matmul, argmax)
    ns2 = {"torch": torch, "F": F, "selected_rgbd_feature": rgbd.t().contiguous(), "mesh_features": mesh.clone(),
           "padding": -torch.ones((d, 1), dtype=torch.float32)}
    exec(ref_lines("utils/pvn3d_eval_utils_kpls.py", 437, 441), ns2)
    out["pad_idx"] = ns2["obj_pts_idx"].numpy()
    out["pad_sim"] = ns2["obj_pts_sim"].numpy()
    np.savez_compressed(os.path.join(HERE, "match_golden.npz"), **out)


def make_dgcnn():
    sys.path.insert(0, REF)
    import models.dgcnn as ref_dgcnn                        # imports cleanly (torch only)

    class _TorchOnCpu(types.ModuleType):                    # models/dgcnn.py:39 hard-codes torch.device('cuda')
        def __getattr__(self, name):
            if name == "device":
                return lambda *_a, **_k: torch.device("cpu")
            return getattr(torch, name)
    ref_dgcnn.torch = _TorchOnCpu("torch_cpu_shim")

    g = torch.Generator().manual_seed(11)
    x = torch.randn((2, 16, 200), generator=g)
    x9 = torch.randn((1, 9, 150), generator=g)
    out = {"x": x.numpy(), "x9": x9.numpy()}
    out["knn_idx"] = ref_dgcnn.knn(x, 20).numpy()
    out["graph"] = ref_dgcnn.get_graph_feature(x, k=20).numpy()
    out["graph9"] = ref_dgcnn.get_graph_feature(x9, k=16, dim9=True).numpy()
    np.savez_compressed(os.path.join(HERE, "dgcnn_golden.npz"), **out)


def make_randla():
    """models/RandLA/RandLANet.py is not importable here (it imports the compiled nearest_neighbors extension), so
    the four functions are EXECUTED from the reference's source text: random_sample :90-105,
    nearest_interpolation :108-120, relative_pos_encoding :720-727, gather_neighbour :730-738."""
    ns = {"torch": torch}
    exec(ref_lines("models/RandLA/RandLANet.py", 90, 105), ns)
    exec(ref_lines("models/RandLA/RandLANet.py", 108, 120), ns)
    exec(ref_lines("models/RandLA/RandLANet.py", 720, 727), ns)
    exec(ref_lines("models/RandLA/RandLANet.py", 730, 738), ns)

    class _Block:                                            # relative_pos_encoding calls self.gather_neighbour
        gather_neighbour = staticmethod(ns["gather_neighbour"])
    g = torch.Generator().manual_seed(21)
    B, C, N, M, K = 2, 12, 160, 40, 16
    feature = torch.randn((B, C, N, 1), generator=g)
    pool_idx = torch.randint(0, N, (B, M, K), generator=g)
    interp_idx = torch.randint(0, N, (B, 3 * M, 1), generator=g)
    xyz = torch.rand((B, N, 3), generator=g)
    neigh_idx = torch.randint(0, N, (B, N, K), generator=g)
    out = {"feature": feature.numpy(), "pool_idx": pool_idx.numpy(), "interp_idx": interp_idx.numpy(),
           "xyz": xyz.numpy(), "neigh_idx": neigh_idx.numpy()}
    out["random_sample"] = ns["random_sample"](feature, pool_idx).numpy()
    out["nearest_interpolation"] = ns["nearest_interpolation"](feature, interp_idx).numpy()
    out["relative_pos_encoding"] = ns["relative_pos_encoding"](_Block(), xyz, neigh_idx).numpy()
    out["gather_neighbour"] = ns["gather_neighbour"](xyz, neigh_idx).numpy()
    np.savez_compressed(os.path.join(HERE, "randla_golden.npz"), **out)


def make_circle():
    """The training-side matching loss, EXECUTED from the reference: CircleLoss is imported from models/loss.py (torch
    only), GeoMatch.matching_loss (models/geoMatch.py:55-83), GeoMatch.pointwise_feature_matching (:102-157) and pdist
    (utils/basic_utils.py:86-93) run from the source text; .cuda() is a no-op while they run."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_loss", os.path.join(REF, "models", "loss.py"))
    ref_loss = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_loss)
    ns = {"torch": torch, "F": F}
    exec(ref_lines("utils/basic_utils.py", 86, 93), ns)                   # pdist
    exec(ref_lines("models/geoMatch.py", 55, 83), ns)                     # matching_loss(self, ...)
    exec(ref_lines("models/geoMatch.py", 102, 157), ns)                   # pointwise_feature_matching(self, ...)

    g = torch.Generator().manual_seed(31)
    B, d, N, M = 3, 64, 96, 64
    bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
    mesh = bf(torch.randn((1, d, M), generator=g))
    i = torch.arange(M, dtype=torch.float64) + 0.5                         # Fibonacci sphere, diameter 0.2 m
    phi, theta = torch.acos(1 - 2 * i / M), np.pi * (1 + 5 ** 0.5) * i
    xyz = (0.1 * torch.stack([torch.cos(theta) * torch.sin(phi), torch.sin(theta) * torch.sin(phi),
                              torch.cos(phi)], dim=1)).float()
    vis = (torch.rand((B, M), generator=g) < 0.6)
    labels = (torch.rand((B, N), generator=g) < 0.45).long()
    labels[2] = 0
    labels[2, :2] = 1                                                      # < 3 foreground rows: the sample is skipped
    match_idx = torch.full((B, N), M, dtype=torch.int32)
    rgbd = bf(torch.randn((B, d, N), generator=g))
    for b in range(B):
        vis_ids = torch.where(vis[b])[0]
        pick = vis_ids[torch.randint(0, len(vis_ids), (N,), generator=g)]
        on = torch.rand((N,), generator=g) < 0.8                           # 20 % of the rows are off the model
        match_idx[b] = torch.where(on, pick, torch.full_like(pick, M)).int()
        near = on.nonzero()[:, 0][::2]                                     # half of the on-model rows look like their vertex
        rgbd[b][:, near] = bf(mesh[0][:, match_idx[b][near].long()] + 0.5 * torch.randn((d, len(near)), generator=g))
    positive_r = 0.045

    class _Emb:
        sys_corr_idx = None
        _buffers = {"xyz": xyz}
    stub = types.SimpleNamespace(feat_dim=d, positive_r=positive_r, circle_loss=ref_loss.CircleLoss(16), model_emb=_Emb())
    stub.matching_loss = types.MethodType(ns["matching_loss"], stub)
    x = {"labels": labels, "match_idx": match_idx, "RT": torch.zeros((B, 3, 4)), "visible_flag": vis.to(torch.uint8)}
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        total = ns["pointwise_feature_matching"](stub, rgbd.clone(), mesh.clone(), x)
        per_sample = []
        for b in range(2):
            xb = {k: v[b:b + 1] for k, v in x.items()}
            per_sample.append(float(ns["pointwise_feature_matching"](stub, rgbd[b:b + 1].clone(), mesh.clone(), xb)))
    finally:
        torch.Tensor.cuda = real_cuda
    out = {"rgbd": rgbd.numpy(), "mesh": mesh.numpy(), "xyz": xyz.numpy(), "vis": vis.numpy().astype(np.uint8),
           "labels": labels.numpy(), "match_idx": match_idx.numpy(), "positive_r": np.float32(positive_r),
           "ref_total": np.float32(float(total)), "ref_per_sample": np.asarray(per_sample, dtype=np.float32)}

    # the DGCNN variant (models/geoMatch_DGCNN.py:53-78 matching_loss with a per-vertex radius positive_r / 1000 * depth,
    # :80-136 pointwise_feature_matching with the e0 pad column and x['origin_labels']), same inputs + poses
    ns2 = {"torch": torch, "F": F, "pdist": ns["pdist"]}
    exec(ref_lines("models/geoMatch_DGCNN.py", 53, 78), ns2)
    exec(ref_lines("models/geoMatch_DGCNN.py", 80, 136), ns2)
    RT = torch.zeros((B, 3, 4))
    for b in range(B):
        q, _ = torch.linalg.qr(torch.randn((3, 3), generator=g))
        RT[b, :, :3] = q * torch.sign(torch.det(q))
        RT[b, :, 3] = torch.tensor([0.05 * b, -0.02, 0.8 + 0.3 * b])
    pos_r_dgcnn = 40.0                                                   # radius = 0.04 * depth (0.8 .. 1.5 m)

    class _Emb2:
        _buffers = {"mesh": torch.cat([xyz.t(), torch.zeros((3, M))], dim=0)[None]}     # [1, 6, M], xyz in channels 0..2
    stub2 = types.SimpleNamespace(feat_dim=d, positive_r=pos_r_dgcnn, circle_loss=ref_loss.CircleLoss(16), model_emb=_Emb2())
    stub2.matching_loss = types.MethodType(ns2["matching_loss"], stub2)
    x2 = {"origin_labels": labels, "match_idx": match_idx, "RT": RT, "visible_flag": vis.to(torch.uint8)}
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        total2 = ns2["pointwise_feature_matching"](stub2, rgbd.clone(), mesh.clone(), x2)
    finally:
        torch.Tensor.cuda = real_cuda
    out.update(RT=RT.numpy(), dgcnn_positive_r=np.float32(pos_r_dgcnn), dgcnn_ref_total=np.float32(float(total2)))

    # the symmetry-aware branch (models/geoMatch.py:138-141 -> matching_loss_sys :86-100): model_emb.sys_corr_idx set,
    # model_emb.sys_idx indexed by the selected scene-point ids exactly as the reference does
    exec(ref_lines("models/geoMatch.py", 86, 100), ns)
    sys_idx = torch.randint(0, N, (N,), generator=g)

    class _Emb3:
        sys_corr_idx = True
        _buffers = {"xyz": xyz}
    _Emb3.sys_idx = sys_idx
    stub3 = types.SimpleNamespace(feat_dim=d, positive_r=positive_r, circle_loss=ref_loss.CircleLoss(16), model_emb=_Emb3())
    stub3.matching_loss_sys = types.MethodType(ns["matching_loss_sys"], stub3)
    real_tensor = torch.tensor
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        total3 = ns["pointwise_feature_matching"](stub3, rgbd.clone(), mesh.clone(), x)
    finally:
        torch.Tensor.cuda = real_cuda
    out.update(sys_idx=sys_idx.numpy(), sys_ref_total=np.float32(float(total3)))
    print("circle golden (matching_loss_sys): total", float(total3))
    print("circle golden (DGCNN variant): total", float(total2))
    np.savez_compressed(os.path.join(HERE, "circle_golden.npz"), **out)
    print("circle golden: total", float(total), "per sample", per_sample)


def make_heads():
    """The head stacks of GeoMatch.__init__ (models/geoMatch.py:33-51, geoMatch_DGCNN.py:30-48) built with the reference's
    own models/pytorch_utils.py: state_dict keys and shapes -> heads_golden.json (what a checkpoint of the reference
    holds for the heads)."""
    import importlib.util, json
    spec = importlib.util.spec_from_file_location("ref_pt_utils", os.path.join(REF, "models/pytorch_utils.py"))
    pt_utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pt_utils)
    out = {}
    for variant, enc_in in (("geoMatch", 128), ("geoMatch_DGCNN", 128)):
        feat_dim = 128
        seg = (pt_utils.Seq(feat_dim).conv1d(128, bn=True).conv1d(128, bn=True).conv1d(128, bn=True)
               .conv1d(2, activation=None))
        enc = (pt_utils.Seq(enc_in).conv1d(128, bn=True).conv1d(128, bn=True).conv1d(128, bn=True)
               .conv1d(feat_dim, activation=None, bias=False))
        nrm = pt_utils.Conv1d(feat_dim, feat_dim, bn=True)
        keys = {}
        for name, mod in (("seg_layer", seg), ("feature_encoding_layer", enc), ("normalize_feature_layer", nrm)):
            for k, v in mod.state_dict().items():
                keys[name + "." + k] = list(v.shape)
        out[variant] = keys
    with open(os.path.join(HERE, "heads_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("heads golden:", len(out["geoMatch"]), "state_dict entries per variant")


def make_pointops():
    """lib/pointops ships no CUDA sources and nothing imports it, but KNNQueryNaive.forward (:396-426) is pure torch:
    it is EXECUTED from the reference text.  QueryAndGroup.forward (:548-585) is executed too, with knnquery_heap bound
    to that naive query and `grouping` to the gather its docstring (:151-155) specifies
    (out[b, c, m, s] = features[b, c, idx[b, m, s]])."""
    from typing import Tuple
    ns = {"torch": torch, "Tuple": Tuple}
    exec(ref_lines("lib/pointops/functions/pointops.py", 396, 426), ns)          # @staticmethod def forward(ctx, ...)
    knn_naive = ns["forward"]

    def grouping(features, idx):
        b, c, n = features.shape
        _, m, k = idx.shape
        return torch.gather(features, 2, idx.long().view(b, 1, m * k).expand(b, c, m * k)).view(b, c, m, k)
    ns2 = {"torch": torch, "grouping": grouping, "knnquery_heap": lambda k, xyz, new: knn_naive(None, k, xyz, new),
           "ballquery": None}
    exec(ref_lines("lib/pointops/functions/pointops.py", 548, 585), ns2)         # def forward(self, xyz, new_xyz, ...)

    class _QG:
        radius, nsample, use_xyz, return_idx = None, 16, True, True
    g = torch.Generator().manual_seed(77)
    b, n, m, c, k = 2, 300, 90, 5, 16
    xyz = torch.rand((b, n, 3), generator=g)
    new_xyz = torch.rand((b, m, 3), generator=g)
    feats = torch.randn((b, c, n), generator=g)
    out = {"xyz": xyz.numpy(), "new_xyz": new_xyz.numpy(), "features": feats.numpy(), "k": np.int64(k)}
    out["knn_idx"] = knn_naive(None, k, xyz, new_xyz).numpy()
    out["knn_self_idx"] = knn_naive(None, k, xyz, None).numpy()
    nf, gxyz, idx = ns2["forward"](_QG(), xyz, new_xyz, feats)
    out["qg_new_features"], out["qg_grouped_xyz"], out["qg_idx"] = nf.numpy(), gxyz.numpy(), idx.numpy()
    _QG.use_xyz = False
    out["qg_features_only"] = ns2["forward"](_QG(), xyz, new_xyz, feats)[0].numpy()
    _QG.use_xyz = True
    out["qg_xyz_only"] = ns2["forward"](_QG(), xyz, new_xyz, None)[0].numpy()
    # the same distances in the order the naive query computes them (for the tie-free check of the fixture)
    dist = (new_xyz[:, :, None, :] - xyz[:, None, :, :]).pow(2).sum(dim=3)
    srt = torch.sort(dist, dim=2)[0]
    out["knn_gap"] = (srt[:, :, 1:k + 1] - srt[:, :, :k]).min().numpy()          # > 0: no ties among the first k + 1
    np.savez_compressed(os.path.join(HERE, "pointops_golden.npz"), **out)
    print("pointops golden: min gap among the first k + 1 distances", float(out["knn_gap"]))


if __name__ == "__main__":
    assert os.path.isdir(REF), "needs the reference tree"
    assert ko.have_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "randla":        # add one fixture without regenerating the others
        make_randla()
    elif len(sys.argv) > 1 and sys.argv[1] == "circle":
        make_circle()
    elif len(sys.argv) > 1 and sys.argv[1] == "pointops":
        make_pointops()
    elif len(sys.argv) > 1 and sys.argv[1] == "heads":
        make_heads()
    else:
        make_knn(); make_match(); make_dgcnn(); make_randla(); make_circle(); make_pointops(); make_heads()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
