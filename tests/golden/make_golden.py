"""Generates the golden fixtures in this directory FROM THE REFERENCE ITSELF.  Run in the build container only
(needs /root/reference; the GPU box never runs this):   python tests/golden/make_golden.py

  knn_golden.npz     outputs of the reference's own compiled nanoflann kNN (oracle/_ref, built from
                     /root/reference/models/RandLA/utils/nearest_neighbors/knn_.cxx) on small seeded clouds
  match_golden.npz   outputs of the reference's matcher lines, EXECUTED from the reference source text
                     (evaluator.py:89-93 and utils/pvn3d_eval_utils_kpls.py:437-441) on seeded descriptors
  dgcnn_golden.npz   outputs of models/dgcnn.py knn() / get_graph_feature() imported from /root/reference
  randla_golden.npz  outputs of RandLANet.py random_sample / nearest_interpolation / relative_pos_encoding /
                     gather_neighbour, EXECUTED from the reference source text
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


if __name__ == "__main__":
    assert os.path.isdir(REF), "needs the reference tree"
    assert ko.have_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "randla":        # add one fixture without regenerating the others
        make_randla()
    else:
        make_knn(); make_match(); make_dgcnn(); make_randla()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
