"""CPU (-m "not gpu"): the C-ABI library loads and exports everything include/gadm.h declares (no compute calls
without a GPU), host-side logic (job tables, sharding over gloo, pose fit, generators), and the rule that the
product package never touches oracle/."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "geometric-aware-dense-matching_b200")


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gadm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gadm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    import gadm_b200
    from gadm_b200 import _lib
    lib = gadm_b200.load_library()
    syms = header_symbols()
    assert len(syms) == 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gadm.h but not exported by libgadm.so"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes signature table must cover the header exactly"
    assert lib.gadm_abi_version() == 3
    assert "#define GADM_ABI_VERSION 3" in open(os.path.join(ROOT, "include", "gadm.h")).read()
    assert ctypes.sizeof(_lib.KnnJob) == 64


def test_error_codes_and_no_gpu_behaviour():
    import gadm_b200
    lib = gadm_b200.load_library()
    assert lib.gadm_strerror(0) == b"ok"
    for code in range(-7, 0):
        assert len(lib.gadm_strerror(code)) > 3
    assert lib.gadm_operand_k(128, 0) == 128 and lib.gadm_operand_k(128, 1) == 384
    assert lib.gadm_operand_k(128, 9) == -2 and lib.gadm_operand_k(0, 0) == -1
    assert lib.gadm_aux_floats(8, 8192) == 8 * 8192 * 7 and lib.gadm_aux_floats(1, 520) == 520 * 7
    if not torch.cuda.is_available():
        # without a device gadm_init fails (CUDA error) and every compute entry point refuses: no CPU fallback
        assert lib.gadm_init(0) in (-5, -6)
        assert lib.gadm_knn_feat(None, 1, 1, 1, 1, 1, None, None) == -7
        assert lib.gadm_match_fwd(*([None] * 7), 1, 1, 8, 64, 1, ctypes.c_float(16.0), 0, 0,
                                  *([None] * 5), 0, None) == -7
        assert lib.gadm_match_workspace_bytes() == 0


def test_knn_workspace_planning_is_host_only():
    import gadm_b200
    from gadm_b200 import ops
    lib = gadm_b200.load_library()
    jobs = ops.make_jobs([(0, 0, 0, 5000, 5000, 5000 * 16, 5000, 5000, 16, 2),
                          (0, 0, 160000, 5000, 5000, 5000, 5000, 5000, 1, 2),     # same cloud: shares the grid
                          (0, 0, 170000, 100, 100, 100, 100, 100, 1, 2)])
    assert lib.gadm_knn3d_workspace_bytes(jobs, 3, 0) == 0                        # BRUTE needs none
    one = lib.gadm_knn3d_workspace_bytes(jobs, 1, 1)
    assert one > 2 * 5000 * 16
    assert lib.gadm_knn3d_workspace_bytes(jobs, 2, 2) == one                      # deduplicated
    bad = ops.make_jobs([(0, 0, 0, 10, 10, 160, 10, 10, 16, 1)])                  # k > n_support
    assert lib.gadm_knn3d_workspace_bytes(bad, 1, 2) == 0


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "oracle/_" not in txt or f.endswith(".md"), f


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from gadm_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GadmError, match="no CPU fallback"):
        _lib.load()


def test_knn_pyramid_job_table_matches_reference_schedule():
    from gadm_b200 import synth
    from gadm_b200.knn import KnnPyramid
    from oracle import knn_oracle as ko
    N, in_size, B = 3200, 64, 3
    cld, sr = synth.depth_cloud(in_size, N, 1)
    calls = ko.schedule(cld, sr)
    pyr = KnnPyramid(N, {s: (in_size // s) ** 2 for s in (2, 4, 8)}, B)
    assert len(pyr.jobs) == 22 and [n[0] for n in pyr.names] == [c[0] for c in calls]
    off = 0
    for job, (name, sup, qry, k) in zip(pyr.jobs, calls):
        assert (job.n_support, job.n_query, job.k, job.batch) == (len(sup), len(qry), k, B), name
        assert job.out_off == off and job.out_bstride == len(qry) * k
        off += B * len(qry) * k
    assert pyr.out_elems == off
    # the packed buffer really holds the clouds where the table says
    cb, srb = synth.frame_batch(B, in_size, N, seed=1)
    pts = pyr.pack(cb, srb).view(B, pyr.P, 3)
    for job, (name, sup, qry, k) in zip(pyr.jobs, calls):
        assert np.array_equal(pts[0, job.support_off: job.support_off + job.n_support].numpy(), sup), name
        assert np.array_equal(pts[0, job.query_off: job.query_off + job.n_query].numpy(), qry), name
    # algorithmic bytes of the schedule at N=12800, in_size=128 ~ 2.87 MB
    big = KnnPyramid(12800, {2: 4096, 4: 1024, 8: 256}, 1)
    assert abs(big.algorithmic_bytes - 2.87e6) < 0.02e6


def test_rt_from_moments_matches_oracle_kabsch():
    from gadm_b200 import matching
    from oracle import match_oracle as mo
    g = torch.Generator().manual_seed(3)
    A, Bp = torch.randn((40, 3), generator=g).double(), torch.randn((40, 3), generator=g).double()
    mom = np.concatenate([[40.0], A.sum(0).numpy(), Bp.sum(0).numpy(), (A.T @ Bp).numpy().reshape(-1)])
    T = matching.rt_from_moments(mom)
    assert np.allclose(T, mo.best_fit_transform(A, Bp).numpy(), atol=1e-9)
    assert matching.sentinel_pose()[2, 3] == -1000 and matching.sentinel_pose().shape == (3, 4)


def test_synth_generators():
    from gadm_b200 import synth
    a = synth.descriptors(2, 100, 64, 64, regime="planted", seed=5)
    b = synth.descriptors(2, 100, 64, 64, regime="planted", seed=5)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert torch.equal(a[0], synth.bf16_round(a[0])), "descriptors must be bf16-representable"
    xyz = synth.fibonacci_sphere(8192, 0.2)
    assert torch.allclose(xyz.norm(dim=1), torch.full((8192,), 0.1), atol=1e-6)
    cld, sr = synth.depth_cloud(128, 12800, 1000, dup_frac=0.1)
    assert cld.shape == (12800, 3) and sr[4].shape == (1024, 3)
    assert len(np.unique(cld, axis=0)) < 12800          # wrap duplicates present
    assert cld[:, 2].min() >= 0.4 and cld[:, 2].max() <= 1.5


def test_frame_sharding():
    from gadm_b200 import sharding
    for n, w in [(256, 8), (10, 4), (3, 8)]:
        rs = [sharding.frame_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1
    bins = sharding.balanced_assignment([6, 1, 1, 1, 1, 1, 1, 6], 2)
    assert sorted(sum(bins, [])) == list(range(8))
    loads = [sum([6, 1, 1, 1, 1, 1, 1, 6][f] for f in b) for b in bins]
    assert abs(loads[0] - loads[1]) <= 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import gadm_b200
from gadm_b200 import sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=world)
n = 7
lo, hi = sharding.frame_range(n, rank, world)
local = (torch.arange(lo, hi).float()[:, None] * torch.ones(1, 5))
full = sharding.gather_outputs(local, n)
assert full.shape == (n, 5) and torch.equal(full[:, 0], torch.arange(n).float()), full
dist.destroy_process_group()
print("ok", rank)
"""


def test_gather_outputs_gloo_world2(tmp_path):
    """N > 1 host path on CPU: world_size 2, gloo, ragged shards (7 frames)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, port], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0 and "ok" in out, out


def test_dgcnn_positive_radius_matches_reference_formula():
    """matching.dgcnn_positive_radius (host logic, torch only) == models/geoMatch_DGCNN.py:64-65 as restated by the
    oracle: positive_r / 1000 * camera-space depth of every model vertex."""
    import numpy as np
    from gadm_b200 import matching
    from oracle import circle_oracle as co
    g = np.load(os.path.join(ROOT, "tests", "golden", "circle_golden.npz"))
    xyz, RT, pr = torch.from_numpy(g["xyz"]), torch.from_numpy(g["RT"]), float(g["dgcnn_positive_r"])
    got = matching.dgcnn_positive_radius(xyz, RT, pr)
    want = torch.stack([co.dgcnn_radius(xyz, RT[b], pr) for b in range(RT.shape[0])])
    assert got.shape == (RT.shape[0], xyz.shape[0]) and torch.allclose(got, want, rtol=1e-6, atol=1e-9)


def test_bf16n_column_scales_stay_below_the_pruning_bound():
    """GADM_MATCH_ARGMAX_BF16N skips a chunk when max(raw, 0) * (1 + 2^-8) cannot beat a running maximum; that is exact
    only if every column scale 1 / ||bf16(normalised column)|| is <= 1 + 2^-8.  Round-to-nearest bf16 has a relative
    error <= 2^-9 per element, so ||.|| >= 1 - 2^-9: checked here on random, sparse, constant and worst-case columns
    (every element just above a rounding midpoint) for d = 64 .. 256, emulating gadm_prep_model's BF16N path."""
    import torch.nn.functional as F
    bound = 1.0 + 2.0 ** -8
    g = torch.Generator().manual_seed(9)
    for d in (64, 128, 192, 256):
        cols = [torch.randn((d, 4096), generator=g), torch.randn((d, 512), generator=g) * (torch.rand((d, 512), generator=g) < 0.05),
                torch.ones((d, 8)), torch.eye(d)[:, :8]]
        # worst case for the norm: every element rounds DOWN by almost half an ulp after the normalisation
        w = torch.full((d, 8), 1.0) + 2.0 ** -8 * 0.999
        cols.append(w * torch.tensor([1.0, 2.0, 0.5, 3.0, 7.0, 0.1, 11.0, 1e-3])[None])
        for m in cols:
            m = m[:, m.norm(dim=0) > 0]
            mt = F.normalize(m.float(), p=2, dim=0).to(torch.bfloat16).float()
            scale = 1.0 / mt.norm(dim=0).clamp_min(1e-12)
            assert float(scale.max()) <= bound, (d, float(scale.max()))


def test_geomatch_cfg_constructor_heads_match_reference_state_dict():
    """GeoMatch(cfg, cls_id) / GeoMatchDGCNN(cfg, cls_id) build heads whose state_dict keys and shapes are those of the
    reference's own pt_utils stacks (tests/golden/heads_golden.json, generated from models/pytorch_utils.py)."""
    import json
    import torch.nn as nn
    from gadm_b200 import matching
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "heads_golden.json")))

    class Emb(nn.Module):
        def forward(self, *a):
            raise AssertionError("not called")
    cfg = {"feat_dim": 128, "neighbor_dis_th": 0.03, "model_d": {5: 200.0}}
    for cls, key in ((matching.GeoMatch, "geoMatch"), (matching.GeoMatchDGCNN, "geoMatch_DGCNN")):
        net = cls(cfg, 5, pcd_emb=Emb(), model_emb=Emb())
        sd = {k: list(v.shape) for k, v in net.state_dict().items()
              if k.split(".")[0] in ("seg_layer", "feature_encoding_layer", "normalize_feature_layer")}
        assert sd == gold[key]
        assert "awl.params" in net.state_dict()                       # AutomaticWeightedLoss(2), loss.py:507-510
    assert abs(matching.GeoMatch(cfg, 5, Emb(), Emb()).positive_r - 0.03 * 200.0 / 1000.0) < 1e-12   # geoMatch.py:24
    assert matching.GeoMatchDGCNN(cfg, 5, Emb(), Emb()).positive_r == 3                               # geoMatch_DGCNN.py:22
    with pytest.raises(NotImplementedError):
        matching.GeoMatch(cfg, 5)                                      # the backbones are outside the package


def test_graft_entry_build_check_passes():
    """__graft_entry__.build() is the driver's "does it build" gate: it must succeed on a CPU-only host (nvcc
    cross-compiles) and its own consistency check (library ABI version == include/gadm.h) must hold."""
    import importlib
    ge = importlib.import_module("__graft_entry__")
    ge.build()
