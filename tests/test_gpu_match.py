"""-m gpu: fused matcher (C ABI) vs oracle/match_oracle.py on the same bf16-representable inputs.

Gates (north star / SURVEY.md 8(c)):
  argmax index   exact on every row whose oracle top-1 margin exceeds 1e-3 (mismatch rate reported below it)
  max_sim        |err| <= 1e-3 (absolute on a cosine in [-1, 1])
  weight         relative error <= 1e-3
  soft_xyz       |err| <= 1e-3 * model diameter
"""
import numpy as np
import pytest
import torch

from oracle import match_oracle as mo

pytestmark = pytest.mark.gpu

TOL = 1e-3


def _check(out, ref, M, diameter, soft=True, rows=None):
    idx, max_sim, weight, soft_xyz = [None if o is None else o.cpu() for o in out]
    sel = slice(None) if rows is None else rows
    idx, max_sim = idx[sel], max_sim[sel]
    decided = ref["margin"] > TOL
    assert decided.float().mean() > 0.5
    assert torch.equal(idx[decided], ref["idx"][decided]), "argmax must be exact where margin > 1e-3"
    assert (max_sim - ref["max_sim"]).abs().max() <= TOL
    # below the margin the kernel may pick the runner-up, but its value must still be the max within tolerance
    if soft:
        weight, soft_xyz = weight[sel], soft_xyz[sel]
        rel = ((weight - ref["weight"]).abs() / ref["weight"]).max()
        assert rel <= TOL, f"weight rel err {rel}"
        err = (soft_xyz - ref["soft_xyz"]).abs().max()
        assert err <= TOL * diameter, f"soft_xyz err {err}"


@pytest.mark.parametrize("regime", ["random", "planted"])
@pytest.mark.parametrize("N,M,d", [(1000, 1024, 128), (777, 520, 64), (300, 8192, 128), (257, 768, 256)])
def test_match_soft_vs_oracle(cuda, regime, N, M, d):
    from gadm_b200 import matching, synth
    rgbd, mesh, _ = synth.descriptors(2, N, M, d, n_obj=1, regime=regime, seed=2000 + N)
    diam = 0.2
    xyz = synth.fibonacci_sphere(M, diam)
    out = matching.match(rgbd.to(cuda), mesh.to(cuda), xyz[None].to(cuda), gamma=16.0, mode="soft")
    torch.cuda.synchronize()
    for b in range(2):
        ref = mo.match_soft(rgbd[b], mesh[0], xyz, gamma=16.0)
        _check([o[b] for o in out], ref, M, diam)


def test_match_argmax_mode_and_bank(cuda):
    """mode='argmax' (the reference's path, evaluator.py:89-93) with a 3-object bank and per-frame obj_id."""
    from gadm_b200 import matching, synth
    B, N, M, d = 4, 640, 1024, 128
    rgbd, mesh, _ = synth.descriptors(B, N, M, d, n_obj=3, regime="planted", seed=5)
    xyz = synth.model_bank_xyz(3, M)
    bank = matching.ModelBank(mesh.to(cuda), xyz.to(cuda))
    obj = [2, 0, 1, 2]
    idx, sim, w, sx = matching.match(rgbd.to(cuda), bank, obj_id=obj, mode="argmax")
    assert w is None and sx is None
    for b in range(B):
        ridx, rsim, margin = mo.match_hard(rgbd[b], mesh[obj[b]])
        ok = margin > TOL
        assert torch.equal(idx[b].cpu()[ok], ridx[ok])
        assert (sim[b].cpu() - rsim).abs().max() <= TOL


@pytest.mark.parametrize("pad_mode", ["minus_one", "e0"])
def test_match_pad_modes(cuda, pad_mode):
    """Padded variants: pvn3d_eval_utils_kpls.py:436-444 (-1 column), geoMatch_DGCNN.py:92-99 (e0 column)."""
    from gadm_b200 import matching, synth
    N, M, d = 900, 512, 128
    rgbd, mesh, _ = synth.descriptors(1, N, M, d, regime="random", seed=77)
    if pad_mode == "minus_one":       # make the pad column win on some rows: rows with all-negative descriptors
        rgbd[0, :, :100] = -rgbd[0, :, :100].abs()
    else:
        rgbd[0, 0, :100] = 30.0
    xyz = synth.fibonacci_sphere(M, 0.2)
    idx, sim, _, _ = matching.match(rgbd.to(cuda), mesh.to(cuda), xyz[None].to(cuda), pad_mode=pad_mode,
                                    mode="argmax")
    ridx, rsim, margin = mo.match_hard(rgbd[0], mesh[0], pad_mode=pad_mode)
    ok = margin > TOL
    assert (ridx == M).sum() >= 50, "test must exercise the pad column"
    assert torch.equal(idx[0].cpu()[ok], ridx[ok])
    assert (sim[0].cpu() - rsim).abs().max() <= TOL


def test_match_mask_rows(cuda):
    """Rows outside the seg mask get idx = -1 (evaluator.py:78-88 selects rows; we keep positions)."""
    from gadm_b200 import matching, synth
    N, M, d = 500, 256, 64
    rgbd, mesh, _ = synth.descriptors(1, N, M, d, regime="planted", seed=9)
    g = torch.Generator().manual_seed(1)
    seg = torch.randn((2, N), generator=g)
    mask = mo.seg_mask(seg)
    xyz = synth.fibonacci_sphere(M, 0.1)
    out = matching.match(rgbd.to(cuda), mesh.to(cuda), xyz[None].to(cuda), mask=mask[None].to(cuda))
    idx = out[0][0].cpu()
    assert torch.all(idx[~mask] == -1)
    ref = mo.match_soft(rgbd[0], mesh[0], xyz, row_mask=mask)
    _check([o[0] for o in out], ref, M, 0.1, rows=mask)


@pytest.mark.parametrize("mode", ["soft", "argmax"])
def test_match_row_compaction_vs_oracle(cuda, mode):
    """evaluator.py:82-93: cls_msk -> rgbd_features[cls_msk] -> normalize -> matmul -> max.  With a mask the rows are
    compacted on the device and only they are matched; compact=True returns the reference's own (compacted) ordering,
    the default scatters back.  Frames with ~20 %, 0, 100 % and 1 selected rows; ragged row / model tiles."""
    from gadm_b200 import matching, synth
    B, N, M, d = 4, 3000, 2056, 128
    rgbd, mesh, _ = synth.descriptors(B, N, M, d, regime="planted", seed=31)
    diam = 0.2
    xyz = synth.fibonacci_sphere(M, diam)
    g = torch.Generator().manual_seed(7)
    mask = torch.rand((B, N), generator=g) < 0.2
    mask[1] = False
    mask[2] = True
    mask[3] = False
    mask[3, 1777] = True
    bank = matching.ModelBank(mesh.to(cuda), xyz[None].to(cuda))
    out = matching.match(rgbd.to(cuda), bank, mask=mask.to(cuda), mode=mode, compact=True)
    n_sel = out[4].cpu()
    assert n_sel.tolist() == mask.sum(1).tolist()
    scat = matching.match(rgbd.to(cuda), bank, mask=mask.to(cuda), mode=mode)
    full = matching.match(rgbd.to(cuda), bank, mode=mode)
    for b in range(B):
        n = int(n_sel[b])
        assert torch.all(out[0][b, n:] == -1) and torch.all(out[1][b, n:] == 0)
        assert torch.all(scat[0][b].cpu()[~mask[b]] == -1) and torch.all(scat[1][b].cpu()[~mask[b]] == 0)
        if n == 0:
            continue
        ref = mo.match_soft(rgbd[b], mesh[0], xyz, row_mask=mask[b])            # the reference's compacted rows
        got = [o[b, :n] if o is not None else None for o in out[:4]]
        if n > 10:
            _check(got, ref, M, diam, soft=mode == "soft")
        else:
            assert (got[1].cpu() - ref["max_sim"]).abs().max() <= TOL
        # scatter-back == compacted results at the selected positions == the unmasked launch at those positions
        for k in range(4):
            if out[k] is None:
                continue
            assert torch.equal(scat[k][b].cpu()[mask[b]], out[k][b, :n].cpu())
            assert torch.equal(scat[k][b].cpu()[mask[b]], full[k][b].cpu()[mask[b]])


@pytest.mark.parametrize("frac", [0.02, 0.1, 0.4])
@pytest.mark.parametrize("B,N,M", [(3, 1500, 2056), (1, 12800, 8192), (2, 900, 520)])
def test_match_few_selected_rows_fewer_units_than_sms(cuda, B, N, M, frac):
    """A mask that keeps so few rows that the persistent ARGMAX kernel has fewer (row block, model tile) units than the
    GPU has SMs -- one object instance with some hundred foreground points.  (The unit count is only known on the
    device then; CTAs without units must not sit between the CTAs that share a row block, or its merge never ends and
    every index stays -1.)  Against the oracle on the compacted rows, both modes."""
    from gadm_b200 import matching, synth
    d = 128
    rgbd, mesh, _ = synth.descriptors(B, N, M, d, n_obj=1, regime="planted", seed=5 + N)
    xyz = synth.fibonacci_sphere(M, 0.2)
    g = torch.Generator().manual_seed(N + int(frac * 100))
    mask = torch.rand((B, N), generator=g) < frac
    bank = matching.ModelBank(mesh.to(cuda), xyz[None].to(cuda))
    for mode in ("argmax", "soft"):
        out = matching.match(rgbd.to(cuda), bank, mask=mask.to(cuda), mode=mode)
        for b in range(B):
            ref = mo.match_soft(rgbd[b], mesh[0], xyz, row_mask=mask[b])
            idx = out[0][b].cpu()
            assert torch.all(idx[~mask[b]] == -1)
            ok = ref["margin"] > TOL
            assert torch.equal(idx[mask[b]][ok], ref["idx"][ok])
            assert (out[1][b].cpu()[mask[b]] - ref["max_sim"]).abs().max() <= TOL


def test_match_obj_id_is_validated(cuda):
    """A bank slot outside [0, n_obj) (e.g. a 1-based YCB-V class id used as is) is rejected on the host."""
    from gadm_b200 import matching, synth
    rgbd, mesh, _ = synth.descriptors(2, 300, 256, 64, n_obj=3, regime="random", seed=3)
    bank = matching.ModelBank(mesh.to(cuda), synth.model_bank_xyz(3, 256).to(cuda))
    with pytest.raises(ValueError):
        matching.match(rgbd.to(cuda), bank, obj_id=[1, 3])
    with pytest.raises(ValueError):
        matching.match(rgbd.to(cuda), bank, obj_id=torch.tensor([-1, 0]))
    # a device-side id cannot be checked without a synchronisation: the kernels clamp it (no out-of-bounds read)
    idx = matching.match(rgbd.to(cuda), bank, obj_id=torch.tensor([7, 0], device=cuda), mode="argmax")[0]
    ref = matching.match(rgbd.to(cuda), bank, obj_id=[2, 0], mode="argmax")[0]
    assert torch.equal(idx, ref)


def test_match_bf16x3_fp32_inputs(cuda):
    """operand_mode='bf16x3': arbitrary fp32 descriptors, ~fp32-faithful similarities (error ~1e-6)."""
    from gadm_b200 import matching, synth
    N, M, d = 512, 1024, 128
    g = torch.Generator().manual_seed(3)
    rgbd, mesh = torch.randn((1, d, N), generator=g), torch.randn((1, d, M), generator=g)
    xyz = synth.fibonacci_sphere(M, 0.2)
    out = matching.match(rgbd.to(cuda), mesh.to(cuda), xyz[None].to(cuda), operand_mode="bf16x3")
    ref = mo.match_soft(rgbd[0], mesh[0], xyz)
    assert (out[1][0].cpu() - ref["max_sim"]).abs().max() < 2e-5
    decided = ref["margin"] > 1e-4
    assert torch.equal(out[0][0].cpu()[decided], ref["idx"][decided])
    _check([o[0] for o in out], ref, M, 0.2)


def test_match_full_size_properties(cuda):
    """BASELINE shape 12800 x 8192 x 128: planted correspondences are recovered, weights in (0, 1],
    soft_xyz inside the model's bounding sphere, argmax mode == soft mode indices, and EVERY row of the frame
    agrees with the oracle (four CPU GEMMs of 3200 rows)."""
    from gadm_b200 import matching, synth
    N, M, d = 12800, 8192, 128
    rgbd, mesh, corr = synth.descriptors(1, N, M, d, regime="planted", seed=2000, sigma=0.5)
    diam = 0.2
    xyz = synth.fibonacci_sphere(M, diam)
    bank = matching.ModelBank(mesh.to(cuda), xyz[None].to(cuda))
    idx, sim, w, sx = matching.match(rgbd.to(cuda), bank)
    idx2, sim2, _, _ = matching.match(rgbd.to(cuda), bank, mode="argmax")
    assert torch.equal(idx, idx2) and torch.equal(sim, sim2)
    assert (idx[0].cpu() == corr[0]).float().mean() > 0.99
    assert torch.all((w > 0) & (w <= 1 + 1e-6))
    assert torch.all(sx.norm(dim=-1) <= diam / 2 * (1 + 1e-4))
    for r0 in range(0, N, 3200):
        rows = slice(r0, r0 + 3200)
        ref = mo.match_soft(rgbd[0][:, rows], mesh[0], xyz)
        _check([idx[0][rows], sim[0][rows], w[0][rows], sx[0][rows]], ref, M, diam)


def test_match_errors(cuda):
    from gadm_b200 import matching, _lib
    with pytest.raises(_lib.GadmError):      # d not a multiple of 64
        matching.match(torch.randn(1, 96, 64, device=cuda), torch.randn(1, 96, 64, device=cuda),
                       torch.zeros(1, 64, 3, device=cuda))
    with pytest.raises((_lib.GadmError, NotImplementedError)):      # CPU tensors: no fallback, no CPU kernel registered
        matching.match(torch.randn(1, 128, 64), torch.randn(1, 128, 64), torch.zeros(1, 64, 3))


def test_frame_poses_recovers_pose(cuda):
    """cal_frame_poses drop-in: planted correspondences + a known rigid transform -> Kabsch recovers it
    (evaluator.py:60-102 + best_fit_transform)."""
    from gadm_b200 import matching, synth
    N, M, d = 2048, 1024, 128
    rgbd, mesh, corr = synth.descriptors(1, N, M, d, regime="planted", seed=31, sigma=0.3)
    xyz = synth.fibonacci_sphere(M, 0.2)
    ang = 0.7
    R = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], dtype=torch.float32)
    t = torch.tensor([0.1, -0.05, 0.8])
    cld = (xyz[corr[0]] @ R.T + t).T.contiguous()              # [3, N]
    seg = torch.stack([torch.zeros(N), torch.ones(N)])          # all foreground
    seg[0, ::7] = 2.0                                           # every 7th point background
    bank = matching.ModelBank(mesh.to(cuda), xyz[None].to(cuda))
    item = (cld.to(cuda), seg.to(cuda), None, rgbd[0].to(cuda), 0, True)
    RT = matching.cal_frame_poses(item, bank)
    assert RT.shape == (3, 4)
    assert np.abs(RT[:, :3] - R.numpy()).max() < 1e-3 and np.abs(RT[:, 3] - t.numpy()).max() < 1e-3
    # oracle Kabsch on the oracle matcher's output
    mask = mo.seg_mask(seg)
    ridx, _, _ = mo.match_hard(rgbd[0], mesh[0], row_mask=mask)
    RT_ref = mo.best_fit_transform(xyz[ridx], cld.T[mask]).numpy()
    assert np.abs(RT - RT_ref).max() < 1e-4
    # early-outs: not detected -> sentinel pose (evaluator.py:69-73)
    RT0 = matching.cal_frame_poses((cld.to(cuda), seg.to(cuda), None, rgbd[0].to(cuda), 0, False), bank)
    assert RT0[2, 3] == -1000


def test_device_pose_fit_vs_best_fit_transform(cuda):
    """gadm_kabsch_moments(_w) + gadm_kabsch_poses (batched 3x3 Jacobi SVD, reflection fix, sentinel) on
    the device against the restated best_fit_transform (utils/pvn3d_eval_utils_kpls.py:43-76) at 1e-5: clean rigid
    motions, heavy noise (reflection fix exercised), planar and collinear point sets (rank-deficient H), too few
    pairs / not detected (evaluator.py:69-72, :83, :96), and the weighted fit."""
    from gadm_b200 import matching, ops, synth
    g = torch.Generator().manual_seed(11)
    B, N, M = 12, 400, 512
    xyz = synth.fibonacci_sphere(M, 0.2)
    bank = matching.ModelBank(synth.bf16_round(torch.randn((1, 64, M), generator=g)).to(cuda), xyz[None].to(cuda))
    idx = torch.randint(0, M, (B, N), generator=g)
    A = xyz[idx]                                                  # [B, N, 3] model points of the matches
    cloud = torch.empty((B, N, 3))
    for b in range(B):
        q, _ = torch.linalg.qr(torch.randn((3, 3), generator=g))
        if torch.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        noise = [0.0, 1e-3, 0.05, 0.3][b % 4]                     # 0.3 >> the model radius: the fit is near-degenerate
        cloud[b] = A[b] @ q.T + torch.randn(3, generator=g) + noise * torch.randn((N, 3), generator=g)
    cloud[8] = torch.randn((N, 3), generator=g) * torch.tensor([1.0, 1.0, 0.0])          # planar camera points
    idx[9] = idx[9, 0]                                                                    # one model vertex only
    A = xyz[idx]
    mask = (torch.rand((B, N), generator=g) < 0.7)
    mask[10] = False; mask[10, :3] = True                                                 # 3 pairs < min_pts
    det = torch.ones(B, dtype=torch.uint8); det[11] = 0
    w = torch.rand((B, N), generator=g) + 0.01
    mom = ops.kabsch_moments(idx.to(cuda), mask.to(torch.uint8).to(cuda), cloud.to(cuda), bank.aux, None, M, 1)
    poses = ops.kabsch_poses(mom, None, det.to(cuda), 5).cpu()
    momw = ops.kabsch_moments_w(idx.to(cuda), mask.to(torch.uint8).to(cuda), w.to(cuda), cloud.to(cuda), bank.aux, None, M, 1)
    posesw = ops.kabsch_poses(momw, mom, det.to(cuda), 5).cpu()
    flips = 0
    for b in range(B):
        sel = mask[b]
        if b in (10, 11):
            for P in (poses[b], posesw[b]):
                assert torch.equal(P[:, :3], torch.eye(3)) and P[2, 3] == -1000 and P[0, 3] == 0 and P[1, 3] == 0
            continue
        ref = mo.best_fit_transform(A[b][sel], cloud[b][sel])
        refw = mo.best_fit_transform_weighted(A[b][sel], cloud[b][sel], w[b][sel])
        H = (A[b][sel].double() - A[b][sel].double().mean(0)).T @ (cloud[b][sel].double() - cloud[b][sel].double().mean(0))
        U, S, Vt = torch.linalg.svd(H)
        flips += int(torch.linalg.det(Vt.T @ U.T) < 0)
        if S[2] > 1e-9 * S[0] and b != 9:                          # unique optimum: compare the matrices
            assert (poses[b].double() - ref).abs().max() < 1e-5, (b, (poses[b].double() - ref).abs().max())
            assert (posesw[b].double() - refw).abs().max() < 1e-5
        # in every case: a proper rotation that attains the optimum's residual
        R = poses[b][:, :3].double()
        assert (R @ R.T - torch.eye(3, dtype=torch.float64)).abs().max() < 1e-6 and abs(float(torch.linalg.det(R)) - 1) < 1e-6
        res = lambda T: ((A[b][sel].double() @ T[:, :3].T + T[:, 3]) - cloud[b][sel].double()).pow(2).sum()
        assert res(poses[b].double()) <= res(ref) * (1 + 1e-6) + 1e-9
    assert flips >= 1, "the reflection fix must be exercised"
    # the whole path on the device: frame_poses_device == the host list of frame_poses
    rgbd, mesh, corr = synth.descriptors(2, 1024, M, 64, regime="planted", seed=5, sigma=0.3)
    bank2 = matching.ModelBank(mesh.to(cuda), xyz[None].to(cuda))
    cld = torch.stack([(xyz[corr[b]] @ torch.eye(3) + torch.tensor([0.0, 0.1 * b, 0.7])).T for b in range(2)])
    seg = torch.stack([torch.zeros((2, 1024)), torch.ones((2, 1024))], dim=1)
    pd = matching.frame_poses_device(cld.to(cuda), seg.to(cuda), rgbd.to(cuda), bank2)
    assert pd.is_cuda and pd.shape == (2, 3, 4)
    assert (pd.cpu()[:, :, :3] - torch.eye(3)).abs().max() < 1e-3 and (pd.cpu()[1, :, 3] - torch.tensor([0.0, 0.1, 0.7])).abs().max() < 1e-3
    pw = matching.frame_poses_device(cld.to(cuda), seg.to(cuda), rgbd.to(cuda), bank2, weighted=True)
    assert (pw.cpu() - pd.cpu()).abs().max() < 1e-3


def test_seg_mask_matches_torch_argmax(cuda):
    """the foreground mask of evaluator.py:78,82 in one kernel (ties -> background, as torch.argmax
    returns the first maximal index)."""
    from gadm_b200 import ops
    g = torch.Generator().manual_seed(12)
    seg = torch.randn((3, 2, 1001), generator=g)
    seg[:, 1, ::7] = seg[:, 0, ::7]                       # exact ties
    got = ops.seg_mask(seg.to(cuda)).cpu()
    assert got.dtype == torch.uint8 and got.shape == (3, 1001)
    assert torch.equal(got.bool(), torch.argmax(seg, dim=1) == 1)


@pytest.fixture
def cfg():
    """gadm_config_set switches for one test; restored to automatic afterwards."""
    from gadm_b200 import _lib
    used = []

    def set_(switches):
        for k, v in switches.items():
            _lib.config_set(k, v)
            used.append(k)
    yield set_
    for k in used:
        _lib.config_set(k, -1)


@pytest.mark.parametrize("env", [{"match.alt": 0, "match.pair": 0, "match.rt": 1},
                                 {"match.alt": 0, "match.pair": 0, "match.rt": 2},
                                 {"match.alt": 0, "match.pair": 1}, {"match.alt": 0, "match.pair": 1, "match.cta2": 1},
                                 {"match.alt": 1}, {"match.ctas": 5}, {"match.ctas": 37}, {"match.ctas": 1},
                                 {"match.alt_cta2": 0}, {"match.alt_cta2": 0, "match.ctas": 5},
                                 {"match.alt_cta2": 0, "match.ctas": 37}, {"match.alt_cta2": 0, "match.ctas": 1}])
def test_match_kernel_variants_agree(cuda, cfg, env):
    """The kernels of the matcher (one row tile per CTA, two row tiles per CTA, paired rows per thread with and
    without CTA pairs -- cta_group::2 MMAs over a cluster of two row blocks --, the
    persistent kernels with their units dealt out to 148 / 37 / 5 / 1 CTAs -- row blocks split over two or more
    CTAs and merged by the last to arrive) are selected per launch; every one of them must meet the same gates on a
    ragged shape, in both modes."""
    from gadm_b200 import matching, synth
    from oracle import match_oracle as mo
    cfg(env)
    N, M, d = 1500, 2056, 128                  # ragged in rows (1500 = 5 * 256 + 220) and in model tiles (2056 = 8 * 256 + 8)
    rgbd, mesh, _ = synth.descriptors(1, N, M, d, regime="planted", seed=77)
    xyz = synth.fibonacci_sphere(M, 0.2)
    ref = mo.match_soft(rgbd[0], mesh[0], xyz)
    ok = ref["margin"] > 1e-3
    for mode in ("soft", "argmax"):
        idx, sim, w, sx = matching.match(rgbd.to(cuda), mesh.to(cuda), xyz[None].to(cuda), mode=mode)
        assert torch.equal(idx[0].cpu()[ok], ref["idx"][ok])
        assert (sim[0].cpu() - ref["max_sim"]).abs().max() <= 1e-3
        if mode == "soft":
            assert ((w[0].cpu() - ref["weight"]).abs() / ref["weight"]).max() <= 1e-3
            assert (sx[0].cpu() - ref["soft_xyz"]).abs().max() <= 1e-3 * 0.2


def test_match_bf16n_operands_and_unit_argmax(cuda):
    """operand_mode='bf16n': model columns are normalised (F.normalize, evaluator.py:90) BEFORE the one rounding to
    bf16.  (1) every mode stays exact with respect to the rounded operands: the oracle fed the rounded normalised
    columns meets the usual gates; (2) mode='argmax_unit' (no per-column scale, the fastest kernel) meets the 1e-3
    gates against the oracle on the ORIGINAL descriptors and agrees with mode='argmax' up to ||bf16 column|| - 1."""
    import torch.nn.functional as F
    from gadm_b200 import matching, synth
    B, N, M, d = 2, 1500, 2056, 128            # ragged rows and model tiles
    rgbd, mesh, _ = synth.descriptors(B, N, M, d, regime="planted", seed=91)
    diam = 0.2
    xyz = synth.fibonacci_sphere(M, diam)
    bank = matching.ModelBank(mesh.to(cuda), xyz[None].to(cuda), operand_mode="bf16n")
    mesh_n = bank.cols[0].float().cpu().T.contiguous()                 # what the tensor core sees, [d, M]
    want = F.normalize(mesh[0], p=2, dim=0)
    assert torch.all((mesh_n - want).abs() <= 2 ** -8 * want.abs() + 1e-30)  # one bf16 rounding of the normalised column
    assert (mesh_n != synth.bf16_round(want)).float().mean() < 1e-3    # (the fp32 norm may differ in the last bit)
    out = matching.match(rgbd.to(cuda), bank, operand_mode="bf16n")
    hard = matching.match(rgbd.to(cuda), bank, operand_mode="bf16n", mode="argmax")
    unit = matching.match(rgbd.to(cuda), bank, operand_mode="bf16n", mode="argmax_unit")
    assert unit[2] is None and unit[3] is None
    # mode="argmax" on a bf16n bank runs GADM_MATCH_ARGMAX_BF16N (chunk pruning): bit-identical to the unpruned kernel
    from gadm_b200 import ops
    from gadm_b200._lib import MATCH_MODES
    rows, rinv, pad = ops.prep_rows(rgbd.to(cuda), 2, 0)
    plain = ops.match_fwd(rows, rinv, pad, bank.cols, bank.aux, None, None, 16.0, 0, MATCH_MODES["argmax"])
    assert torch.equal(plain[0], hard[0]) and torch.equal(plain[1], hard[1])
    for b in range(B):
        ref_n = mo.match_soft(rgbd[b], mesh_n, xyz)                    # (1) exact w.r.t. the rounded operands
        _check([o[b] for o in out], ref_n, M, diam)
        assert torch.equal(hard[0][b], out[0][b])
        ref = mo.match_soft(rgbd[b], mesh[0], xyz)                     # (2) the reference on the original inputs
        ok = ref["margin"] > TOL
        assert torch.equal(unit[0][b].cpu()[ok], ref["idx"][ok])
        assert (unit[1][b].cpu() - ref["max_sim"]).abs().max() <= TOL
        assert (mesh_n.norm(dim=0) - 1).abs().max() < 2 ** -8          # the search ignores this factor ...
        same = (unit[0][b] == hard[0][b]).cpu()
        assert same.float().mean() > 0.999
        # ... but the winner's similarity is reported with its true column scale
        assert torch.allclose(unit[1][b].cpu()[same], hard[1][b].cpu()[same], atol=2e-6)
    with pytest.raises(ValueError):                                    # unit scales need normalised columns
        matching.match(rgbd.to(cuda), matching.ModelBank(mesh.to(cuda), xyz[None].to(cuda)), mode="argmax_unit")


@pytest.mark.parametrize("ctas", [0, 2, 14, 74])
@pytest.mark.parametrize("B,N,M", [(3, 1500, 2056), (1, 257, 520), (2, 3333, 8192), (5, 700, 264)])
def test_match_argmax_cta_pairs_change_nothing(cuda, cfg, B, N, M, ctas):
    """The persistent ARGMAX kernel runs as CTA pairs by default (cta_group::2 MMAs, half a model tile per CTA, units =
    pairs of row blocks; match.alt_cta2 = 0: single CTAs).  Indices and similarities are bit-identical between the two: odd numbers of row
    blocks per frame (the second CTA of the last pair has no rows), ragged model tiles, row blocks split over several
    pairs and merged, compacted rows (device-side row counts), every ARGMAX flavour."""
    from gadm_b200 import matching, ops, synth
    d = 128
    rgbd, mesh, _ = synth.descriptors(B, N, M, d, n_obj=1, regime="random", seed=31 + N)
    xyz = synth.fibonacci_sphere(M, 0.2)[None].to(cuda)
    g = torch.Generator().manual_seed(N)
    mask = (torch.rand((B, N), generator=g) < 0.4).to(cuda)
    bank = matching.ModelBank(mesh.to(cuda), xyz)
    bank_n = matching.ModelBank(mesh.to(cuda), xyz, operand_mode="bf16n")

    def run():
        outs = [matching.match(rgbd.to(cuda), bank, mode="argmax"),
                matching.match(rgbd.to(cuda), bank, mode="argmax", mask=mask),
                matching.match(rgbd.to(cuda), bank_n, mode="argmax", operand_mode="bf16n"),
                matching.match(rgbd.to(cuda), bank_n, mode="argmax_unit", operand_mode="bf16n")]
        return [(o[0].clone(), o[1].clone()) for o in outs]
    cfg({"match.alt_cta2": 0})
    want = run()
    cfg({"match.alt_cta2": 1, **({"match.ctas": ctas} if ctas else {})})
    got = run()
    for (wi, ws), (gi, gs) in zip(want, got):
        assert torch.equal(wi, gi) and torch.equal(ws, gs)


@pytest.mark.parametrize("env", [{}, {"match.alt": 0}, {"match.alt": 0, "match.rt": 1},
                                 {"match.alt": 0, "match.pair": 1}, {"match.alt": 0, "match.pair": 1, "match.cta2": 1},
                                 {"match.ctas": 3}, {"match.ctas": 11}, {"match.ctas": 50}, {"match.alt_cta2": 0},
                                 {"match.alt_cta2": 0, "match.ctas": 3}, {"match.alt_cta2": 0, "match.ctas": 11}])
def test_match_exact_ties_first_index_wins(cuda, cfg, env):
    """torch.max returns the FIRST maximal index of the scores it is given (evaluator.py:93).  Model vertices duplicated bit for bit across
    groups, chunks, column slices, model tiles and (alternating kernel) beyond the tiles after which the slices
    exchange their running maxima: on every row whose best vertex has copies the smallest index must win -- the stash
    look-up, the slice / quad merges and the 'no record' sentinel of the exchanges all have to break ties that way."""
    from gadm_b200 import matching, synth
    cfg(env)
    N, M, d = 777, 4608, 128                   # 18 model tiles of 256 (24 of 192), ragged rows
    g = torch.Generator().manual_seed(5)
    mesh = synth.bf16_round(torch.randn((1, d, M), generator=g))
    src = torch.randperm(512, generator=g)[:200]                      # originals in the first two tiles
    copies = {}
    for i, c in enumerate(src.tolist()):
        # a copy in the same 8-column group / chunk / slice / tile, and copies in later tiles (before and after
        # the exchanges that follow tiles 0, 1, 3, 7, 15)
        offs = [(c // 8) * 8 + (c + 3) % 8, c + 32 if c % 64 < 32 else c - 32, (c + 64) % 256 + (c // 256) * 256,
                c + 256 * (1 + i % 3), c + 256 * (4 + i % 12)]
        for o in offs:
            if o != c and o not in copies and o not in src.tolist() and 0 <= o < M:
                mesh[0][:, o] = mesh[0][:, c]
                copies[o] = c
    corr = torch.randint(0, M, (1, N), generator=g)                    # planted AFTER the duplication
    rgbd = synth.bf16_round(mesh[0][:, corr[0]] + 0.3 * torch.randn((d, N), generator=g))[None]
    # rows planted on a duplicated vertex (or on one of its copies) must report the smallest member of the set
    groups = {}
    for o, c in copies.items():
        groups.setdefault(c, {c}).add(o)
    want = corr[0].clone()
    member = {}
    for c, grp in groups.items():
        for x in grp:
            member[x] = min(grp)
    hit = torch.tensor([int(c) in member for c in corr[0].tolist()])
    assert hit.sum() > 30
    for i in torch.where(hit)[0].tolist():
        want[i] = member[int(corr[0][i])]
    xyz = synth.fibonacci_sphere(M, 0.2)
    # (the CPU reference is no arbiter here: its SGEMM sums different columns in different orders, so bit-identical
    # vertices get scores that differ in the last bit and torch.max picks any of them; the tensor core accumulates
    # identical operands identically, so the kernel's scores tie exactly and the contract "first index" is testable)
    ref_idx, _, _ = mo.match_hard(rgbd[0], mesh[0])
    assert all(member.get(int(a), int(a)) == int(b) for a, b in zip(ref_idx[hit].tolist(), want[hit].tolist()))
    for mode in ("argmax", "soft"):
        idx = matching.match(rgbd.to(cuda), mesh.to(cuda), xyz[None].to(cuda), mode=mode)[0][0].cpu()
        assert torch.equal(idx[hit], want[hit]), f"{mode}: a later copy displaced the first maximal index"


@pytest.mark.parametrize("N,M,d", [(5, 8, 64), (129, 264, 128), (128, 256, 64), (257, 8, 128), (1, 8192, 128)])
def test_match_tiny_and_boundary_shapes(cuda, N, M, d):
    """Smallest legal sizes and tile boundaries (one scene point, 8 model vertices, exactly one row tile, one row past
    it): every launch path (single-row-tile fallbacks included), both modes, with and without an all-zero mask."""
    from gadm_b200 import matching, synth
    rgbd, mesh, _ = synth.descriptors(2, N, M, d, regime="random", seed=700 + N + M)
    diam = 0.2
    xyz = synth.fibonacci_sphere(M, diam)
    bank = matching.ModelBank(mesh.to(cuda), xyz[None].to(cuda))
    soft = matching.match(rgbd.to(cuda), bank)
    hard = matching.match(rgbd.to(cuda), bank, mode="argmax")
    for b in range(2):
        ref = mo.match_soft(rgbd[b], mesh[0], xyz)
        ok = ref["margin"] > TOL
        for out in (soft, hard):
            assert torch.equal(out[0][b].cpu()[ok], ref["idx"][ok])
            assert (out[1][b].cpu() - ref["max_sim"]).abs().max() <= TOL
        assert ((soft[2][b].cpu() - ref["weight"]).abs() / ref["weight"]).max() <= TOL
        assert (soft[3][b].cpu() - ref["soft_xyz"]).abs().max() <= TOL * diam
    none = torch.zeros((2, N), dtype=torch.bool, device=cuda)
    for mode in ("soft", "argmax"):
        idx, sim, w, sx = matching.match(rgbd.to(cuda), bank, mask=none, mode=mode)
        assert torch.all(idx == -1) and torch.all(sim == 0)
        if mode == "soft":
            assert torch.all(w == 0) and torch.all(sx == 0)
