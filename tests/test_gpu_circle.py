"""-m gpu: fused CircleLoss forward (gadm_circle_loss_fwd) vs oracle/circle_oracle.py (which
reproduces the reference's own loss to 1e-6 on the golden fixture) on the same bf16-representable inputs.
Gate: fp32 path, 1e-3 relative on the scalar loss and on every row's loss (+1e-3 absolute for rows near 0)."""
import os

import numpy as np
import pytest
import torch

from oracle import circle_oracle as co

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-3


def _check(rgbd, mesh, labels, match_idx, xyz, vis, r, cuda, obj_id=None, bank=None):
    from gadm_b200 import matching
    if bank is None:
        bank = matching.ModelBank(mesh.to(cuda), xyz.to(cuda))
    total, rows, lse_p, lse_n = matching.circle_match_loss(rgbd.to(cuda), bank, labels.to(cuda), match_idx.to(cuda),
                                                           vis.to(cuda), r, obj_id=obj_id, return_rows=True)
    total2 = matching.circle_match_loss(rgbd.to(cuda), mesh.to(cuda), labels.to(cuda), match_idx.to(cuda), vis.to(cuda),
                                        r, model_xyz=xyz.to(cuda), obj_id=obj_id)        # raw features instead of a bank
    assert float(total2) == float(total)
    B = rgbd.shape[0]
    per = []
    for b in range(B):
        o = 0 if obj_id is None else int(obj_id[b])
        if int((labels[b] == 1).sum()) < 3:
            continue
        idxs, want, wp, wn = co.sample_rows(rgbd[b], mesh[o], labels[b], match_idx[b], xyz[o], vis[b], r)
        got = rows[b].cpu()[idxs]
        assert torch.all((got - want).abs() <= TOL * want.abs() + TOL), f"row loss, sample {b}"
        fin = torch.isfinite(wp)
        assert torch.all((lse_p[b].cpu()[idxs][fin] - wp[fin]).abs() <= TOL * wp[fin].abs() + TOL)
        assert torch.all((lse_n[b].cpu()[idxs] - wn).abs() <= TOL * wn.abs() + TOL)
        off = torch.ones(labels.shape[1], dtype=torch.bool)
        off[idxs] = False
        assert torch.all(rows[b].cpu()[off] == 0)          # rows outside the foreground take no part
        per.append(want.mean())
    want_total = torch.stack(per).mean() if per else torch.tensor(0.0)
    assert abs(float(total) - float(want_total)) <= TOL * abs(float(want_total)) + 1e-6
    return float(total)


def test_circle_loss_golden_fixture(cuda):
    g = np.load(os.path.join(GOLD, "circle_golden.npz"))
    t = lambda k: torch.from_numpy(g[k])
    total = _check(t("rgbd"), t("mesh"), t("labels"), t("match_idx").long(), t("xyz")[None], t("vis"),
                   float(g["positive_r"]), cuda)
    assert abs(total - float(g["ref_total"])) <= TOL * float(g["ref_total"])      # the reference's own number


def test_circle_loss_dgcnn_variant_golden(cuda):
    """models/geoMatch_DGCNN.py:53-78, :80-136: e0 pad column, per-vertex positive radius positive_r / 1000 * depth."""
    from gadm_b200 import matching
    g = np.load(os.path.join(GOLD, "circle_golden.npz"))
    t = lambda k: torch.from_numpy(g[k])
    xyz, RT, pr = t("xyz"), t("RT"), float(g["dgcnn_positive_r"])
    rad = matching.dgcnn_positive_radius(xyz.to(cuda), RT.to(cuda), pr)
    want_rad = torch.stack([co.dgcnn_radius(xyz, RT[b], pr) for b in range(3)])
    assert torch.allclose(rad.cpu(), want_rad, rtol=1e-6, atol=1e-9)
    total, rows, _, _ = matching.circle_match_loss(t("rgbd").to(cuda), t("mesh").to(cuda), t("labels").to(cuda),
                                                   t("match_idx").to(cuda), t("vis").to(cuda), rad,
                                                   model_xyz=xyz[None].to(cuda), pad_mode="e0", return_rows=True)
    assert abs(float(total) - float(g["dgcnn_ref_total"])) <= TOL * float(g["dgcnn_ref_total"])
    for b in range(2):
        idxs, want, _, _ = co.sample_rows(t("rgbd")[b], t("mesh")[0], t("labels")[b], t("match_idx")[b].long(), xyz,
                                          t("vis")[b], want_rad[b], pad="e0")
        assert torch.all((rows[b].cpu()[idxs] - want).abs() <= TOL * want.abs() + TOL)


def test_circle_loss_sys_variant_golden(cuda):
    """GeoMatch.matching_loss_sys (models/geoMatch.py:86-100, the sys_corr_idx branch :138-141): the positives of a row
    are exactly the columns match_idx[n] and match_idx[sys_idx[n]].  Forward against the reference's own number, gradients
    against torch autograd through the oracle."""
    from gadm_b200 import matching
    g = np.load(os.path.join(GOLD, "circle_golden.npz"))
    t = lambda k: torch.from_numpy(g[k])
    rgbd = t("rgbd").to(cuda).requires_grad_(True)
    mesh = t("mesh").to(cuda).requires_grad_(True)
    total = matching.circle_match_loss(rgbd, mesh, t("labels").to(cuda), t("match_idx").to(cuda), None, None,
                                       model_xyz=t("xyz")[None].to(cuda), sys_idx=t("sys_idx").to(cuda))
    assert abs(float(total.detach()) - float(g["sys_ref_total"])) <= TOL * float(g["sys_ref_total"])
    total.backward()
    r2, m2 = t("rgbd").clone().requires_grad_(True), t("mesh").clone().requires_grad_(True)
    co.batch_loss_sys(r2, m2, t("labels"), t("match_idx"), t("sys_idx")).backward()
    for got, want in ((rgbd.grad.cpu(), r2.grad), (mesh.grad.cpu(), m2.grad)):
        assert (got - want).abs().max() <= 2e-3 * want.abs().max() + 1e-6


def test_circle_loss_ragged_bank(cuda):
    """Ragged rows / model tiles (1500 rows, 2056 vertices), a 2-object bank with per-frame obj_id, planted matches,
    per-frame visibility, 15 % of the rows off the model."""
    from gadm_b200 import synth
    B, N, M, d = 3, 1500, 2056, 128
    g = torch.Generator().manual_seed(41)
    mesh = synth.bf16_round(torch.randn((2, d, M), generator=g))
    xyz = torch.stack([synth.fibonacci_sphere(M, 0.2), synth.fibonacci_sphere(M, 0.15)])
    obj = [1, 0, 1]
    vis = torch.rand((B, M), generator=g) < 0.5
    labels = (torch.rand((B, N), generator=g) < 0.6).long()
    match_idx = torch.full((B, N), M, dtype=torch.int64)
    rgbd = synth.bf16_round(torch.randn((B, d, N), generator=g))
    for b in range(B):
        ids = torch.where(vis[b])[0]
        pick = ids[torch.randint(0, len(ids), (N,), generator=g)]
        on = torch.rand((N,), generator=g) < 0.85
        match_idx[b] = torch.where(on, pick, torch.full_like(pick, M))
        sel = on.nonzero()[:, 0]
        rgbd[b][:, sel] = synth.bf16_round(mesh[obj[b]][:, match_idx[b][sel]] + 0.7 * torch.randn((d, len(sel)), generator=g))
    total = _check(rgbd, mesh, labels, match_idx, xyz, vis, 0.012, cuda, obj_id=obj)
    assert total > 0


def test_circle_loss_bwd_split_is_the_fp32_gradient_with_the_norms_folded_in(cuda):
    """gadm_circle_loss_bwd_split against gadm_circle_loss_bwd on the same inputs: hi + lo of group g, element e =
    G[.., 8 g + e] * rinv_i * scale_j to 2^-16 relative (two bf16 parts), the pad column's gradient in g_pad, zeros in
    the padding; ragged N and M (the guard paths of the kernel)."""
    from gadm_b200 import ops, synth
    from gadm_b200.ops import OPERAND_MODES, PAD_MODES
    B, N, M, d = 2, 333, 520, 64
    g = torch.Generator().manual_seed(7)
    rgbd, mesh, _ = synth.descriptors(B, N, M, d, seed=11)
    xyz = synth.fibonacci_sphere(M, 0.2)[None].to(cuda)
    rows, rinv, pad_sim = ops.prep_rows(rgbd.to(cuda), OPERAND_MODES["bf16"], PAD_MODES["minus_one"])
    cols, aux = ops.prep_model(mesh[:1].to(cuda), xyz, OPERAND_MODES["bf16"])
    planes = torch.empty((4, B, M), device=cuda)
    planes[:3] = xyz[0].t()[:, None, :].expand(3, B, M)
    planes[3] = 0.05 ** 2
    mi = torch.randint(0, M + 1, (B, N), generator=g).to(cuda)
    fg = torch.ones((B, N), dtype=torch.uint8, device=cuda)
    _, lp, ln = ops.circle_loss_fwd(rows, rinv, pad_sim, cols, aux, planes, mi, fg, None, 16.0, 0.2)
    w = torch.rand((B, N), generator=g).to(cuda)
    G = ops.circle_loss_bwd(rows, rinv, pad_sim, cols, aux, planes, mi, None, 16.0, 0.2, lp, ln, w)
    G2, g_pad = ops.circle_loss_bwd_split(rows, rinv, pad_sim, cols, aux, planes, mi, None, 16.0, 0.2, lp, ln, w)
    Mp = M + 8
    assert G2.shape == (B, N, 2 * Mp) and G2.dtype == torch.bfloat16
    parts = G2.float().view(B, N, Mp // 8, 2, 8)
    got = (parts[:, :, :, 0] + parts[:, :, :, 1]).reshape(B, N, Mp)
    scale = aux[:M].view(1, 1, M)
    want = G[:, :, :M] * rinv[..., None] * scale
    err = (got[:, :, :M] - want).abs()
    assert bool((err <= 2.0 ** -15 * want.abs() + 1e-37).all()), float((err / want.abs().clamp(min=1e-30)).max())
    assert bool((parts[:, :, :, 1].abs() <= 2.0 ** -8 * parts[:, :, :, 0].abs() + 1e-37).all())   # lo is the remainder of hi
    assert float(got[:, :, M:].abs().max()) == 0.0
    assert torch.equal(g_pad, G[:, :, M])
    assert float(G.abs().max()) > 0


@pytest.mark.parametrize("B,N,M,d,sys2", [(2, 333, 520, 64, False), (1, 700, 1000, 128, False), (2, 130, 264, 128, True),
                                          (1, 1300, 8192, 128, False), (3, 50, 72, 64, False), (1, 128, 128, 128, True)])
def test_circle_loss_bwd_fused_products_match_the_library_products(cuda, B, N, M, d, sys2):
    """gadm_circle_loss_bwd_fused: G2 and g_pad bit-identical to gadm_circle_loss_bwd_split, and dF -- the second MMA of
    the kernel, G'' from shared memory against the resident model tile read MN-major -- equal to the same product
    computed from G2 by a library GEMM (fp32 accumulation in both, in different orders over 2 (M + 8) terms: 5e-5 of
    the largest entry).  Ragged rows / model
    tiles, both positive rules."""
    from gadm_b200 import ops, synth
    from gadm_b200.ops import OPERAND_MODES, PAD_MODES
    g = torch.Generator().manual_seed(17 + N)
    rgbd, mesh, _ = synth.descriptors(B, N, M, d, seed=3 + N)
    xyz = synth.fibonacci_sphere(M, 0.2)[None].to(cuda)
    rows, rinv, pad_sim = ops.prep_rows(rgbd.to(cuda), OPERAND_MODES["bf16"], PAD_MODES["minus_one"])
    cols, aux = ops.prep_model(mesh[:1].to(cuda), xyz, OPERAND_MODES["bf16"])
    planes = torch.empty((4, B, M), device=cuda)
    planes[:3] = xyz[0].t()[:, None, :].expand(3, B, M)
    planes[3] = 0.05 ** 2
    mi = torch.randint(0, M + 1, (B, N), generator=g).to(cuda)
    mi2 = torch.randint(0, M + 1, (B, N), generator=g).to(cuda) if sys2 else None
    fg = torch.ones((B, N), dtype=torch.uint8, device=cuda)
    _, lp, ln = ops.circle_loss_fwd(rows, rinv, pad_sim, cols, aux, planes, mi, fg, None, 16.0, 0.2, mi2)
    w = torch.rand((B, N), generator=g).to(cuda)
    G2, g_pad = ops.circle_loss_bwd_split(rows, rinv, pad_sim, cols, aux, planes, mi, None, 16.0, 0.2, lp, ln, w, mi2)
    G2f, g_padf, dF, _ = ops.circle_loss_bwd_fused(rows, rinv, pad_sim, cols, aux, planes, mi, None, 16.0, 0.2, lp, ln, w, mi2)
    assert torch.equal(G2.view(torch.int16), G2f.view(torch.int16)) and torch.equal(g_pad, g_padf)
    Mp = M + 8
    k = torch.arange(2 * Mp, device=cuda)
    col_of_k = (k // 16) * 8 + k % 8
    cols_p = torch.zeros((1, Mp, d), dtype=torch.bfloat16, device=cuda)
    cols_p[:, :M] = cols
    want = torch.bmm(G2, cols_p[:, col_of_k].expand(B, -1, -1), out_dtype=torch.float32)
    err = (dF - want).abs().max()
    assert float(want.abs().max()) > 0
    assert err <= 5e-5 * want.abs().max(), f"dF max err {float(err)} vs max {float(want.abs().max())}"
    # model_side=True: the model-side product in the kernel as well (third MMA, both operands MN-major, fp32 reductions
    # over the row tiles of a frame); dL/dsim is not written at all
    G0, g_padm, dFm, dM = ops.circle_loss_bwd_fused(rows, rinv, pad_sim, cols, aux, planes, mi, None, 16.0, 0.2, lp, ln, w,
                                                    mi2, True)
    assert G0.numel() == 0 and torch.equal(g_padm, g_pad) and torch.equal(dFm, dF)
    t = torch.bmm(G2.transpose(1, 2), rows, out_dtype=torch.float32)
    want_m = t.view(B, Mp // 8, 2, 8, d).sum(2).reshape(B, Mp, d)
    assert dM.shape == (B, Mp, d) and float(dM[:, M:].abs().max()) == 0.0
    errm = (dM - want_m).abs().max()
    assert float(want_m.abs().max()) > 0
    assert errm <= 5e-5 * want_m.abs().max(), f"dM max err {float(errm)} vs max {float(want_m.abs().max())}"


@pytest.mark.parametrize("grad_gemm,gate", [("fp32", 1e-3), ("tf32", 3e-3), ("bf16x2", 1e-3), ("fused", 1e-3), ("flash", 1e-3)])
def test_circle_loss_gradients_vs_autograd_of_the_reference_math(cuda, grad_gemm, gate):
    """d loss / d rgbd and d loss / d mesh against torch autograd through the oracle (the reference's own formulas,
    ap / an detached as at loss.py:479-480) on the CPU.  Gate: 1e-3 of the largest gradient entry (3e-3 when the two
    library gradient GEMMs are allowed to run in tf32)."""
    from gadm_b200 import matching, synth
    B, N, M, d = 2, 300, 520, 64
    g = torch.Generator().manual_seed(51)
    mesh = synth.bf16_round(torch.randn((1, d, M), generator=g))
    xyz = synth.fibonacci_sphere(M, 0.2)[None]
    vis = torch.rand((B, M), generator=g) < 0.6
    labels = (torch.rand((B, N), generator=g) < 0.5).long()
    match_idx = torch.full((B, N), M, dtype=torch.int64)
    rgbd = synth.bf16_round(torch.randn((B, d, N), generator=g))
    for b in range(B):
        ids = torch.where(vis[b])[0]
        pick = ids[torch.randint(0, len(ids), (N,), generator=g)]
        on = torch.rand((N,), generator=g) < 0.8
        match_idx[b] = torch.where(on, pick, torch.full_like(pick, M))
        sel = on.nonzero()[:, 0]
        rgbd[b][:, sel] = synth.bf16_round(mesh[0][:, match_idx[b][sel]] + 0.8 * torch.randn((d, len(sel)), generator=g))
    r = 0.03
    a = rgbd.clone().requires_grad_(True)
    m = mesh.clone().requires_grad_(True)
    want = co.batch_loss(a, m[0], labels, match_idx, xyz[0], vis, r)
    want.backward()
    ad = rgbd.to(cuda).requires_grad_(True)
    md = mesh.to(cuda).requires_grad_(True)
    got = matching.circle_match_loss(ad, md, labels.to(cuda), match_idx.to(cuda), vis.to(cuda), r, model_xyz=xyz.to(cuda),
                                     grad_gemm=grad_gemm)
    assert abs(float(got.detach()) - float(want.detach())) <= TOL * abs(float(want.detach()))
    (2.0 * got).backward()                                   # upstream gradient 2: checks the chain through g_total
    for name, gd, wd in (("rgbd", ad.grad.cpu() / 2, a.grad), ("mesh", md.grad.cpu() / 2, m.grad)):
        err = (gd - wd).abs().max()
        assert err <= gate * wd.abs().max(), f"d loss / d {name}: max err {err} vs max |grad| {wd.abs().max()}"
        assert wd.abs().max() > 0


@pytest.mark.parametrize("variant", ["e0_per_vertex_radius", "sys", "two_objects"])
@pytest.mark.parametrize("grad_gemm", ["bf16x2", "fused", "flash"])
def test_circle_loss_backward_modes_agree_on_every_variant(cuda, variant, grad_gemm):
    """The tensor-core / in-kernel gradient products against the fp32 library products (grad_gemm="fp32", itself gated
    against torch autograd above) on the variants that test does not cover: the DGCNN variant (e0 pad column, per-vertex
    radii), the symmetry-aware positives, a bank of two objects with per-frame object ids.  1e-3 of the largest entry."""
    from gadm_b200 import matching, synth
    B, N, M, d = 3, 420, 520, 128
    g = torch.Generator().manual_seed(77)
    n_obj = 2 if variant == "two_objects" else 1
    mesh = synth.bf16_round(torch.randn((n_obj, d, M), generator=g))
    xyz = torch.stack([synth.fibonacci_sphere(M, 0.2 - 0.03 * o) for o in range(n_obj)])
    obj = [1, 0, 1] if n_obj == 2 else None
    vis = torch.rand((B, M), generator=g) < 0.6
    labels = (torch.rand((B, N), generator=g) < 0.6).long()
    match_idx = torch.randint(0, M + 1, (B, N), generator=g)
    rgbd = synth.bf16_round(torch.randn((B, d, N), generator=g))
    kw = dict(model_xyz=xyz.to(cuda))
    if obj is not None:
        kw["obj_id"] = obj
    if variant == "e0_per_vertex_radius":
        kw["pad_mode"] = "e0"
        radius = (0.02 + 0.03 * torch.rand((B, M), generator=g)).to(cuda)
    else:
        radius = 0.03
    if variant == "sys":
        kw["sys_idx"] = torch.randperm(N, generator=g).to(cuda)
    grads = {}
    for mode in ("fp32", grad_gemm):
        a = rgbd.to(cuda).requires_grad_(True)
        m = mesh.to(cuda).requires_grad_(True)
        loss = matching.circle_match_loss(a, m, labels.to(cuda), match_idx.to(cuda), vis.to(cuda), radius, grad_gemm=mode, **kw)
        loss.backward()
        grads[mode] = (float(loss.detach()), a.grad.clone(), m.grad.clone())
    assert grads["fp32"][0] == grads[grad_gemm][0] and grads["fp32"][0] > 0
    for k in (1, 2):
        ref, got = grads["fp32"][k], grads[grad_gemm][k]
        assert float(ref.abs().max()) > 0
        assert float((got - ref).abs().max()) <= 1e-3 * float(ref.abs().max())


def test_circle_loss_errors(cuda):
    from gadm_b200 import matching, synth, _lib
    rgbd, mesh, _ = synth.descriptors(1, 256, 256, 64, seed=3)
    xyz = synth.fibonacci_sphere(256, 0.2)
    bank = matching.ModelBank(mesh.to(cuda), xyz[None].to(cuda))
    args = (rgbd.to(cuda), bank, torch.ones((1, 256), dtype=torch.long), torch.zeros((1, 256), dtype=torch.long),
            torch.ones((1, 256), dtype=torch.uint8))
    with pytest.raises(_lib.GadmError):                      # 2^logit would leave the fp32 range
        matching.circle_match_loss(*args, 0.01, gamma=40.0)
    with pytest.raises(ValueError):                          # raw features need the model coordinates
        matching.circle_match_loss(args[0], mesh.to(cuda), args[2], args[3], args[4], 0.01)
    # no foreground sample with >= 3 rows: the reference returns 0 (geoMatch.py:151-152)
    assert float(matching.circle_match_loss(args[0], bank, torch.zeros((1, 256), dtype=torch.long), args[3], args[4], 0.01)) == 0.0


def test_geomatch_module_forward_contract(cuda):
    """gadm_b200.matching.GeoMatch keeps the reference's forward contract (models/geoMatch.py:159-200): end_points
    keys and shapes; eval + match_in_forward adds the fused matcher's outputs; training adds 'match_loss' (fused
    CircleLoss, equal to the oracle on the same features), 'seg_loss', 'loss', and gradients reach every head."""
    import torch.nn as nn
    from gadm_b200 import matching, synth
    B, N, M, d, C = 2, 640, 512, 64, 32
    torch.manual_seed(0)

    class Pcd(nn.Module):
        def __init__(self):
            super().__init__()
            self.c = nn.Conv1d(9, C, 1)
        def forward(self, inputs):
            return self.c(inputs['cld_rgb_nrm'])

    class Mesh(nn.Module):
        def __init__(self):
            super().__init__()
            self.f = nn.Parameter(torch.randn(d, M))
        def forward(self):
            return self.f

    xyz = synth.fibonacci_sphere(M, 0.2)
    net = matching.GeoMatchHead(Pcd(), Mesh(), nn.Conv1d(C, d, 1), nn.Conv1d(C, 2, 1), nn.Conv1d(d, C, 1), model_xyz=xyz,
                            match_in_forward=True, positive_r=0.03,
                            seg_loss_func=lambda seg, lab: nn.functional.cross_entropy(seg, lab)).to(cuda)
    g = torch.Generator().manual_seed(1)
    vis = torch.rand((B, M), generator=g) < 0.7
    labels = (torch.rand((B, N), generator=g) < 0.5).long()
    match_idx = torch.full((B, N), M, dtype=torch.int32)
    for b in range(B):
        ids = torch.where(vis[b])[0]
        pick = ids[torch.randint(0, len(ids), (N,), generator=g)]
        match_idx[b] = torch.where(torch.rand((N,), generator=g) < 0.8, pick, torch.full_like(pick, M)).int()
    inputs = {'cld_rgb_nrm': torch.randn((B, 9, N), generator=g).to(cuda), 'labels': labels.to(cuda),
              'match_idx': match_idx.to(cuda), 'visible_flag': vis.to(torch.uint8).to(cuda)}
    net.eval()
    with torch.no_grad():
        ep = net(inputs)
    assert ep['seg'].shape == (B, 2, N) and ep['mesh'].shape == (1, d, M) and ep['rgbd'].shape == (B, d, N)
    assert ep['match_idx'].shape == (B, N) and ep['match_xyz'].shape == (B, N, 3) and 'loss' not in ep
    net.train()
    ep = net(inputs)
    assert {'loss', 'seg_loss', 'match_loss', 'seg', 'mesh', 'rgbd'} <= set(ep) and 'match_idx' not in ep
    bf = synth.bf16_round                                       # the fused loss sees the features rounded to bf16
    want = co.batch_loss(bf(ep['rgbd'].detach().cpu()), bf(ep['mesh'].detach().cpu()[0]), labels, match_idx.long(), xyz,
                         vis, 0.03)
    assert abs(float(ep['match_loss'].detach()) - float(want)) <= TOL * abs(float(want))
    ep['loss'].backward()
    for name, prm in net.named_parameters():
        if name.startswith("normalize_feature_layer"):           # feeds the segmentation branch only
            continue
        assert prm.grad is not None and torch.isfinite(prm.grad).all() and prm.grad.abs().max() > 0, name


def test_geomatch_cfg_constructors_forward(cuda):
    """GeoMatch(cfg, cls_id) (models/geoMatch.py:13-52, :159-200) and the DGCNN variant (geoMatch_DGCNN.py:11-183) with stub
    backbones: end_points keys and shapes ('mesh' [1, d, M] vs [d, M]), training losses equal to the fused loss computed
    directly (incl. the symmetry-aware branch when model_emb.sys_corr_idx is set), eval-mode matching."""
    import torch.nn as nn
    from gadm_b200 import matching, synth
    B, N, M, d = 2, 600, 256, 128
    g = torch.Generator().manual_seed(3)
    xyz = synth.fibonacci_sphere(M, 0.2)

    class Pcd(nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = nn.Conv1d(9, 128, 1)

        def forward(self, x):
            return self.lin(x['cld_rgb_nrm'] if isinstance(x, dict) else x)

    class Mesh(nn.Module):
        sys_corr_idx = None

        def __init__(self):
            super().__init__()
            self.f = nn.Parameter(torch.randn((d, M), generator=g))
            self.register_buffer('xyz', xyz)
            self.register_buffer('mesh', torch.cat([xyz.t(), torch.zeros((3, M))], 0)[None])

        def forward(self):
            return self.f
    cfg = {"feat_dim": d, "neighbor_dis_th": 0.15, "model_d": {1: 200.0}}
    vis = (torch.rand((B, M), generator=g) < 0.7).to(torch.uint8)
    labels = (torch.rand((B, N), generator=g) < 0.5).long()
    match_idx = torch.randint(0, M + 1, (B, N), generator=g).int()
    RT = torch.eye(3, 4)[None].repeat(B, 1, 1)
    RT[:, 2, 3] = 0.9
    inputs = {'cld_rgb_nrm': torch.randn((B, 9, N), generator=g).to(cuda), 'labels': labels.to(cuda),
              'origin_labels': labels.to(cuda), 'match_idx': match_idx.to(cuda), 'visible_flag': vis.to(cuda),
              'RT': RT.to(cuda)}
    for cls, mesh_shape in ((matching.GeoMatch, (1, d, M)), (matching.GeoMatchDGCNN, (d, M))):
        net = cls(cfg, 1, pcd_emb=Pcd(), model_emb=Mesh(), match_in_forward=True).to(cuda)
        net.train()
        ep = net(inputs)
        assert tuple(ep['seg'].shape) == (B, 2, N) and tuple(ep['rgbd'].shape) == (B, d, N)
        assert tuple(ep['mesh'].shape) == mesh_shape
        assert set(('loss', 'seg_loss', 'match_loss')) <= set(ep)
        if cls is matching.GeoMatch:
            want = matching.circle_match_loss(ep['rgbd'], ep['mesh'], inputs['labels'], inputs['match_idx'], inputs['visible_flag'],
                                              net.positive_r, model_xyz=xyz.to(cuda))
        else:
            rad = matching.dgcnn_positive_radius(xyz.to(cuda), inputs['RT'], 3)
            want = matching.circle_match_loss(ep['rgbd'], ep['mesh'][None], inputs['origin_labels'], inputs['match_idx'],
                                              inputs['visible_flag'], rad, model_xyz=xyz.to(cuda), pad_mode="e0")
        assert torch.allclose(ep['match_loss'], want, rtol=1e-6)
        assert torch.allclose(ep['loss'], net.awl(ep['seg_loss'], ep['match_loss']))
        ep['loss'].backward()
        assert net.model_emb.f.grad is not None and net.pcd_emb.lin.weight.grad is not None
        net.eval()
        with torch.no_grad():
            ep = net(inputs)
        assert 'loss' not in ep and tuple(ep['match_idx'].shape) == (B, N) and tuple(ep['match_xyz'].shape) == (B, N, 3)
    # symmetric object: model_emb.sys_corr_idx set -> matching_loss_sys (geoMatch.py:138-141)
    net = matching.GeoMatch(cfg, 1, pcd_emb=Pcd(), model_emb=Mesh()).to(cuda)
    net.model_emb.sys_corr_idx = True
    net.model_emb.sys_idx = torch.randint(0, N, (N,), generator=g).to(cuda)
    net.train()
    ep = net(inputs)
    want = matching.circle_match_loss(ep['rgbd'], ep['mesh'], inputs['labels'], inputs['match_idx'], None, None,
                                      model_xyz=xyz.to(cuda), sys_idx=net.model_emb.sys_idx)
    assert torch.allclose(ep['match_loss'], want, rtol=1e-6)


def test_cal_frame_poses_single_argument(cuda):
    """evaluator.cal_frame_poses(item) with the module-level model container (evaluator.py:28-58, :60-102): the item's own
    mesh_features are used and the model points come from model3ds.models_3d[cls_id]."""
    from gadm_b200 import matching, synth
    N, M, d = 1536, 512, 128
    rgbd, mesh, corr = synth.descriptors(1, N, M, d, regime="planted", seed=8, sigma=0.3)
    xyz = synth.fibonacci_sphere(M, 0.2)
    t = torch.tensor([0.05, 0.02, 0.9])
    cld = (xyz[corr[0]] + t).T.contiguous()
    seg = torch.stack([torch.zeros(N), torch.ones(N)])
    item = (cld.to(cuda), seg.to(cuda), mesh[0].to(cuda), rgbd[0].to(cuda), torch.tensor(7), True)
    matching.set_model_container(None)
    with pytest.raises(RuntimeError):
        matching.cal_frame_poses(item)
    matching.set_model_container(matching.ModelContainer({7: xyz.numpy(), 9: xyz.numpy() * 2}))
    RT = matching.cal_frame_poses(item)
    assert RT.shape == (3, 4) and np.abs(RT[:, :3] - np.eye(3)).max() < 1e-3 and np.abs(RT[:, 3] - t.numpy()).max() < 1e-3
    bank = matching.ModelBank(mesh.to(cuda), xyz[None].to(cuda))
    assert np.array_equal(RT, matching.cal_frame_poses(item, bank))
    assert matching.cal_frame_poses(item[:5] + (False,))[2, 3] == -1000
    matching.set_model_container(None)
