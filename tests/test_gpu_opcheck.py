"""-m gpu: torch.library.opcheck over every torch.ops.gadm.* op -- the schema (mutation / aliasing annotations), the
fake (meta) implementation against the shapes and dtypes the CUDA op really returns, and the autograd registration of
the differentiable gathers."""
import pytest
import torch

pytestmark = pytest.mark.gpu

BASIC = ("test_schema", "test_faketensor")
AUTOGRAD = BASIC + ("test_autograd_registration",)


def _cases(dev):
    from gadm_b200 import ops
    g = torch.Generator().manual_seed(77)
    B, d, N, M, n_obj = 2, 64, 256, 128, 3
    feat = torch.randn((B, d, N), generator=g).to(dev)
    mesh = torch.randn((n_obj, d, M), generator=g).to(dev)
    xyz = torch.randn((n_obj, M, 3), generator=g).to(dev)
    mask = (torch.rand((B, N), generator=g) < 0.4).to(torch.uint8).to(dev)
    obj = torch.tensor([2, 0], dtype=torch.int32, device=dev)
    cloud = torch.randn((B, N, 3), generator=g).to(dev)
    rows, rinv, pad = ops.prep_rows(feat, 0, 0)
    rows_p, rinv_p, pad_p = ops.prep_rows(feat, 0, 1)
    cols, aux = ops.prep_model(mesh, xyz, 0)
    pos, row_map, n_sel = ops.compact_rows(mask)
    idx, max_sim, weight, soft_xyz = ops.match_fwd(rows, rinv, pad, cols, aux, None, obj, 16.0, 0, 1)
    mom = ops.kabsch_moments(idx, mask, cloud, aux, obj, M, n_obj)
    det = torch.ones((B,), dtype=torch.uint8, device=dev)
    x = torch.randn((B, 16, N), generator=g).to(dev)
    kidx = torch.randint(0, N, (B, N, 8), generator=g).to(dev)
    kidx_m = torch.randint(0, N, (B, 64, 8), generator=g).to(dev)
    pc = torch.randn((B, N, 5), generator=g).to(dev)
    seg = torch.randn((B, 2, N), generator=g).to(dev)
    gidx = torch.randint(0, N, (B, 64, 8), generator=g).to(torch.int32).to(dev)

    planes = torch.empty((4, B, M), device=dev)
    planes[:3] = xyz[obj.long()].permute(2, 0, 1)
    planes[3] = 0.05
    mi = torch.randint(0, M + 1, (B, N), generator=g).to(dev)
    closs, lse_p, lse_n = ops.circle_loss_fwd(rows_p, rinv_p, pad_p, cols, aux, planes, mi, mask, obj, 16.0, 0.25)

    def grad(t):
        return t.clone().requires_grad_(True)

    return [
        ("prep_rows", (feat, 0, 0), BASIC),
        ("prep_rows", (feat, 1, 1), BASIC),
        ("prep_rows", (feat.to(torch.bfloat16), 0, 0), BASIC),
        ("prep_model", (mesh, xyz, 0), BASIC),
        ("prep_model", (mesh, xyz, 1), BASIC),
        ("compact_rows", (mask,), BASIC),
        ("prep_rows_sel", (feat, pos, 0, 1), BASIC),
        ("match_fwd", (rows, rinv, pad, cols, aux, None, obj, 16.0, 0, 0), BASIC),
        ("match_fwd", (rows_p, rinv_p, pad_p, cols, aux, mask, obj, 16.0, 1, 1), BASIC),
        ("match_fwd_sel", (rows, rinv, pad, cols, aux, n_sel, row_map, obj, 16.0, 0, 1), BASIC),
        ("match_fwd_sel", (rows, rinv, pad, cols, aux, n_sel, None, obj, 16.0, 0, 0), BASIC),
        ("pack_match_outputs", (idx, max_sim, weight, soft_xyz,
                                torch.empty((B, N, 6), dtype=torch.int32, device=dev)), BASIC),
        ("pack_indices_u16", (torch.randint(0, 60000, (1001,), generator=g).to(torch.int32).to(dev),
                              torch.empty((1001,), dtype=torch.uint16, device=dev)), BASIC),
        ("kabsch_moments", (idx, mask, cloud, aux, obj, M, n_obj), BASIC),
        ("kabsch_moments_w", (idx, None, weight, cloud, aux, obj, M, n_obj), BASIC),
        ("kabsch_poses", (mom, mom, det, 4), BASIC),
        ("knn3d", (cloud, cloud[:, :100].contiguous(), 4, 0), BASIC),
        ("knn_feat", (x, 8, 16), BASIC),
        ("graph_feature", (grad(x), kidx), AUTOGRAD),
        ("graph_feature_bwd", (torch.randn((B, 32, N, 8), device=dev), kidx), BASIC),
        ("group_fwd", (grad(x), gidx), AUTOGRAD),
        ("group_bwd", (torch.randn((B, 16, 64, 8), device=dev), gidx, N), BASIC),
        ("gather_neighbour", (grad(pc), kidx_m), AUTOGRAD),
        ("gather_neighbour_bwd", (torch.randn((B, 64, 8, 5), device=dev), kidx_m, N), BASIC),
        ("gather_max", (grad(x), kidx_m), AUTOGRAD),
        ("gather_max_bwd", (x, kidx_m, torch.randn((B, 16, 64), device=dev)), BASIC),
        ("relative_pos_encoding", (grad(cloud), kidx), AUTOGRAD),
        ("seg_mask", (seg,), BASIC),
        ("circle_loss_fwd", (rows_p, rinv_p, pad_p, cols, aux, planes, mi, mask, obj, 16.0, 0.25), BASIC),
        ("circle_loss_fwd", (rows_p, rinv_p, pad_p, cols, aux, planes, mi, None, obj, 16.0, 0.25, mi), BASIC),
        ("circle_loss_bwd", (rows_p, rinv_p, pad_p, cols, aux, planes, mi, obj, 16.0, 0.25, lse_p, lse_n,
                             torch.ones((B, N), device=dev)), BASIC),
        ("circle_loss_bwd_split", (rows_p, rinv_p, pad_p, cols, aux, planes, mi, obj, 16.0, 0.25, lse_p, lse_n,
                                   torch.ones((B, N), device=dev)), BASIC),
        ("circle_loss_bwd_fused", (rows_p, rinv_p, pad_p, cols, aux, planes, mi, obj, 16.0, 0.25, lse_p, lse_n,
                                   torch.ones((B, N), device=dev)), BASIC),
        ("circle_loss_bwd_fused", (rows_p, rinv_p, pad_p, cols, aux, planes, mi, obj, 16.0, 0.25, lse_p, lse_n,
                                   torch.ones((B, N), device=dev), None, True), BASIC),
    ]


def test_opcheck_every_op(cuda):
    import gadm_b200  # noqa: F401  (registers torch.ops.gadm.*)
    seen = set()
    for name, args, utils in _cases(cuda):
        op = getattr(torch.ops.gadm, name).default
        torch.library.opcheck(op, args, test_utils=utils)
        seen.add(name)
    from gadm_b200 import ops
    registered = {v._qualname.split("::")[1] for v in vars(ops).values() if isinstance(v, torch.library.CustomOpDef)}
    assert registered == seen, registered ^ seen
