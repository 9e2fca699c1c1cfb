"""TEST INFRASTRUCTURE (oracle) -- never imported by the product package.

CPU restatement (torch, fp32) of the training-side matching loss of the reference:
  /root/reference/models/geoMatch.py:102-157   GeoMatch.pointwise_feature_matching (per-sample loop, -1 pad column)
  /root/reference/models/geoMatch.py:55-83     GeoMatch.matching_loss (positive mask from the model-space radius)
  /root/reference/utils/basic_utils.py:86-89   pdist (sqrt(sum((A - B)^2) + 1e-7))
  /root/reference/models/loss.py:441-490       CircleLoss.log_sum_exp / forward
Pinned by tests/golden/circle_golden.npz, which tests/golden/make_golden.py produces by EXECUTING those reference
lines (CircleLoss imported from models/loss.py, the two geoMatch methods and pdist run from the source text with
.cuda() made a no-op)."""
import torch
import torch.nn.functional as F


def pdist(A, B):
    """basic_utils.py:86-89, dist_type 'L2'."""
    D2 = torch.sum((A.unsqueeze(1) - B.unsqueeze(0)).pow(2), 2)
    return torch.sqrt(D2 + 1e-7)


def positive_mask(match_idx, mesh_xyz, vis_flag, positive_r):
    """geoMatch.py:55-79.  match_idx [n] (M = off the model), mesh_xyz [M, 3], vis_flag [M] -> bool [n, M + 1].
    positive_r: a scalar (geoMatch.py:67) or a per-vertex tensor [M] (geoMatch_DGCNN.py:64-65: positive_r / 1000 * the
    camera-space depth of the vertex; only the entries of visible vertices are used)."""
    n_node = len(mesh_xyz)
    in_mesh = match_idx != n_node                                               # :59
    vis = vis_flag.to(torch.bool)
    mask = torch.zeros((len(match_idx), n_node), dtype=torch.bool)              # :72
    if in_mesh.any():
        gt_pt = mesh_xyz[match_idx[in_mesh]]                                    # :63
        r = positive_r[vis] if torch.is_tensor(positive_r) and positive_r.dim() else positive_r
        near = pdist(gt_pt, mesh_xyz[vis]) < r                                  # :65-66, :73
        sub = torch.zeros((int(in_mesh.sum()), n_node), dtype=torch.bool)
        sub[:, vis] = near                                                      # :74-75
        mask[in_mesh] = sub                                                     # :76
    return torch.cat([mask, (~in_mesh).unsqueeze(1)], dim=1)                    # :78 (pad column)


def _masked_lse(logit, mask):
    """loss.py:441-461 with a {0,1} mask: the LSE over the masked-in entries, -inf when there is none."""
    neg = torch.full_like(logit, float("-inf"))
    return torch.logsumexp(torch.where(mask, logit, neg), dim=-1)


def circle_rows(sim, mask, m=0.2, gamma=16.0):
    """loss.py:475-490 without the final mean: per-row softplus(LSE_p + LSE_n), and the two LSEs."""
    ap = torch.clamp_min(-sim.detach() + 1 + m, min=0.0)                         # :479 (masking = restricting the LSE)
    an = torch.clamp_min(sim.detach() + m, min=0.0)                              # :480 (both detached, as there)
    logit_p = -ap * (sim - (1 - m)) * gamma                                      # :482, :488
    logit_n = an * (sim - m) * gamma                                             # :483, :489
    lse_p, lse_n = _masked_lse(logit_p, mask), _masked_lse(logit_n, ~mask)       # :491-492
    return F.softplus(lse_p + lse_n), lse_p, lse_n                               # :494


def dgcnn_radius(mesh_xyz, RT, positive_r):
    """geoMatch_DGCNN.py:64-65 for all vertices: positive_r / 1000 * (R x + t)_z.  RT [3, 4] -> [M]."""
    proj = torch.matmul(mesh_xyz, RT[:, :3].t()) + RT[:, 3:].t()
    return positive_r / 1000.0 * proj[:, 2]


def sample_rows(rgbd_feature, mesh_feature, labels, match_idx, mesh_xyz, vis_flag, positive_r, m=0.2, gamma=16.0,
                pad="minus_one"):
    """One iteration of the loop at geoMatch.py:125-149 (pad "minus_one") or geoMatch_DGCNN.py:107-127 (pad "e0";
    there all rows are normalised first, which is the same arithmetic per row).  rgbd_feature [d, N], mesh_feature
    [d, M], labels [N], match_idx [N] -> (idxs, loss_rows, lse_p, lse_n) over the foreground rows."""
    d = rgbd_feature.shape[0]
    padding = -torch.ones((d, 1))                                                # geoMatch.py:117
    if pad == "e0":                                                              # geoMatch_DGCNN.py:95-96
        padding = torch.zeros((d, 1))
        padding[0] = 1
    mesh_padded = F.normalize(torch.cat([mesh_feature, padding], dim=1), p=2, dim=0)   # :117-119 / DGCNN :97-98
    idxs = torch.where(labels == 1)[0]                                           # :127
    selected = F.normalize(rgbd_feature.transpose(0, 1).index_select(0, idxs), p=2, dim=1)         # :131, :134
    sim = torch.matmul(selected, mesh_padded)                                    # :136
    mask = positive_mask(match_idx.index_select(0, idxs).long(), mesh_xyz, vis_flag, positive_r)   # :143-149
    return (idxs,) + circle_rows(sim, mask, m, gamma)


def batch_loss(rgbd, mesh_feature, labels, match_idx, mesh_xyz, vis_flags, positive_r, m=0.2, gamma=16.0,
               pad="minus_one"):
    """geoMatch.py:102-157 (geoMatch_DGCNN.py:80-136 with pad "e0" and positive_r a [B, M] tensor): mean over the
    samples with >= 3 foreground rows of the mean row loss; 0 if none."""
    per = []
    for i in range(rgbd.shape[0]):
        if int((labels[i] == 1).sum()) < 3:                                      # :128-129
            continue
        r = positive_r[i] if torch.is_tensor(positive_r) and positive_r.dim() == 2 else positive_r
        _, rows, _, _ = sample_rows(rgbd[i], mesh_feature, labels[i], match_idx[i], mesh_xyz, vis_flags[i], r, m, gamma,
                                    pad)
        per.append(rows.mean())
    return torch.stack(per).mean() if per else torch.tensor(0.0)                 # :151-156


def sys_positive_mask(match_idx, idxs, sys_idx, M):
    """GeoMatch.matching_loss_sys (models/geoMatch.py:86-100): for the selected rows idxs, the positives are the columns
    match_idx[idxs] and match_idx[sys_idx[idxs]] of the [len(idxs), M + 1] similarity."""
    n = len(idxs)
    rows = torch.arange(n)
    rows = torch.cat((rows, rows), dim=0)                                        # :91-92
    cols = torch.cat((match_idx[idxs], match_idx[sys_idx[idxs]]), dim=0).long()  # :93
    mask = torch.zeros((n, M + 1), dtype=torch.bool)
    mask.index_put_((rows, cols), torch.tensor(True))                            # :95-96
    return mask


def batch_loss_sys(rgbd, mesh_feature, labels, match_idx, sys_idx, m=0.2, gamma=16.0):
    """pointwise_feature_matching (models/geoMatch.py:102-157) on the sys_corr_idx branch (:138-141)."""
    import torch.nn.functional as F
    d, M = mesh_feature.shape[-2:]
    mesh = mesh_feature.reshape(d, M)
    padded = F.normalize(torch.cat([mesh, -torch.ones((d, 1))], dim=1), p=2, dim=0)   # :117-119
    losses = []
    for b in range(rgbd.shape[0]):
        idxs = torch.where(labels[b] == 1)[0]                                    # :127
        if len(idxs) < 3:
            continue
        sim = F.normalize(rgbd[b].t()[idxs], p=2, dim=1) @ padded                # :131-136
        mask = sys_positive_mask(match_idx[b].long(), idxs, sys_idx.long(), M)
        losses.append(circle_rows(sim, mask, m, gamma)[0].mean())
    return torch.stack(losses).mean() if losses else torch.tensor(0.0)
