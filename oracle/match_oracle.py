"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the geoMatch matching head (torch fp32, CPU).

Follows, line for line in meaning:
  * live inference matcher      /root/reference/evaluator.py:77-93   (seg argmax :78, transpose :79,
    mask :82-88, F.normalize rows :89, F.normalize columns :90, matmul :91, torch.max :93)
  * padded variant (-1 column)  /root/reference/utils/pvn3d_eval_utils_kpls.py:436-444 and
    models/geoMatch.py:117-119  (pad, THEN normalise columns; "idx != M" marks a match)
  * padded variant (e0 column)  /root/reference/models/geoMatch_DGCNN.py:92-99
  * soft correspondence: NOT in the reference.  Defined here as
        w = softmax(gamma * S[:, :M], dim=1); weight = w.max(1); soft_xyz = w @ model_xyz
    with gamma defaulting to 16 (the only temperature in the reference: CircleLoss(16),
    models/geoMatch.py:27).  The pad column never enters the softmax.
F.normalize uses eps=1e-12; torch.max / argmax on CPU return the FIRST maximal index.
"""
import torch
import torch.nn.functional as F

PAD_NONE, PAD_MINUS_ONE, PAD_E0 = "none", "minus_one", "e0"


def _padded_mesh(mesh_features: torch.Tensor, pad_mode: str) -> torch.Tensor:
    d = mesh_features.shape[0]
    if pad_mode == PAD_NONE:
        return mesh_features
    if pad_mode == PAD_MINUS_ONE:                      # geoMatch.py:117-118
        pad = -torch.ones((d, 1), dtype=mesh_features.dtype)
    elif pad_mode == PAD_E0:                           # geoMatch_DGCNN.py:95-97
        pad = torch.zeros((d, 1), dtype=mesh_features.dtype)
        pad[0] = 1
    else:
        raise ValueError(pad_mode)
    return torch.cat([mesh_features, pad], dim=1)


def similarity(rgbd_features, mesh_features, row_mask=None, pad_mode=PAD_NONE):
    """rgbd_features [d, N], mesh_features [d, M] fp32 -> S [n_sel, M(+1)] fp32."""
    rows = rgbd_features.transpose(0, 1)               # evaluator.py:79
    if row_mask is not None:
        rows = rows[row_mask]                          # evaluator.py:88
    rows = F.normalize(rows, p=2, dim=1)               # evaluator.py:89
    cols = F.normalize(_padded_mesh(mesh_features, pad_mode), p=2, dim=0)   # evaluator.py:90
    return torch.matmul(rows, cols)                    # evaluator.py:91


def seg_mask(seg_features):
    """seg_features [2, N] -> bool [N]  (evaluator.py:78,82)."""
    return torch.argmax(seg_features, dim=0) == 1


def match_ref(rgbd_features, mesh_features, row_mask=None, pad_mode=PAD_NONE):
    """Exactly evaluator.py:89-93 and nothing else: normalize, normalize, matmul, torch.max -> (idx, max_sim).
    This is what the CPU arms of bench.py time (match_hard below adds a top-2 for the parity margin, which the
    reference does not run)."""
    S = similarity(rgbd_features, mesh_features, row_mask, pad_mode)
    max_sim, idx = torch.max(S, dim=1)                 # evaluator.py:93
    return idx, max_sim


def match_hard(rgbd_features, mesh_features, row_mask=None, pad_mode=PAD_NONE):
    """-> (idx int64 [n_sel], max_sim fp32 [n_sel], margin fp32 [n_sel]).  evaluator.py:93.
    margin = top1 - top2 of the row, used by the parity gate (exact where margin > 1e-3)."""
    S = similarity(rgbd_features, mesh_features, row_mask, pad_mode)
    max_sim, idx = torch.max(S, dim=1)
    top2 = torch.topk(S, 2, dim=1).values
    return idx, max_sim, top2[:, 0] - top2[:, 1]


def match_soft(rgbd_features, mesh_features, model_xyz, gamma=16.0, row_mask=None, pad_mode=PAD_NONE,
               dtype=torch.float32):
    """Extension oracle.  model_xyz [M, 3].  -> dict(idx, max_sim, margin, weight, soft_xyz).
    dtype=torch.float64 gives the fp64 shadow used for error budgeting."""
    M = mesh_features.shape[1]
    S = similarity(rgbd_features.to(dtype), mesh_features.to(dtype), row_mask, pad_mode)
    max_sim, idx = torch.max(S, dim=1)
    top2 = torch.topk(S, 2, dim=1).values
    w = torch.softmax(gamma * S[:, :M], dim=1)
    return dict(idx=idx, max_sim=max_sim, margin=top2[:, 0] - top2[:, 1],
                weight=w.max(dim=1).values, soft_xyz=w @ model_xyz.to(dtype))


def best_fit_transform(A, B):
    """Kabsch, restating /root/reference/utils/pvn3d_eval_utils_kpls.py:43-76 in torch fp64.
    A [n,3] model points, B [n,3] camera points -> T [3,4] with B ~ R A + t."""
    A = A.double(); B = B.double()
    ca, cb = A.mean(0), B.mean(0)                      # :58-59
    H = (A - ca).T @ (B - cb)                          # :60-63
    U, S, Vt = torch.linalg.svd(H)                     # :64
    R = Vt.T @ U.T                                     # :65
    if torch.linalg.det(R) < 0:                        # :67-69
        Vt = Vt.clone(); Vt[2, :] *= -1
        R = Vt.T @ U.T
    t = cb - R @ ca                                    # :71
    return torch.cat([R, t[:, None]], dim=1)


def best_fit_transform_weighted(A, B, w):
    """Extension (weights = the soft-correspondence weights): weighted Procrustes, same steps as best_fit_transform with
    weighted centroids and H = sum w (A - ca)(B - cb)^T."""
    A = A.double(); B = B.double(); w = w.double()
    W = w.sum()
    ca, cb = (w[:, None] * A).sum(0) / W, (w[:, None] * B).sum(0) / W
    H = ((A - ca) * w[:, None]).T @ (B - cb)
    U, S, Vt = torch.linalg.svd(H)
    R = Vt.T @ U.T
    if torch.linalg.det(R) < 0:
        Vt = Vt.clone(); Vt[2, :] *= -1
        R = Vt.T @ U.T
    t = cb - R @ ca
    return torch.cat([R, t[:, None]], dim=1)
