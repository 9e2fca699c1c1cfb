"""Host side of the geoMatch matching head, mirroring the reference interfaces.

  match(...)              the fused matcher (new entry point)
  ModelBank               per-object descriptors + xyz prepared once (the reference keeps `models_3d` and
                          mesh features per class id: evaluator.py:28-58, train_lm.py:331-340)
  cal_frame_poses(item)   drop-in for evaluator.cal_frame_poses (evaluator.py:60-102): same item tuple,
                          same early-outs and sentinel pose, returns the [3,4] pose
  best_fit_transform      Kabsch from GPU-accumulated moments (utils/pvn3d_eval_utils_kpls.py:43-76)
  GeoMatch                nn.Module with the reference's forward signature and end_points keys
                          (models/geoMatch.py:159-200) around pluggable embedding networks
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import MATCH_MODES, OPERAND_MODES, PAD_MODES


class ModelBank:
    """Model-side operands for n_obj objects: bf16 descriptors [n_obj, M, K'] + aux {x,y,z,1/|m|}."""

    def __init__(self, mesh_features, model_xyz, operand_mode="bf16"):
        if mesh_features.dim() == 2:
            mesh_features = mesh_features.unsqueeze(0)
        if model_xyz.dim() == 2:
            model_xyz = model_xyz.unsqueeze(0)
        self.n_obj, self.d, self.M = mesh_features.shape
        self.operand_mode = operand_mode
        self.model_xyz = model_xyz.contiguous().float()
        self.cols, self.aux = ops.prep_model(mesh_features.contiguous().float(), self.model_xyz,
                                             OPERAND_MODES[operand_mode])

    @property
    def device(self):
        return self.cols.device


def _check_obj_id(obj_id, n_obj):
    """obj_id indexes the bank's slots [0, n_obj); the reference's cls_id is 1-based for YCB-V (map it first).  Host
    values are validated here; a device tensor is validated by the caller's data (the kernels read it unclamped)."""
    if isinstance(obj_id, torch.Tensor) and obj_id.is_cuda:
        return
    ids = torch.as_tensor(obj_id).reshape(-1)
    if ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= n_obj):
        raise ValueError(f"obj_id must lie in [0, {n_obj}) (bank slots; a 1-based class id needs cls_id - 1), got "
                         f"[{int(ids.min())}, {int(ids.max())}]")


def match(rgbd, mesh, model_xyz=None, obj_id=None, mask=None, gamma=16.0, pad_mode="none",
          operand_mode="bf16", mode="soft", compact=None):
    """Dense scene-to-model correspondence.

    rgbd [B, d, N] fp32 (end_points['rgbd']); mesh: ModelBank, or [n_obj | 1, d, M] fp32 (end_points['mesh'])
    with model_xyz [n_obj, M, 3]; obj_id int [B] selects the object per frame (bank slot, 0-based); mask [B, N]
    (bool/uint8) marks rows to match.  With a mask the selected rows are COMPACTED on the device first
    (evaluator.py:82-88) and only they are matched, so the cost follows the foreground fraction:
      compact=False / None  results at the rows' original positions, idx = -1 and zeros elsewhere
      compact=True          results in compacted order -- exactly the reference's rgbd_features[cls_msk] ordering --
                            as [B, N] tensors whose first n_sel[b] rows are valid; a fifth value n_sel int32 [B] is
                            returned (no host synchronisation anywhere)
    Returns (idx int64 [B,N], max_sim f32 [B,N], weight f32 [B,N], soft_xyz f32 [B,N,3]); with
    mode="argmax" weight/soft_xyz are None (the reference's hard-argmax path, evaluator.py:89-93);
    mode="argmax_unit" is the same on a bank with operand_mode="bf16n" (columns normalised before the bf16
    rounding), without the per-column scale: fastest, similarities within ~2e-4 of mode="argmax".
    idx == M means the pad column won (pad_mode "minus_one" / "e0")."""
    if rgbd.dim() == 2:
        rgbd = rgbd.unsqueeze(0)
    bank = mesh if isinstance(mesh, ModelBank) else ModelBank(
        mesh, model_xyz if model_xyz is not None else
        torch.zeros((mesh.shape[0] if mesh.dim() == 3 else 1, mesh.shape[-1], 3), device=mesh.device),
        operand_mode)
    if bank.operand_mode != operand_mode:
        raise ValueError(f"bank was prepared with operand_mode={bank.operand_mode!r}")
    if mode == "argmax_unit" and bank.operand_mode != "bf16n":
        raise ValueError("mode='argmax_unit' drops the column scales: it needs a bank prepared with operand_mode='bf16n'")
    B, d, N = rgbd.shape
    if d != bank.d:
        raise ValueError(f"descriptor dim mismatch: scene {d} vs model {bank.d}")
    pm = PAD_MODES[pad_mode]
    if obj_id is not None:
        _check_obj_id(obj_id, bank.n_obj)
        obj_id = torch.as_tensor(obj_id, device=rgbd.device).to(torch.int32).contiguous()
    feat = rgbd.contiguous() if rgbd.dtype == torch.bfloat16 else rgbd.contiguous().float()
    # exact argmax on a bank whose columns were normalised before the rounding: same results, chunk pruning allowed
    kmode = "argmax_bf16n" if mode == "argmax" and bank.operand_mode == "bf16n" else mode
    n_sel = None
    if mask is not None:
        pos, row_map, n_sel = ops.compact_rows(mask.to(torch.uint8).contiguous())
        rows, rinv, pad_sim = ops.prep_rows_sel(feat, pos, OPERAND_MODES[operand_mode], pm)
        idx, max_sim, weight, soft_xyz = ops.match_fwd_sel(rows, rinv, pad_sim, bank.cols, bank.aux, n_sel,
                                                           None if compact else row_map, obj_id, float(gamma), pm,
                                                           MATCH_MODES[kmode])
    else:
        if compact:
            raise ValueError("compact=True needs a mask")
        rows, rinv, pad_sim = ops.prep_rows(feat, OPERAND_MODES[operand_mode], pm)
        idx, max_sim, weight, soft_xyz = ops.match_fwd(rows, rinv, pad_sim, bank.cols, bank.aux, None, obj_id,
                                                       float(gamma), pm, MATCH_MODES[kmode])
    out = (idx, max_sim, None, None) if mode != "soft" else (idx, max_sim, weight, soft_xyz)
    return out + (n_sel,) if compact else out


class _CircleMatchLoss(torch.autograd.Function):
    """Forward: gadm_circle_loss_fwd on bf16 operands prepared from (rgbd, mesh).  Backward: dL/dsim recomputed by
    gadm_circle_loss_bwd (nothing of size [N, M] is kept between the passes), two library GEMMs (G M^ and G^T F^) and
    the backward of the two F.normalize calls; the bf16 rounding of the operands is straight-through."""

    @staticmethod
    def forward(ctx, rgbd, mesh, model_xyz, labels, match_idx, visible_flag, sel, oid, radius, gamma, margin, pad_mode,
                grad_gemm, match_idx2):
        B, d, N = rgbd.shape
        dev = rgbd.device
        rows, rinv, pad_sim = ops.prep_rows(rgbd.contiguous().float(), OPERAND_MODES["bf16"], PAD_MODES[pad_mode])
        cols, aux = ops.prep_model(mesh.contiguous().float(), model_xyz, OPERAND_MODES["bf16"])
        vis = visible_flag.to(dev).bool()                                            # [B, M]
        xyz_f = model_xyz[sel]                                                       # [B, M, 3]
        planes = torch.empty((4, B, xyz_f.shape[1]), dtype=torch.float32, device=dev)
        planes[:3] = torch.where(vis[None], xyz_f.permute(2, 0, 1), xyz_f.new_full((), 1e18))
        planes[3] = radius * radius                                                  # [B, M] squared positive radius
        fg = (labels.to(dev) == 1).to(torch.uint8).contiguous()
        mi = match_idx.to(dev).long().contiguous()
        mi2 = None if match_idx2 is None else match_idx2.to(dev).long().contiguous()
        loss, lse_p, lse_n = ops.circle_loss_fwd(rows, rinv, pad_sim, cols, aux, planes, mi, fg, oid, gamma, margin, mi2)
        cnt = fg.sum(dim=1)
        use = cnt >= 3                                                               # geoMatch.py:128-129
        n_use = use.sum().clamp(min=1)
        row_w = (fg * use[:, None]).float() / (cnt.clamp(min=1)[:, None] * n_use)    # d total / d loss_row
        total = (loss * row_w).sum()
        ctx.save_for_backward(rows, rinv, pad_sim, cols, aux, planes, mi, sel, lse_p, lse_n, row_w)
        ctx.mi2 = mi2
        ctx.oid, ctx.cfg, ctx.n_obj, ctx.grad_gemm = oid, (gamma, margin, pad_mode), mesh.shape[0], grad_gemm
        ctx.mark_non_differentiable(loss, lse_p, lse_n)
        return total, loss, lse_p, lse_n

    @staticmethod
    def backward(ctx, g_total, _g_loss, _g_p, _g_n):
        rows, rinv, pad_sim, cols, aux, planes, mi, sel, lse_p, lse_n, row_w = ctx.saved_tensors
        gamma, margin, pad_mode = ctx.cfg
        B, N, d = rows.shape
        n_obj, M, _ = cols.shape
        w = (torch.sigmoid(lse_p + lse_n) * row_w * g_total).contiguous()            # softplus' = sigmoid
        if ctx.grad_gemm in ("bf16x2", "fused", "flash"):
            return _circle_backward_split(ctx, w, fused={"bf16x2": 0, "fused": 1, "flash": 2}[ctx.grad_gemm])
        G = ops.circle_loss_bwd(rows, rinv, pad_sim, cols, aux, planes, mi, ctx.oid, gamma, margin, lse_p, lse_n,
                                w, ctx.mi2)                                          # [B, N, M + 8]
        f_hat = rows.float() * rinv[..., None]                                       # [B, N, d]
        scale = aux[: n_obj * M].view(n_obj, M, 1)
        m_hat = torch.zeros((n_obj, M + 8, d), dtype=torch.float32, device=rows.device)
        m_hat[:, :M] = cols.float() * scale
        if pad_mode == "minus_one":
            m_hat[:, M] = -(d ** -0.5)                                               # the normalised -1 pad column
        else:
            m_hat[:, M, 0] = 1.0                                                     # e0 (geoMatch_DGCNN.py:95-98)
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = ctx.grad_gemm == "tf32"              # library GEMMs: fp32 unless asked
        try:
            d_fhat = torch.bmm(G, m_hat[sel])                                        # [B, N, d]
            d_mhat_b = torch.bmm(G.transpose(1, 2), f_hat)[:, :M]                    # [B, M, d]
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32
        d_mhat = torch.zeros((n_obj, M, d), dtype=torch.float32, device=rows.device).index_add_(0, sel, d_mhat_b)
        # backward of F.normalize: x^ = x / |x|  =>  dx = (dx^ - (dx^ . x^) x^) / |x|
        d_f = (d_fhat - (d_fhat * f_hat).sum(-1, keepdim=True) * f_hat) * rinv[..., None]
        mh = m_hat[:, :M]
        d_m = (d_mhat - (d_mhat * mh).sum(-1, keepdim=True) * mh) * scale
        return (d_f.transpose(1, 2).contiguous(), d_m.transpose(1, 2).contiguous(), None, None, None, None, None, None,
                None, None, None, None, None, None)


def _circle_backward_split(ctx, w, fused=0):
    """grad_gemm = "bf16x2": dL/dsim leaves the kernel with both norms folded in and split into two bf16 parts
    (gadm_circle_loss_bwd_split), so the two gradient products run as bf16 tensor-core GEMMs with fp32 accumulation
    on the forward pass's own bf16 operands -- exact products, 16 mantissa bits of G (tf32 keeps 10 of G AND rounds the
    other operand) -- with the parts interleaved along K; the norms come back as fp32 row scalings of the results."""
    rows, rinv, pad_sim, cols, aux, planes, mi, sel, lse_p, lse_n, row_w = ctx.saved_tensors
    gamma, margin, pad_mode = ctx.cfg
    B, N, d = rows.shape
    n_obj, M, _ = cols.shape
    Mp = M + 8
    dev = rows.device
    dF = dMb = None
    if fused == 2 and rows.shape[2] <= 128:   # both gradient products inside the kernel: dL/dsim never leaves the SM
        _, g_pad, dF, dMb = ops.circle_loss_bwd_fused(rows, rinv, pad_sim, cols, aux, planes, mi, ctx.oid, gamma, margin,
                                                      lse_p, lse_n, w, ctx.mi2, True)
    elif fused == 1 and rows.shape[2] <= 128:  # the scene-side product inside the kernel, the model-side one a GEMM
        G2, g_pad, dF, _ = ops.circle_loss_bwd_fused(rows, rinv, pad_sim, cols, aux, planes, mi, ctx.oid, gamma, margin,
                                                     lse_p, lse_n, w, ctx.mi2, False)
    else:
        G2, g_pad = ops.circle_loss_bwd_split(rows, rinv, pad_sim, cols, aux, planes, mi, ctx.oid, gamma, margin, lse_p,
                                              lse_n, w, ctx.mi2)                     # [B, N, 2 Mp] bf16, [B, N]
    scale = aux[: n_obj * M].view(n_obj, M, 1)
    if dF is None:
        k = torch.arange(2 * Mp, device=dev)
        col_of_k = (k // 16) * 8 + k % 8                                             # K2 of gadm.h
        cols_p = torch.zeros((n_obj, Mp, d), dtype=torch.bfloat16, device=dev)
        cols_p[:, :M] = cols
        cols2 = cols_p[:, col_of_k]                                                  # [n_obj, 2 Mp, d]
    m_pad = torch.zeros((d,), dtype=torch.float32, device=dev)
    if pad_mode == "minus_one":
        m_pad[:] = -(d ** -0.5)
    else:
        m_pad[0] = 1.0
    if dF is None:
        dF = torch.bmm(G2, cols2[sel], out_dtype=torch.float32)
    d_fhat = dF / rinv[..., None] + g_pad[..., None] * m_pad
    if dMb is None:
        t = torch.bmm(G2.transpose(1, 2), rows, out_dtype=torch.float32)             # [B, 2 Mp, d]
        d_mhat_b = t.view(B, Mp // 8, 2, 8, d).sum(2).reshape(B, Mp, d)[:, :M]       # hi + lo
    else:
        d_mhat_b = dMb[:, :M]
    d_mhat = torch.zeros((n_obj, M, d), dtype=torch.float32, device=dev).index_add_(0, sel, d_mhat_b) / scale
    f_hat = rows.float() * rinv[..., None]
    mh = cols.float() * scale
    d_f = (d_fhat - (d_fhat * f_hat).sum(-1, keepdim=True) * f_hat) * rinv[..., None]
    d_m = (d_mhat - (d_mhat * mh).sum(-1, keepdim=True) * mh) * scale
    return (d_f.transpose(1, 2).contiguous(), d_m.transpose(1, 2).contiguous(), None, None, None, None, None, None,
            None, None, None, None, None, None)


def dgcnn_positive_radius(model_xyz, RT, positive_r):
    """Per-vertex positive radius of the DGCNN variant (models/geoMatch_DGCNN.py:64-65):
    positive_r / 1000 * the camera-space depth of every model vertex.  model_xyz [M, 3], RT [B, 3, 4] -> [B, M]."""
    z = torch.matmul(model_xyz, RT[:, :, :3].transpose(1, 2))[..., 2] + RT[:, None, 2, 3]
    return positive_r / 1000.0 * z


def circle_match_loss(rgbd, mesh, labels, match_idx, visible_flag, positive_r, model_xyz=None, obj_id=None,
                      gamma=16.0, margin=0.2, return_rows=False, pad_mode="minus_one", grad_gemm="fused", sys_idx=None):
    """The matching loss of GeoMatch.pointwise_feature_matching (models/geoMatch.py:102-157 + :55-83 +
    CircleLoss.forward, models/loss.py:475-490) for a whole batch in one fused launch, differentiable with respect
    to rgbd and mesh.

    rgbd [B, d, N] fp32 (end_points['rgbd']); mesh [n_obj | 1, d, M] fp32 (end_points['mesh']) with model_xyz
    [n_obj, M, 3], or a ModelBank (no gradient to the model then); labels [B, N] (rows with label == 1 take part,
    x['labels']); match_idx [B, N] int (ground-truth vertex, M = off the model, x['match_idx']); visible_flag [B, M]
    (x['visible_flag']); positive_r: metres (geoMatch.py:24), a scalar, or a [B, M] tensor of per-vertex radii
    (dgcnn_positive_radius: the DGCNN variant, which also uses pad_mode="e0" and labels = x['origin_labels']).
    grad_gemm: how the two gradient products of the backward pass (G M^ and G^T F^) run.  "fp32": library
    fp32 GEMMs on the fp32 dL/dsim (gradients within ~1e-6 of torch autograd through the reference formulas, 2.8x the
    time of the default); "fused" (default) is described below; "tf32": the same with tf32 allowed (~3x faster backward, gradient error ~5e-4 of
    the largest entry instead of ~1e-6); "bf16x2": dL/dsim leaves the kernel split into two bf16 parts with both norms
    folded in and the products run as bf16 tensor-core GEMMs on the forward pass's own bf16 operands (exact products,
    16 mantissa bits of dL/dsim: more accurate than tf32 and faster than fp32); "fused": as "bf16x2", with the
    scene-side product accumulated inside the kernel by a second MMA per model tile (fastest accurate path); "flash":
    the model-side product inside the kernel as well (third MMA per tile, fp32 reductions into global memory): nothing
    of size [N, M] exists at any time, 10 % slower than "fused".  Both need d <= 128 and fall back to "bf16x2".
    sys_idx (int [>= N], or None): the symmetry-aware variant GeoMatch.matching_loss_sys (models/geoMatch.py:86-100, used
    when model_emb.sys_corr_idx is set, :138-141): the positives of scene point n are exactly the two columns
    match_idx[n] and match_idx[sys_idx[n]] -- no radius, no visibility (positive_r / visible_flag are ignored).
    Returns the scalar the reference returns: the mean over samples with >= 3 foreground rows of the mean row loss
    (0 if there is none); return_rows=True adds the per-row (loss, lse_p, lse_n) tensors."""
    if grad_gemm not in ("fp32", "tf32", "bf16x2", "fused", "flash"):
        raise ValueError("grad_gemm must be 'fp32', 'tf32', 'bf16x2', 'fused' or 'flash'")
    if isinstance(mesh, ModelBank):
        bank = mesh
        if bank.operand_mode != "bf16":
            raise ValueError("circle_match_loss expects a bank prepared with operand_mode='bf16'")
        # a prepared bank holds only the rounded operands: rebuild channel-major fp32 descriptors from them (they are
        # bf16-representable, so the forward pass is unchanged; no gradient reaches the caller's model features)
        model_xyz = bank.model_xyz
        mesh = bank.cols.float().transpose(1, 2).contiguous()
    elif mesh.dim() == 2:
        mesh = mesh.unsqueeze(0)
    if model_xyz is None:
        raise ValueError("model_xyz [n_obj, M, 3] is needed for the positive mask")
    if model_xyz.dim() == 2:
        model_xyz = model_xyz.unsqueeze(0)
    B = rgbd.shape[0]
    dev = rgbd.device
    n_obj = mesh.shape[0]
    oid = None if obj_id is None else torch.as_tensor(obj_id, device=dev).to(torch.int32).contiguous()
    sel = oid.long() if oid is not None else (torch.arange(B, device=dev) if n_obj == B
                                              else torch.zeros(B, dtype=torch.long, device=dev))
    M = mesh.shape[-1]
    mi2 = None
    if sys_idx is not None:                                                   # geoMatch.py:92: match_idx[sys_cor[idxs]]
        N = rgbd.shape[2]
        sidx = torch.as_tensor(sys_idx, device=dev).long()[:N]
        mi2 = match_idx.to(dev).long()[:, sidx]
        positive_r = 0.0 if positive_r is None else positive_r
        if visible_flag is None:
            visible_flag = torch.ones((B, M), dtype=torch.uint8, device=dev)
    radius = (positive_r.to(dev).float() if torch.is_tensor(positive_r) else
              torch.full((1, 1), float(positive_r), device=dev)).expand(B, M)
    total, loss, lse_p, lse_n = _CircleMatchLoss.apply(rgbd, mesh, model_xyz.contiguous().float().to(dev), labels,
                                                       match_idx, visible_flag, sel, oid, radius, float(gamma),
                                                       float(margin), pad_mode, grad_gemm, mi2)
    return (total, loss, lse_p, lse_n) if return_rows else total


def rt_from_moments(mom):
    """best_fit_transform (utils/pvn3d_eval_utils_kpls.py:43-76) from {n, sum A, sum B, sum A B^T}.
    mom: float64 [16] (numpy).  H = AA^T BB = sum(A B^T) - n cA cB^T."""
    n = mom[0]
    ca, cb = mom[1:4] / n, mom[4:7] / n
    H = mom[7:16].reshape(3, 3) - n * np.outer(ca, cb)
    U, S, Vt = np.linalg.svd(H)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        Vt[2, :] *= -1
        R = Vt.T @ U.T
    T = np.zeros((3, 4))
    T[:, :3] = R
    T[:, 3] = cb - R @ ca
    return T


def sentinel_pose():
    """evaluator.py:69-71: identity rotation, t_z = -1000."""
    rt = np.eye(4, dtype=np.float32)
    rt[2, 3] = -1000
    return rt[:3, :]


def frame_poses_device(cld, seg, rgbd, bank, obj_id=None, det=None, min_pts=5, weighted=False, gamma=16.0):
    """Batched evaluator.cal_frame_poses entirely on the device: cld [B,>=3,N], seg [B,2,N], rgbd [B,d,N] ->
    poses float32 [B, 3, 4] (CUDA tensor; no host synchronisation -- the reference syncs per frame at evaluator.py:83,
    :87, :99).  seg argmax -> row compaction -> fused matcher -> moments -> batched 3x3 SVD (gadm_kabsch_poses).
    weighted=True fits a weighted Procrustes with the matcher's softmax weights (the soft-correspondence extension);
    the sentinel rule still counts matched rows."""
    mask = ops.seg_mask(seg.contiguous().float())                            # evaluator.py:78,82 (uint8 [B, N])
    out = match(rgbd, bank, obj_id=obj_id, mask=mask, mode="soft" if weighted else "argmax", gamma=gamma,
                operand_mode=bank.operand_mode)
    cloud = cld[:, :3, :].transpose(1, 2).contiguous().float()               # evaluator.py:85
    oid = None if obj_id is None else torch.as_tensor(obj_id, device=rgbd.device).to(torch.int32)
    mom = ops.kabsch_moments(out[0], mask, cloud, bank.aux, oid, bank.M, bank.n_obj)
    d8 = None if det is None else torch.as_tensor(det, device=rgbd.device).to(torch.uint8).contiguous()
    if weighted:
        momw = ops.kabsch_moments_w(out[0], mask, out[2].contiguous(), cloud, bank.aux, oid, bank.M, bank.n_obj)
        return ops.kabsch_poses(momw, mom, d8, min_pts)
    return ops.kabsch_poses(mom, None, d8, min_pts)


def frame_poses(cld, seg, rgbd, bank, obj_id=None, det=None, min_pts=5, weighted=False):
    """Batched evaluator.cal_frame_poses: -> list of numpy [3,4] poses (what the reference's callers consume).
    Everything runs on the device (frame_poses_device); ONE device->host copy of B x 12 floats ends the batch."""
    poses = frame_poses_device(cld, seg, rgbd, bank, obj_id=obj_id, det=det, min_pts=min_pts, weighted=weighted)
    return list(poses.cpu().numpy())


class ModelContainer:
    """The evaluator's module-level model store (evaluator.py:28-58): model point sets per object id, in metres
    (np.load(...)[:MODEL_PT_NUM, :3] / 1000 there), plus the optional symmetric correspondence indices.  The
    reference builds it from its dataset config at import time (`model3ds = ModelContainer(ycbv_cfg)`); here the
    caller hands over the arrays (mesh files are outside this package): set_model_container(ModelContainer({...}))."""

    def __init__(self, models_3d, sys_corr_idx=None, feat_dim=None):
        self.models_3d = {int(k): np.asarray(v, dtype=np.float32)[:, :3] for k, v in models_3d.items()}
        self.sys_corr_idx = {} if sys_corr_idx is None else {str(k): v for k, v in sys_corr_idx.items()}
        self.feat_dim = feat_dim
        self._dev = {}

    def xyz(self, cls_id, device):
        key = (int(cls_id), str(device))
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(self.models_3d[int(cls_id)]).to(device)
        return self._dev[key]


model3ds = None          # evaluator.py:58


def set_model_container(container):
    global model3ds
    model3ds = container


def cal_frame_poses(item, bank=None):
    """Drop-in for evaluator.cal_frame_poses(item) (evaluator.py:60-102); item =
    (cld [>=3,N], seg_features [2,N], mesh_features [d,M], rgbd_features [d,N], cls_id, det) -> numpy [3,4].

    With ONE argument it behaves as the reference does: the model points come from the module-level container
    (`model3ds.models_3d[cls_id]`, evaluator.py:99; install it with set_model_container) and the model descriptors are
    the item's own mesh_features.  With a prepared ModelBank (second argument) the item's mesh_features are ignored and
    cls_id selects the bank slot when the bank holds more than one object."""
    cld, seg_features, mesh_features, rgbd_features, cls_id, det = item
    if not det:
        return sentinel_pose()
    cid = int(cls_id.item()) if torch.is_tensor(cls_id) else int(cls_id)
    if bank is None:
        if model3ds is None:
            raise RuntimeError("cal_frame_poses(item) needs the module-level model container "
                               "(matching.set_model_container(ModelContainer({cls_id: xyz})), evaluator.py:28-58)")
        xyz = model3ds.xyz(cid, rgbd_features.device)
        M = xyz.shape[0]
        bank = ModelBank(mesh_features[:, :M].contiguous(), xyz)          # evaluator.py:90 normalises mesh_features
        oid = None
    else:
        oid = None if bank.n_obj == 1 else [cid]
    return frame_poses(cld[None], seg_features[None], rgbd_features[None], bank, obj_id=oid, det=[det])[0]


class GeoMatchHead(nn.Module):
    """The reference's GeoMatch forward contract (models/geoMatch.py:159-200, geoMatch_DGCNN.py:138-183) around
    pluggable embedding networks (the FFB6D / SplineCNN / DGCNN backbones are out of scope).

    pcd_emb(inputs) -> [B, C_emb, N]; model_emb() -> [d, M].  Heads keep the reference's layer names
    (seg_layer, feature_encoding_layer, normalize_feature_layer) so checkpoints load unchanged when the
    same head modules are supplied.  forward(inputs, end_points=None) returns end_points with
    'seg' [B,2,N], 'mesh' [1,d,M], 'rgbd' [B,d,N]; in eval mode with match_in_forward=True it adds
    'match_idx', 'match_sim', 'match_weight', 'match_xyz' from the fused kernel.  In training mode (geoMatch.py:188-195)
    it adds 'match_loss' (the fused CircleLoss, circle_match_loss) when positive_r and model_xyz are given, and
    'seg_loss' / 'loss' through the pluggable seg_loss_func(seg, labels) / awl(seg_loss, match_loss) (the reference's
    FocalLoss and AutomaticWeightedLoss are outside the path); without awl, 'loss' is their plain sum."""

    def __init__(self, pcd_emb, model_emb, feature_encoding_layer, seg_layer, normalize_feature_layer=None,
                 model_xyz=None, match_in_forward=False, gamma=16.0, operand_mode="bf16", positive_r=None,
                 seg_loss_func=None, awl=None):
        super().__init__()
        self.positive_r, self.seg_loss_func, self.awl = positive_r, seg_loss_func, awl
        self.pcd_emb, self.model_emb = pcd_emb, model_emb
        self.feature_encoding_layer, self.seg_layer = feature_encoding_layer, seg_layer
        self.normalize_feature_layer = normalize_feature_layer
        self.match_in_forward, self.gamma, self.operand_mode = match_in_forward, gamma, operand_mode
        if model_xyz is not None:
            self.register_buffer("xyz", model_xyz.float())
        else:
            self.xyz = None

    def forward(self, inputs, end_points=None):
        if not end_points:
            end_points = {}
        rgbd_emb = self.pcd_emb(inputs)                                   # geoMatch.py:178
        mesh_features = self.model_emb()                                  # :179
        rgbd_features = self.feature_encoding_layer(rgbd_emb)             # :180
        if self.normalize_feature_layer is not None:
            rgbd_emb = rgbd_emb + self.normalize_feature_layer(rgbd_features)   # :181-182
        seg_features = self.seg_layer(rgbd_emb)                           # :183
        mesh_features = mesh_features.unsqueeze(0)                        # :184
        if self.training and self.positive_r is not None and self.xyz is not None and 'match_idx' in inputs:
            match_loss = circle_match_loss(rgbd_features, mesh_features, inputs['labels'], inputs['match_idx'],
                                           inputs['visible_flag'], self.positive_r, model_xyz=self.xyz,
                                           gamma=16.0, margin=0.2)        # :190 (CircleLoss(16), m = 0.2: :27, :81)
            end_points['match_loss'] = match_loss
            if self.seg_loss_func is not None:
                seg_loss = self.seg_loss_func(seg_features, inputs['labels'])               # :191
                end_points['seg_loss'] = seg_loss
                end_points['loss'] = self.awl(seg_loss, match_loss) if self.awl is not None else seg_loss + match_loss
            else:
                end_points['loss'] = match_loss
        end_points['seg'] = seg_features
        end_points['mesh'] = mesh_features
        end_points['rgbd'] = rgbd_features
        if self.match_in_forward and not self.training:
            xyz = self.xyz if self.xyz is not None else None
            mask = ops.seg_mask(seg_features.detach().contiguous().float())
            idx, sim, w, sxyz = match(rgbd_features.detach(), mesh_features.detach(), xyz, mask=mask,
                                      gamma=self.gamma, operand_mode=self.operand_mode)
            end_points.update(match_idx=idx, match_sim=sim, match_weight=w, match_xyz=sxyz)
        return end_points


# ------------------------------------------------------------------------------------------------ reference constructors
class _BN1d(nn.Sequential):
    """models/pytorch_utils.py:42-54 (BatchNorm1d wrapper: child 'bn', weight 1, bias 0)."""

    def __init__(self, n):
        super().__init__()
        self.add_module("bn", nn.BatchNorm1d(n))
        nn.init.constant_(self[0].weight, 1.0)
        nn.init.constant_(self[0].bias, 0)


def _conv1d(in_size, out_size, bn=False, activation=True, bias=True):
    """models/pytorch_utils.py:70-160 (Conv1d = _ConvBase): children 'conv' [, 'normlayer'] [, 'activation'], kernel 1,
    kaiming-normal weight, zero bias, no bias next to a batch norm -- the same module names, so the reference's
    checkpoints load into it."""
    m = nn.Sequential()
    conv = nn.Conv1d(in_size, out_size, kernel_size=1, bias=bias and not bn)
    nn.init.kaiming_normal_(conv.weight)
    if conv.bias is not None:
        nn.init.constant_(conv.bias, 0)
    m.add_module("conv", conv)
    if bn:
        m.add_module("normlayer", _BN1d(out_size))
    if activation:
        m.add_module("activation", nn.ReLU(inplace=True))
    return m


def _seq(in_channels, layers):
    """models/pytorch_utils.py:272-316 (Seq(...).conv1d(...)...): children '0', '1', ..."""
    m, c = nn.Sequential(), in_channels
    for i, (out, kw) in enumerate(layers):
        m.add_module(str(i), _conv1d(c, out, **kw))
        c = out
    return m


class FocalLoss(nn.Module):
    """models/loss.py:15-46 with alpha=None, size_average=True (the reference uses FocalLoss(gamma=2),
    geoMatch.py:29): mean of -(1 - p_t)^gamma log p_t with p_t detached."""

    def __init__(self, gamma=0):
        super().__init__()
        self.gamma = gamma

    def forward(self, input, target):
        logp = F.log_softmax(input.transpose(1, 2).reshape(-1, input.size(1)), dim=-1)
        logpt = logp.gather(1, target.reshape(-1, 1).long()).view(-1)
        pt = logpt.detach().exp()
        return (-1 * (1 - pt) ** self.gamma * logpt).mean()


class AutomaticWeightedLoss(nn.Module):
    """models/loss.py:496-516: sum_i 0.5 / p_i^2 * loss_i + log(1 + p_i^2) with learnable p (ones)."""

    def __init__(self, num=2):
        super().__init__()
        self.params = nn.Parameter(torch.ones(num))

    def forward(self, *x):
        return sum(0.5 / (self.params[i] ** 2) * loss + torch.log(1 + self.params[i] ** 2) for i, loss in enumerate(x))


class GeoMatch(GeoMatchHead):
    """GeoMatch(cfg, cls_id) -- the reference's constructor (models/geoMatch.py:13-52) and forward contract
    (:159-200).  cfg['feat_dim'], cfg['neighbor_dis_th'], cfg['model_d'][cls_id] are read as there; the heads
    (seg_layer, feature_encoding_layer, normalize_feature_layer) are built with the reference's layer stack and module
    names, so its state_dict keys match.  The two backbones are outside this package: pass them as
    pcd_emb= / model_emb= (or cfg['pcd_emb'] / cfg['model_emb']); model_emb() -> [d, M], with the model points in
    model_emb._buffers['xyz'] and, for symmetric objects, model_emb.sys_corr_idx / model_emb.sys_idx as in the reference
    (the symmetry-aware loss of :138-141 is then used)."""

    def __init__(self, cfg, cls_id, pcd_emb=None, model_emb=None, match_in_forward=False, gamma=16.0,
                 operand_mode="bf16"):
        feat_dim = cfg["feat_dim"]
        positive_r = cfg["neighbor_dis_th"] * cfg["model_d"][cls_id] / 1000.0            # geoMatch.py:24
        pcd_emb = pcd_emb if pcd_emb is not None else cfg.get("pcd_emb")
        model_emb = model_emb if model_emb is not None else cfg.get("model_emb")
        if pcd_emb is None or model_emb is None:
            raise NotImplementedError("the FFB6D / SplineCNN embedding networks are outside this package: pass "
                                      "pcd_emb= and model_emb= (or cfg['pcd_emb'] / cfg['model_emb'])")
        bn = dict(bn=True)
        seg_layer = _seq(feat_dim, [(128, bn), (128, bn), (128, bn), (2, dict(activation=False))])          # :33-39
        feature_encoding_layer = _seq(self._enc_in(feat_dim), [(128, bn), (128, bn), (128, bn),
                                                               (feat_dim, dict(activation=False, bias=False))])   # :40-46
        normalize_feature_layer = _conv1d(feat_dim, feat_dim, bn=True)                                      # :48-51
        xyz = getattr(model_emb, "_buffers", {}).get("xyz")
        super().__init__(pcd_emb, model_emb, feature_encoding_layer, seg_layer, normalize_feature_layer,
                         model_xyz=None, match_in_forward=match_in_forward, gamma=gamma, operand_mode=operand_mode,
                         positive_r=positive_r, seg_loss_func=FocalLoss(gamma=2), awl=AutomaticWeightedLoss(2))
        self.feat_dim, self.cls_id = feat_dim, cls_id
        self._xyz_from_emb = xyz is not None

    @staticmethod
    def _enc_in(feat_dim):
        return 128                                                                      # pt_utils.Seq(128), :41

    def _model_xyz(self):
        return self.model_emb._buffers["xyz"].contiguous() if self._xyz_from_emb else self.xyz       # :148

    def _pcd(self, inputs):
        return self.pcd_emb(inputs)                                                     # :178

    def _mesh_out(self, mesh_features):
        return mesh_features.unsqueeze(0)                                               # :184  'mesh' [1, d, M]

    def _match_loss(self, rgbd_features, mesh_features, inputs):
        sys_idx = self.model_emb.sys_idx if getattr(self.model_emb, "sys_corr_idx", None) is not None else None
        return circle_match_loss(rgbd_features, mesh_features.reshape(1, *mesh_features.shape[-2:]), inputs['labels'],
                                 inputs['match_idx'], inputs.get('visible_flag'), self.positive_r,
                                 model_xyz=self._model_xyz(), gamma=16.0, margin=0.2, sys_idx=sys_idx)   # :27, :81, :98

    def forward(self, inputs, end_points=None):
        if not end_points:
            end_points = {}
        rgbd_emb = self._pcd(inputs)
        mesh_features = self.model_emb()                                                # :179
        rgbd_features = self.feature_encoding_layer(rgbd_emb)                           # :180
        rgbd_emb = rgbd_emb + self.normalize_feature_layer(rgbd_features)               # :181-182
        seg_features = self.seg_layer(rgbd_emb)                                         # :183
        mesh_features = self._mesh_out(mesh_features)
        if self.training:                                                               # :188-195
            match_loss = self._match_loss(rgbd_features, mesh_features, inputs)
            seg_loss = self.seg_loss_func(seg_features, inputs[self._label_key])
            end_points['loss'] = self.awl(seg_loss, match_loss)
            end_points['seg_loss'] = seg_loss
            end_points['match_loss'] = match_loss
        end_points['seg'] = seg_features
        end_points['mesh'] = mesh_features
        end_points['rgbd'] = rgbd_features
        if self.match_in_forward and not self.training:
            mask = ops.seg_mask(seg_features.detach().contiguous().float())
            idx, sim, w, sxyz = match(rgbd_features.detach(), mesh_features.detach().reshape(1, *mesh_features.shape[-2:]),
                                      self._model_xyz(), mask=mask, gamma=self.gamma, operand_mode=self.operand_mode)
            end_points.update(match_idx=idx, match_sim=sim, match_weight=w, match_xyz=sxyz)
        return end_points

    _label_key = 'labels'


class GeoMatchDGCNN(GeoMatch):
    """models/geoMatch_DGCNN.py:11-183: the same head on DGCNN embeddings.  Differences kept: positive_r = 3 (a
    per-vertex radius positive_r / 1000 * camera depth, :22, :64-65), feature_encoding_layer starts from feat_dim
    (:38), the point embedding takes inputs['cld_rgb_nrm'] (:157-159), the model points are channels 0..2 of
    model_emb._buffers['mesh'] (:111), the pad column is e0 (:95-98), rows are picked by x['origin_labels'] (:107),
    and 'mesh' is returned as model_emb() gives it -- NOT unsqueezed (:160, :180)."""

    def __init__(self, cfg, cls_id, pcd_emb=None, model_emb=None, match_in_forward=False, gamma=16.0,
                 operand_mode="bf16"):
        cfg = dict(cfg, neighbor_dis_th=0.0, model_d={cls_id: 0.0})
        super().__init__(cfg, cls_id, pcd_emb, model_emb, match_in_forward, gamma, operand_mode)
        self.positive_r = 3                                                             # :22

    @staticmethod
    def _enc_in(feat_dim):
        return feat_dim                                                                 # pt_utils.Seq(self.feat_dim), :38

    def _model_xyz(self):
        return self.model_emb._buffers['mesh'][0, :3].t().contiguous()                  # :111

    def _pcd(self, inputs):
        return self.pcd_emb(inputs['cld_rgb_nrm'])                                      # :157-159

    def _mesh_out(self, mesh_features):
        return mesh_features                                                            # :160, :180

    def _match_loss(self, rgbd_features, mesh_features, inputs):
        xyz = self._model_xyz()
        radius = dgcnn_positive_radius(xyz, inputs['RT'].to(xyz.device).float(), self.positive_r)    # :64-65
        return circle_match_loss(rgbd_features, mesh_features.reshape(1, *mesh_features.shape[-2:]),
                                 inputs['origin_labels'], inputs['match_idx'], inputs['visible_flag'], radius,
                                 model_xyz=xyz, gamma=16.0, margin=0.2, pad_mode="e0")

    _label_key = 'labels'
