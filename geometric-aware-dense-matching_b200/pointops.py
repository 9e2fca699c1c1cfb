"""Drop-ins for the two pointops entry points on the path (lib/pointops/functions/pointops.py):
knnquery / knnquery_heap (:435-493) and grouping (:149-178), plus RandLA's gather_neighbour
(models/RandLA/RandLANet.py:729-738)."""
import torch

from . import ops


def knnquery(nsample, xyz, new_xyz=None):
    """xyz (b,n,3), new_xyz (b,m,3) [defaults to xyz] -> idx int32 (b,m,nsample), non-differentiable."""
    if new_xyz is None:
        new_xyz = xyz
    with torch.no_grad():
        return ops.knn3d(xyz.contiguous().float(), new_xyz.contiguous().float(), int(nsample), ops.KNN_ALGOS["auto"])


knnquery_heap = knnquery


def grouping(features, idx):
    """features (b,c,n) fp32, idx (b,m,nsample) int32 -> (b,c,m,nsample); differentiable w.r.t. features."""
    assert features.is_contiguous() and idx.is_contiguous()
    return ops.group_fwd(features, idx)


def gather_neighbour(pc, neighbor_idx):
    """pc (B,N,C), neighbor_idx (B,M,K) int64 -> (B,M,K,C)."""
    return ops.gather_neighbour(pc.contiguous().float(), neighbor_idx.contiguous().long())


class QueryAndGroup(torch.nn.Module):
    """kNN grouping (pointops.py:536-585, the nsample/knn branch): returns (b, 3+c, m, nsample)."""

    def __init__(self, nsample=32, use_xyz=True):
        super().__init__()
        self.nsample, self.use_xyz = nsample, use_xyz

    def forward(self, xyz, new_xyz=None, features=None, idx=None):
        if new_xyz is None:
            new_xyz = xyz
        if idx is None:
            idx = knnquery(self.nsample, xyz, new_xyz)                       # :566
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping(xyz_trans, idx)                               # :568
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)    # :570
        if features is None:
            return grouped_xyz
        grouped_features = grouping(features.contiguous(), idx)              # :572
        return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
