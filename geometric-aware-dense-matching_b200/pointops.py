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
    """lib/pointops/functions/pointops.py:536-585 with the reference's constructor and return values: kNN grouping
    (radius=None; the ball query is not on this path), forward -> (new_features (b, 3+c | c | 3, m, nsample),
    grouped_xyz (b, 3, m, nsample)) and, with return_idx=True, the int64 neighbour indices as a third value."""

    def __init__(self, radius=None, nsample=32, use_xyz=True, return_idx=False):
        super().__init__()
        if radius is not None:
            raise NotImplementedError("ball query (radius) is outside the kNN path this package replaces")
        self.radius, self.nsample, self.use_xyz, self.return_idx = radius, nsample, use_xyz, return_idx

    def forward(self, xyz, new_xyz=None, features=None, idx=None):
        if new_xyz is None:
            new_xyz = xyz
        if idx is None:
            idx = knnquery_heap(self.nsample, xyz, new_xyz)                      # :566
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping(xyz_trans, idx)                                   # :568  (b, 3, m, nsample)
        grouped_xyz_diff = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)   # :570
        if features is not None:
            grouped_features = grouping(features.contiguous(), idx)              # :572
            new_features = torch.cat([grouped_xyz_diff, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz_diff
        if self.return_idx:
            return new_features, grouped_xyz, idx.long()
        return new_features, grouped_xyz
