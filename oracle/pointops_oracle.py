"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the two pointops entry points the path names.

The CUDA sources of lib/pointops are absent from the reference (lib/pointops/pointops.egg-info/SOURCES.txt:6-24) and
nothing imports the package, but its own pure-torch oracle ships: this restatement is PINNED by
tests/golden/pointops_golden.npz, produced by EXECUTING KNNQueryNaive.forward
(/root/reference/lib/pointops/functions/pointops.py:396-426) and QueryAndGroup.forward (:548-585) from the reference
source text (tests/golden/make_golden.py::make_pointops).  Grouping follows the docstring contract :151-155 and the
scatter-add of its backward :166-176.
torch.sort is not stable across equal keys by contract; ties are canonicalised by (dist, index) here (the fixture is
tie-free among the first k + 1 distances, asserted by the test).
"""
import torch


def knnquery_naive(nsample, xyz, new_xyz=None):
    """xyz (b,n,3), new_xyz (b,m,3) -> idx int32 (b,m,nsample); pointops.py:405-424.
    dist = (new - xyz)^2 summed over the coordinate axis (torch reduction order: x, y, z)."""
    if new_xyz is None:
        new_xyz = xyz
    diff = new_xyz[:, :, None, :] - xyz[:, None, :, :]
    dist = diff.pow(2).sum(dim=3)
    idxs = torch.sort(dist, dim=2, stable=True)[1]      # stable => ties by ascending index
    return idxs[:, :, :nsample].int(), torch.sort(dist, dim=2, stable=True)[0][:, :, :nsample]


def query_and_group(nsample, xyz, new_xyz=None, features=None, use_xyz=True):
    """QueryAndGroup.forward, kNN branch (pointops.py:548-585) -> (new_features, grouped_xyz, idx int64)."""
    if new_xyz is None:
        new_xyz = xyz
    idx = knnquery_naive(nsample, xyz, new_xyz)[0]
    grouped_xyz = grouping(xyz.transpose(1, 2).contiguous(), idx)
    diff = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
    if features is not None:
        gf = grouping(features, idx)
        new = torch.cat([diff, gf], dim=1) if use_xyz else gf
    else:
        new = diff
    return new, grouped_xyz, idx.long()


def grouping(features, idx):
    """features (b,c,n), idx (b,m,s) int -> (b,c,m,s): out[b,c,m,s] = features[b,c,idx[b,m,s]]."""
    b, c, n = features.shape
    _, m, s = idx.shape
    gather_idx = idx.long().view(b, 1, m * s).expand(b, c, m * s)
    return torch.gather(features, 2, gather_idx).view(b, c, m, s)


def grouping_backward(grad_out, idx, n):
    """grad_out (b,c,m,s) -> grad_features (b,c,n): scatter-add (pointops.py:166-176)."""
    b, c, m, s = grad_out.shape
    g = torch.zeros(b, c, n, dtype=grad_out.dtype)
    gather_idx = idx.long().view(b, 1, m * s).expand(b, c, m * s)
    return g.scatter_add_(2, gather_idx, grad_out.reshape(b, c, m * s))
