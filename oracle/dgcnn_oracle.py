"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the DGCNN dynamic-graph kNN and edge-feature gather.

Restates /root/reference/models/dgcnn.py:21-27 (knn) and :30-56 (get_graph_feature) with the same
torch operations (so the fp32 rounding of the distance form  -|xi|^2 + 2 xi.xj - |xj|^2  is the
reference's), minus the hard-coded torch.device('cuda') at :39.  tests/golden/dgcnn_*.npz were produced
by importing the reference module itself (tests/golden/make_golden.py) and pin this restatement.
"""
import torch


def pairwise_neg_sqdist(x):
    """x [B, C, N] -> [B, N, N]; dgcnn.py:22-24 (note the order of the three terms)."""
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    return -xx - inner - xx.transpose(2, 1)


def knn(x, k):
    """x [B, C, N] -> idx int64 [B, N, k], nearest first, self included (dgcnn.py:26)."""
    return pairwise_neg_sqdist(x).topk(k=k, dim=-1)[1]


def knn_with_gaps(x, k):
    """idx plus, per row, the minimum gap between adjacent ranks 1..k+1 of the distance values --
    the parity gate for feature-space kNN is exact only where that gap exceeds the fp32 noise."""
    pd = pairwise_neg_sqdist(x)
    vals, idx = pd.topk(k=min(k + 1, pd.shape[-1]), dim=-1)
    gaps = (vals[..., :-1] - vals[..., 1:]).min(dim=-1).values
    return idx[..., :k], gaps, vals


def get_graph_feature(x, k=20, idx=None, dim9=False):
    """x [B, C, N] -> [B, 2C, N, k] = cat(neighbour - centre, centre).  dgcnn.py:30-56."""
    B, C, N = x.shape
    if idx is None:
        idx = knn(x[:, :3], k) if dim9 else knn(x, k)             # :35-38
    rows = x.transpose(2, 1).contiguous()                          # [B, N, C]   :48
    flat = (idx + torch.arange(B).view(-1, 1, 1) * N).view(-1)     # :41-45
    nbr = rows.view(B * N, C)[flat].view(B, N, k, C)               # :49-50
    ctr = rows.view(B, N, 1, C).expand(B, N, k, C)                 # :51
    return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2).contiguous()   # :53
