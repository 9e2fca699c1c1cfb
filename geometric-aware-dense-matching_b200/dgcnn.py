"""Drop-ins for models/dgcnn.py:21-56 (knn, get_graph_feature) on the fused CUDA kernels.  get_graph_feature is
differentiable in x (as the reference's gather / cat), the neighbour indices carry no gradient."""
from . import ops


def knn(x, k):
    """x [B, C, N] fp32 CUDA -> idx int64 [B, N, k]; k nearest in feature space, self first
    (models/dgcnn.py:21-27; the [B,N,N] distance matrix is never materialised)."""
    x = x.contiguous().float()
    return ops.knn_feat(x, int(k), x.shape[1])


def get_graph_feature(x, k=20, idx=None, dim9=False):
    """x [B, C, N] -> [B, 2C, N, k] = cat(neighbour - centre, centre)  (models/dgcnn.py:30-56)."""
    B, N = x.size(0), x.size(2)
    x = x.view(B, -1, N).contiguous().float()
    if idx is None:
        idx = ops.knn_feat(x, int(k), 3 if dim9 else x.shape[1])          # dgcnn.py:35-38
    return ops.graph_feature(x, idx.contiguous())
