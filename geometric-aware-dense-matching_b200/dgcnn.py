"""Drop-ins for models/dgcnn.py:21-56 (knn, get_graph_feature) on the fused CUDA kernels.  get_graph_feature is
differentiable in x (as the reference's gather / cat), the neighbour indices carry no gradient."""
from . import ops


def knn(x, k):
    """x [B, C, N] fp32 CUDA -> idx int64 [B, N, k]; k nearest in feature space, self first
    (models/dgcnn.py:21-27; the [B,N,N] distance matrix is never materialised)."""
    x = x.contiguous().float()
    return ops.knn_feat(x, int(k), x.shape[1])


def knn_xyz(x, k):
    """dim9 branch of get_graph_feature (models/dgcnn.py:38): kNN over the first three channels.  Three channels are a
    point cloud: the exact grid search of the 3-D kNN (gadm_knn3d, d2 = dx^2 + dy^2 + dz^2, ties by index) finds them
    in a tenth of the time of the feature-space kernel.  Its metric is the direct form of the reference's
    -|xi|^2 + 2 xi.xj - |xj|^2; the two orderings differ only where the reference's own fp32 cancellation noise
    (~1e-7 |x|^2) decides (same gates as knn_feat in tests/test_gpu_dgcnn_pointops.py).  k <= 32."""
    pts = x[:, :3].transpose(1, 2).contiguous().float()
    return ops.knn3d(pts, pts, int(k), ops.KNN_ALGOS["auto"]).long()


def get_graph_feature(x, k=20, idx=None, dim9=False):
    """x [B, C, N] -> [B, 2C, N, k] = cat(neighbour - centre, centre)  (models/dgcnn.py:30-56)."""
    B, N = x.size(0), x.size(2)
    x = x.view(B, -1, N).contiguous().float()
    if idx is None:
        idx = knn_xyz(x, k) if dim9 else ops.knn_feat(x, int(k), x.shape[1])          # dgcnn.py:35-38
    return ops.graph_feature(x, idx.contiguous())
