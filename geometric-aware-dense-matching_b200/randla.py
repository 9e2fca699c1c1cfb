"""Drop-ins for the RandLA-Net consumers of the kNN index tensors, each one fused kernel
instead of a materialised torch.gather + a second pass:

  random_sample(feature, pool_idx)          models/RandLA/RandLANet.py:90-105    gather + max over the k neighbours
  nearest_interpolation(feature, idx)       models/RandLA/RandLANet.py:107-120   1-NN gather
  relative_pos_encoding(xyz, neigh_idx)     models/RandLA/RandLANet.py:720-727   [|p-q|, p-q, p, q] per (point, nbr)
  gather_neighbour(pc, neighbor_idx)        models/RandLA/RandLANet.py:729-738

Differentiable in the features / coordinates like the reference's torch.gather compositions (scatter-add backward
kernels, registered as autograd formulas of the custom ops); the index tensors carry no gradient."""
from . import ops
from .pointops import gather_neighbour  # noqa: F401


def random_sample(feature, pool_idx):
    """feature [B, d, N, 1] (or [B, d, N]), pool_idx [B, N', k] -> [B, d, N', 1]."""
    f = feature.squeeze(3) if feature.dim() == 4 else feature
    out = ops.gather_max(f.contiguous().float(), pool_idx.contiguous().long())
    return out.unsqueeze(3)


def nearest_interpolation(feature, interp_idx):
    """feature [B, C, N, 1] (or [B, C, N]), interp_idx [B, up, 1] -> [B, C, up, 1]."""
    f = feature.squeeze(3) if feature.dim() == 4 else feature
    idx = interp_idx.reshape(interp_idx.shape[0], interp_idx.shape[1], 1)
    return ops.gather_max(f.contiguous().float(), idx.contiguous().long()).unsqueeze(3)


def relative_pos_encoding(xyz, neigh_idx):
    """xyz [B, N, 3], neigh_idx [B, N, k] -> [B, N, k, 10]."""
    return ops.relative_pos_encoding(xyz.contiguous().float(), neigh_idx.contiguous().long())
