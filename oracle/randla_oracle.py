"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the RandLA-Net consumers of the kNN indices, line by line from
/root/reference/models/RandLA/RandLANet.py (plain torch, imports nothing from the reference at run time).
Pinned by tests/golden/randla_golden.npz, which tests/golden/make_golden.py produces by EXECUTING the reference's own
functions (Network.random_sample / nearest_interpolation are staticmethods; relative_pos_encoding and
gather_neighbour are called on the reference's Building_block class without constructing it)."""
import torch


def random_sample(feature, pool_idx):
    """RandLANet.py:90-105."""
    feature = feature.squeeze(dim=3)
    num_neigh = pool_idx.shape[-1]
    d = feature.shape[1]
    batch_size = pool_idx.shape[0]
    pool_idx = pool_idx.reshape(batch_size, -1)
    pool_features = torch.gather(feature, 2, pool_idx.unsqueeze(1).repeat(1, feature.shape[1], 1))
    pool_features = pool_features.reshape(batch_size, d, -1, num_neigh)
    return pool_features.max(dim=3, keepdim=True)[0]


def nearest_interpolation(feature, interp_idx):
    """RandLANet.py:107-120."""
    feature = feature.squeeze(dim=3)
    batch_size = interp_idx.shape[0]
    up_num_points = interp_idx.shape[1]
    interp_idx = interp_idx.reshape(batch_size, up_num_points)
    interpolated = torch.gather(feature, 2, interp_idx.unsqueeze(1).repeat(1, feature.shape[1], 1))
    return interpolated.unsqueeze(3)


def gather_neighbour(pc, neighbor_idx):
    """RandLANet.py:729-738."""
    batch_size, num_points, d = pc.shape
    index_input = neighbor_idx.reshape(batch_size, -1)
    features = torch.gather(pc, 1, index_input.unsqueeze(-1).repeat(1, 1, pc.shape[2])).contiguous()
    return features.reshape(batch_size, num_points, neighbor_idx.shape[-1], d)


def relative_pos_encoding(xyz, neigh_idx):
    """RandLANet.py:720-727."""
    neighbor_xyz = gather_neighbour(xyz, neigh_idx)
    xyz_tile = xyz.unsqueeze(2).repeat(1, 1, neigh_idx.shape[-1], 1)
    relative_xyz = xyz_tile - neighbor_xyz
    relative_dis = torch.sqrt(torch.sum(torch.pow(relative_xyz, 2), dim=-1, keepdim=True))
    return torch.cat([relative_dis, relative_xyz, xyz_tile, neighbor_xyz], dim=-1)
