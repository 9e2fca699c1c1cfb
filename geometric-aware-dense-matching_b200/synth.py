"""Synthetic workloads in the shapes BASELINE.json names.  Deterministic by seed, built on
the CPU with torch/numpy; no dataset or checkpoint is needed.  Descriptors are rounded to bf16 once so the
oracle and the kernel consume the same bf16-representable values (DESIGN.md "Precision contract")."""
import math

import numpy as np
import torch

# LINEMOD intrinsics (reference: common.py:161-163)
LM_K = np.array([[572.4114, 0.0, 325.2611], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]])
# LM-O object diameters in metres (reference: config/lmo_cfg.py:6-22, mm there)
LMO_DIAMETERS = [0.10210, 0.24750, 0.16736, 0.17249, 0.20141, 0.15455, 0.12426, 0.26148]
# config/ycbv_cfg.py:2-24 (mm -> m), objects 1..21
YCBV_DIAMETERS = [0.172063, 0.269573, 0.198377, 0.120543, 0.196463, 0.089797, 0.142543, 0.114053, 0.129540, 0.197796,
                  0.259534, 0.259566, 0.161922, 0.124990, 0.226170, 0.237299, 0.203973, 0.121365, 0.174746, 0.217094,
                  0.102903]


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def fibonacci_sphere(M, diameter, dtype=torch.float32):
    """M quasi-uniform points on a sphere of the given diameter -> [M, 3]."""
    i = torch.arange(M, dtype=torch.float64) + 0.5
    phi = torch.acos(1 - 2 * i / M)
    theta = math.pi * (1 + 5 ** 0.5) * i
    r = diameter / 2
    return torch.stack([r * torch.cos(theta) * torch.sin(phi), r * torch.sin(theta) * torch.sin(phi),
                        r * torch.cos(phi)], dim=1).to(dtype)


def descriptors(B, N, M, d, n_obj=1, regime="random", seed=2000, sigma=0.5):
    """-> rgbd [B, d, N], mesh [n_obj, d, M] fp32 (bf16-representable), corr [B, N] (planted index or -1).
    regime "random": iid N(0,1); "planted": f_i = m_{c(i)} + sigma * eps with c(i) uniform over M."""
    g = torch.Generator().manual_seed(seed)
    mesh = bf16_round(torch.randn((n_obj, d, M), generator=g))
    if regime == "random":
        rgbd = bf16_round(torch.randn((B, d, N), generator=g))
        corr = torch.full((B, N), -1, dtype=torch.int64)
    elif regime == "planted":
        corr = torch.randint(0, M, (B, N), generator=g)
        rgbd = torch.empty((B, d, N))
        for b in range(B):
            src = mesh[b % n_obj]                                   # object b % n_obj
            rgbd[b] = src[:, corr[b]] + sigma * torch.randn((d, N), generator=g)
        rgbd = bf16_round(rgbd)
    else:
        raise ValueError(regime)
    return rgbd, mesh, corr


def model_bank_xyz(n_obj, M, diameters=LMO_DIAMETERS):
    return torch.stack([fibonacci_sphere(M, diameters[o % len(diameters)]) for o in range(n_obj)], dim=0)


def depth_cloud(in_size=128, n_points=12800, seed=1000, dup_frac=0.0, invalid_frac=0.0):
    """A crop of an RGB-D frame as the reference's dataset produces it (datasets/lm/linemod_pbr.py:485-527):
    smooth random depth surface z in [0.4, 1.5] m + N(0, 2 mm), back-projected with LINEMOD K on an
    in_size x in_size pixel grid; n_points pixels chosen by shuffled mask (wrap-padding duplicates when
    dup_frac > 0); image-grid clouds sr2dptxyz[s] = every s-th pixel.
    -> cld [n_points, 3] fp32, {1,2,4,8: [(in_size/s)^2, 3]} fp32 (numpy)."""
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[:in_size, :in_size].astype(np.float64)
    u0, v0 = 260.0, 180.0                                   # crop origin inside the 640x480 frame
    z = np.full((in_size, in_size), 0.0)
    for _ in range(4):                                      # a few random low-frequency waves
        fx, fy = rng.uniform(0.5, 3.0, 2) * 2 * np.pi / in_size
        z += rng.uniform(0.05, 0.15) * np.sin(fx * xs + rng.uniform(0, 6.28)) * np.cos(fy * ys + rng.uniform(0, 6.28))
    z = 0.95 + z
    z = np.clip(z + rng.normal(0, 0.002, z.shape), 0.4, 1.5)
    if invalid_frac > 0:                                    # holes in the depth map -> (0,0,0) points
        z = np.where(rng.random(z.shape) < invalid_frac, 0.0, z)
    X = (xs + u0 - LM_K[0, 2]) * z / LM_K[0, 0]
    Y = (ys + v0 - LM_K[1, 2]) * z / LM_K[1, 1]
    xyz = np.stack([X, Y, z], axis=-1).astype(np.float32)   # [h, w, 3]
    valid = np.flatnonzero(xyz[..., 2].reshape(-1) > 1e-8)
    n_unique = int(round(n_points * (1 - dup_frac)))
    choose = valid.copy()
    if len(choose) > n_unique:
        m = np.zeros(len(choose), dtype=int); m[:n_unique] = 1
        rng.shuffle(m)
        choose = choose[m.nonzero()]
    if len(choose) < n_points:
        choose = np.pad(choose, (0, n_points - len(choose)), "wrap")    # linemod_pbr.py:492
    rng.shuffle(choose)
    cld = xyz.reshape(-1, 3)[choose]
    sr = {}
    for s in (1, 2, 4, 8):
        n = in_size // s
        yy, xx = np.mgrid[:n, :n]
        sr[s] = xyz[yy * s, xx * s].reshape(-1, 3).copy()
    return cld, sr


def frame_batch(B, in_size=128, n_points=12800, seed=1000, dup_frac=0.0):
    """B frames -> cld [B, N, 3], {s: [B, P_s, 3]} torch fp32."""
    clds, srs = [], {1: [], 2: [], 4: [], 8: []}
    for b in range(B):
        c, sr = depth_cloud(in_size, n_points, seed + b, dup_frac)
        clds.append(torch.from_numpy(c))
        for s in srs:
            srs[s].append(torch.from_numpy(sr[s]))
    return torch.stack(clds), {s: torch.stack(v) for s, v in srs.items()}
