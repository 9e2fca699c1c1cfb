"""Geometric 3-D kNN behind the reference's interfaces.

  DataProcessing.knn_search(support_pts, query_pts, k)   models/RandLA/helper_tool.py:161-170
        numpy [B,N1,3], [B,N2,3] -> np.int32 [B,N2,k]   (host buffers in, host buffers out)
  knn_search_cuda(...)                                   same on CUDA tensors, no host round trip
  KnnPyramid                                             the 22 kNN calls one sample needs
        (datasets/lm/linemod_pbr.py:534-569) for a batch of frames in ONE library call
Ties are ordered by ascending index (nanoflann orders them by tree traversal)."""
import numpy as np
import torch

from . import ops


def knn_search_cuda(support_pts, query_pts, k, algo="auto"):
    """support [B,N1,3], query [B,N2,3] CUDA fp32 -> int32 [B,N2,k]."""
    return ops.knn3d(support_pts.contiguous().float(), query_pts.contiguous().float(), int(k), ops.KNN_ALGOS[algo])


class DataProcessing:
    @staticmethod
    def knn_search(support_pts, query_pts, k, device=None):
        """:param support_pts: points you have, B*N1*3
        :param query_pts: points you want to know the neighbour index, B*N2*3
        :param k: Number of neighbours in knn search
        :return: neighbor_idx: neighboring points indexes, B*N2*k   (np.int32, helper_tool.py:170)"""
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        s = torch.from_numpy(np.ascontiguousarray(support_pts, dtype=np.float32)).to(dev, non_blocking=True)
        q = torch.from_numpy(np.ascontiguousarray(query_pts, dtype=np.float32)).to(dev, non_blocking=True)
        return knn_search_cuda(s, q, k).cpu().numpy()


class KnnPyramid:
    """All kNN index tensors of a batch of samples, as the reference's dataset builds them one call at a time
    (linemod_pbr.py:528-569): 4 down-sampling levels x {self k=16, up 1-NN, r2p k=16, p2r 1-NN} + 3 up-sampling
    levels x {r2p k=16, p2r 1-NN} = 22 calls, here one job table -> one gadm_knn3d call.

    cld [B, N, 3]; sr2dptxyz {1|2|4|8: [B, (in_size/s)^2, 3]} (only 2, 4, 8 are used).
    __call__ returns {name: int32 [B, n_query, k]} with the reference's key names."""

    RGB_DS_SR = [4, 8, 8, 8]      # linemod_pbr.py:528
    RGB_UP_SR = [4, 2, 2]         # linemod_pbr.py:557

    def __init__(self, n_points, grid_sizes, batch, k_nei=16, algo="auto", device=None):
        self.N, self.B, self.k, self.algo = n_points, batch, k_nei, algo
        self.grid_sizes = dict(grid_sizes)          # {2: P2, 4: P4, 8: P8}
        # flat point buffer layout per batch item: [cld (N) | sr2 | sr4 | sr8]
        self.off = {"cld": 0}
        o = n_points
        for s in (2, 4, 8):
            self.off[s] = o
            o += self.grid_sizes[s]
        self.P = o
        calls, out = [], 0
        n_lvl = [n_points // (4 ** i) for i in range(5)]
        def add(name, s_off, ns, q_off, nq, k):
            nonlocal out
            calls.append((name, (s_off, q_off, out, self.P, self.P, 0, ns, nq, k, batch), nq, k))
            out += batch * nq * k
        for i in range(4):
            sr = self.RGB_DS_SR[i]
            add("cld_nei_idx%d" % i, 0, n_lvl[i], 0, n_lvl[i], k_nei)                    # :534-536
            add("cld_interp_idx%d" % i, 0, n_lvl[i + 1], 0, n_lvl[i], 1)                  # :539-541
            add("r2p_ds_nei_idx%d" % i, self.off[sr], self.grid_sizes[sr], 0, n_lvl[i + 1], k_nei)   # :546-548
            add("p2r_ds_nei_idx%d" % i, 0, n_lvl[i + 1], self.off[sr], self.grid_sizes[sr], 1)       # :550-552
        for i in range(3):
            sr = self.RGB_UP_SR[i]
            lvl = n_lvl[4 - i - 1]
            add("r2p_up_nei_idx%d" % i, self.off[sr], self.grid_sizes[sr], 0, lvl, k_nei)            # :559-562
            add("p2r_up_nei_idx%d" % i, 0, lvl, self.off[sr], self.grid_sizes[sr], 1)                # :564-567
        # out_bstride = nq * k for every job (outputs of one job are [B, nq, k] contiguous)
        fixed = []
        for name, j, nq, k in calls:
            j = list(j); j[5] = nq * k
            fixed.append(tuple(j))
        self.names = [(c[0], c[1][2], c[2], c[3]) for c in calls]
        self.jobs = ops.make_jobs(fixed)
        self.out_elems = out
        self.workspace = None
        self.n_queries = sum(c[2] for c in calls)
        self.pairs_brute = sum(c[1][6] * c[2] for c in calls)
        # algorithmic bytes per frame: 12 N_s + 12 N_q + 4 k N_q per call
        self.algorithmic_bytes = sum(12 * c[1][6] + 12 * c[2] + 4 * c[3] * c[2] for c in calls)

    def pack(self, cld, sr2dptxyz):
        """-> flat [B * P, 3] point buffer in the layout the job table indexes."""
        parts = [cld] + [sr2dptxyz[s] for s in (2, 4, 8)]
        return torch.cat([p.float() for p in parts], dim=1).contiguous().view(-1, 3)

    def run_packed(self, pts, out=None):
        lib_ws = self.workspace
        idx = ops.knn3d_jobs(pts, pts, self.jobs, self.out_elems, self.algo, workspace=lib_ws, out=out)
        return idx

    def __call__(self, cld, sr2dptxyz):
        pts = self.pack(cld, sr2dptxyz)
        flat = self.run_packed(pts)
        return self.unpack(flat)

    def unpack(self, flat):
        """flat index buffer (int32, or the uint16 transport form of pipeline.FrameStream) -> the reference's dict of
        int32 [B, nq, k] arrays (datasets/lm/linemod_pbr.py:534-569 keys)."""
        if flat.dtype != torch.int32:
            flat = flat.to(torch.int32)
        out = {}
        for name, off, nq, k in self.names:
            out[name] = flat[off: off + self.B * nq * k].view(self.B, nq, k)
        # linemod_pbr.py:538 / :543: the pooling indices are the first N/4 rows of the self-kNN
        for i in range(4):
            nei = out["cld_nei_idx%d" % i]
            out["cld_sub_idx%d" % i] = nei[:, : nei.shape[1] // 4, :]
        return out
