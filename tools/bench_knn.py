"""kNN-pyramid-only timing (CUDA events) at the BASELINE shape: 8 frames x the 22-call schedule."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gadm_b200  # noqa
from gadm_b200 import ops, synth
from gadm_b200.knn import KnnPyramid

dev = torch.device("cuda", 0)
B, N = 8, 12800
cld, sr = synth.frame_batch(B, 128, N, seed=2000)
pyr = KnnPyramid(N, {s: (128 // s) ** 2 for s in (2, 4, 8)}, B)
ws = ops._lib.load().gadm_knn3d_workspace_bytes(pyr.jobs, len(pyr.jobs), ops.KNN_ALGOS["auto"])
pyr.workspace = torch.empty((max(ws, 16),), dtype=torch.uint8, device=dev)
pts = pyr.pack(cld.to(dev), {s: v.to(dev) for s, v in sr.items()})
for _ in range(3):
    pyr.run_packed(pts)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    pyr.run_packed(pts)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"knn pyramid {os.environ.get('TAG','')}: {ms:.4f} ms per 8-frame batch, {pyr.n_queries * B / ms / 1e3:.1f} M queries/s, "
      f"{pyr.algorithmic_bytes * B / ms / 1e6:.1f} GB/s algorithmic", flush=True)
