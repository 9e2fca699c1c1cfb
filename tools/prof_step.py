"""Profiling driver: a few launches of each hot kernel at the BASELINE shapes (for ncu; not a benchmark)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gadm_b200  # noqa
from gadm_b200 import ops, synth
from gadm_b200._lib import MATCH_MODES
from gadm_b200.knn import KnnPyramid

dev = torch.device("cuda", 0)
B, N, M, D = 8, 12800, 8192, 128
what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rgbd, mesh, _ = synth.descriptors(B, N, M, D, n_obj=8, regime="planted", seed=2000)
xyz = synth.model_bank_xyz(8, M).to(dev)
obj = torch.arange(B, dtype=torch.int32, device=dev)
cols, aux = ops.prep_model(mesh.to(dev), xyz, 0)
cols_n, aux_n = ops.prep_model(mesh.to(dev), xyz, 2)          # bf16n: columns normalised before the rounding
rows, rinv, pad = ops.prep_rows(rgbd.to(dev), 0, 0)
cld, sr = synth.frame_batch(B, 128, N, seed=2000)
pyr = KnnPyramid(N, {s: (128 // s) ** 2 for s in (2, 4, 8)}, B)
pts = pyr.pack(cld.to(dev), {s: v.to(dev) for s, v in sr.items()})
xf = torch.randn((16, 64, 4096), generator=torch.Generator().manual_seed(5000)).to(dev)   # a quarter of config 5
for _ in range(reps):
    if what in ("all", "dgcnn"):
        ops.knn_feat(xf, 20, 64)          # tensor-core feature-space kNN (knn_feat_tc_kernel)
    if what in ("all", "match"):
        ops.match_fwd(rows, rinv, pad, cols, aux, None, obj, 16.0, 0, MATCH_MODES["argmax"])
        ops.match_fwd(rows, rinv, pad, cols, aux, None, obj, 16.0, 0, MATCH_MODES["soft"])
        ops.match_fwd(rows, rinv, pad, cols_n, aux_n, None, obj, 16.0, 0, MATCH_MODES["argmax_unit"])
    if what in ("all", "knn"):
        pyr.run_packed(pts)
if what == "circle":     # CircleLoss kernels: forward, dL/dsim (fp32 / split bf16), fused dF, fused dF + dM
    M8 = M
    planes = torch.empty((4, B, M8), device=dev)
    planes[:3] = xyz[0].t()[:, None, :].expand(3, B, M8)
    planes[3] = 0.006 ** 2
    rows1, rinv1, pad1 = ops.prep_rows(rgbd.to(dev), 0, 1)
    mi = torch.randint(0, M8, (B, N), generator=torch.Generator().manual_seed(1)).to(dev)
    fg = torch.ones((B, N), dtype=torch.uint8, device=dev)
    w = torch.full((B, N), 1.0 / (B * N), device=dev)
    for _ in range(reps):
        _, lp_, ln_ = ops.circle_loss_fwd(rows1, rinv1, pad1, cols, aux, planes, mi, fg, obj, 16.0, 0.2)
        ops.circle_loss_bwd(rows1, rinv1, pad1, cols, aux, planes, mi, obj, 16.0, 0.2, lp_, ln_, w)
        ops.circle_loss_bwd_split(rows1, rinv1, pad1, cols, aux, planes, mi, obj, 16.0, 0.2, lp_, ln_, w)
        ops.circle_loss_bwd_fused(rows1, rinv1, pad1, cols, aux, planes, mi, obj, 16.0, 0.2, lp_, ln_, w)
        ops.circle_loss_bwd_fused(rows1, rinv1, pad1, cols, aux, planes, mi, obj, 16.0, 0.2, lp_, ln_, w, None, True)
torch.cuda.synchronize()
print("done")
