"""Small shapes of every kernel for quick triage (compute-sanitizer is closed on the GPU pool; run it plainly or under a sanitizer elsewhere): ragged tiles, both matcher modes and row-tile
variants, pad modes, masks, kNN grid + brute, DGCNN, gathers."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gadm_b200  # noqa
from gadm_b200 import matching, ops, synth, randla, dgcnn
from gadm_b200.knn import KnnPyramid

dev = torch.device("cuda", 0)
for (B, N, M, d) in [(2, 300, 520, 64), (1, 700, 1000, 128), (1, 130, 264, 256)]:
    rgbd, mesh, _ = synth.descriptors(B, N, M, d, n_obj=B, regime="planted", seed=3)
    xyz = synth.model_bank_xyz(B, M).to(dev)
    mask = (torch.rand((B, N)) > 0.3).to(dev)
    for mode in ("soft", "argmax"):
        for pad in ("none", "minus_one"):
            matching.match(rgbd.to(dev), mesh.to(dev), xyz, mask=mask, pad_mode=pad, mode=mode)
rgbd, mesh, _ = synth.descriptors(1, 200, 264, 64, regime="random", seed=4)
matching.match(rgbd.to(dev), mesh.to(dev), synth.model_bank_xyz(1, 264).to(dev), operand_mode="bf16x3")
cld, sr = synth.frame_batch(2, 64, 3200, seed=5)
pyr = KnnPyramid(3200, {s: (64 // s) ** 2 for s in (2, 4, 8)}, 2)
pyr(cld.to(dev), {s: v.to(dev) for s, v in sr.items()})
x = torch.randn((2, 16, 300)).to(dev)
dgcnn.get_graph_feature(x, k=8)
f = torch.randn((2, 5, 300, 1)).to(dev)
idx = torch.randint(0, 300, (2, 77, 16)).to(dev)
randla.random_sample(f, idx)
randla.relative_pos_encoding(torch.rand((2, 300, 3)).to(dev), torch.randint(0, 300, (2, 300, 16)).to(dev))
ops.seg_mask(torch.randn((2, 2, 300)).to(dev))
torch.cuda.synchronize()
print("sanitize_small done")
