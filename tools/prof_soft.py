"""Profiling driver (for ncu; not a benchmark): the SOFT matcher at the BASELINE shape, once per kernel variant --
single CTAs (the default) and CTA pairs (cta_group::2, match.cta2 = 1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gadm_b200  # noqa
from gadm_b200 import ops, synth, _lib
from gadm_b200._lib import MATCH_MODES

dev = torch.device("cuda", 0)
B, N, M, D = 8, 12800, 8192, 128
rgbd, mesh, _ = synth.descriptors(B, N, M, D, n_obj=8, regime="planted", seed=2000)
xyz = synth.model_bank_xyz(8, M).to(dev)
obj = torch.arange(B, dtype=torch.int32, device=dev)
cols, aux = ops.prep_model(mesh.to(dev), xyz, 0)
rows, rinv, pad = ops.prep_rows(rgbd.to(dev), 0, 0)
for cta2 in (1, 0, 1, 0):
    _lib.config_set("match.cta2", cta2)
    ops.match_fwd(rows, rinv, pad, cols, aux, None, obj, 16.0, 0, MATCH_MODES["soft"])
torch.cuda.synchronize()
print("done")
