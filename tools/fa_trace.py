"""One match_fa_kernel launch of the trace build (GADM_LIB=.../libgadm_trace.so): clock64 accounting of two CTAs."""
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
obj = torch.arange(B, dtype=torch.int32, device=dev) % 8
cols, aux = ops.prep_model(mesh.to(dev), xyz, 0)
rows, rinv, pad = ops.prep_rows(rgbd.to(dev), 0, 0)
for dbg in [int(x) for x in os.environ.get("DBG", "0,9,8").split(",")]:
    _lib.config_set("match.dbg", dbg)
    torch.cuda.synchronize()
    print(f"---- dbg={dbg}", flush=True)
    ops.match_fwd(rows, rinv, pad, cols, aux, None, obj, 16.0, 0, MATCH_MODES["soft"])
    torch.cuda.synchronize()
