"""match_fa_kernel timing ablations (gadm_config_set "match.dbg": results are wrong, only the time matters)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gadm_b200  # noqa
from gadm_b200 import ops, synth, _lib
from gadm_b200._lib import MATCH_MODES

dev = torch.device("cuda", 0)
B, N, M, D = 8, 12800, 8192, 128
for regime in ("planted",):
    rgbd, mesh, _ = synth.descriptors(B, N, M, D, n_obj=8, regime=regime, seed=2000)
    xyz = synth.model_bank_xyz(8, M).to(dev)
    obj = torch.arange(B, dtype=torch.int32, device=dev) % 8
    cols, aux = ops.prep_model(mesh.to(dev), xyz, 0)
    rows, rinv, pad = ops.prep_rows(rgbd.to(dev), 0, 0)
    flop = 2.0 * N * M * D * B
    names = {0: "as built", 1: "no PV MMAs", 8: "no epilogue arithmetic", 9: "no epilogue arithmetic, no PV",
             25: "9 + no P store", 41: "9 + UMMA thread spins on p_full", 73: "9 + epilogue spins on s_full",
             121: "9 + no P store + both spin", 32: "UMMA thread spins", 96: "both spin"}
    names.update({2: "no exponentials", 4: "no maximum tree", 6: "no exp, no max tree", 3: "no PV, no exp", 7: "no PV, no exp, no max"})
    for dbg in (0, 1, 2, 4, 6, 3, 7, 8, 9):
        _lib.config_set("match.dbg", dbg)
        for _ in range(3):
            ops.match_fwd(rows, rinv, pad, cols, aux, None, obj, 16.0, 0, MATCH_MODES["soft"])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(30):
            ops.match_fwd(rows, rinv, pad, cols, aux, None, obj, 16.0, 0, MATCH_MODES["soft"])
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 30
        print(f"{regime:8s} dbg={dbg:2d} {names[dbg]:32s} {ms:.4f} ms  frac {flop/ms/1e9/1658.8:.3f}", flush=True)
    _lib.config_set("match.dbg", -1)
