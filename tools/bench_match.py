"""Matcher-only timing (CUDA events) at the BASELINE shape: both modes; GADM_MATCH_DBG=1/2 for ceilings."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gadm_b200  # noqa
from gadm_b200 import ops, synth
from gadm_b200._lib import MATCH_MODES

dev = torch.device("cuda", 0)
B, N, M, D = int(os.environ.get("B", "8")), int(os.environ.get("N", "12800")), 8192, int(os.environ.get("D", "128"))
regime = os.environ.get("REGIME", "planted")
rgbd, mesh, _ = synth.descriptors(B, N, M, D, n_obj=8, regime=regime, seed=2000)
xyz = synth.model_bank_xyz(8, M).to(dev)
obj = torch.arange(B, dtype=torch.int32, device=dev) % 8
cols, aux = ops.prep_model(mesh.to(dev), xyz, 0)
rows, rinv, pad = ops.prep_rows(rgbd.to(dev), 0, 0)
flop = 2.0 * N * M * D * B
if os.environ.get("OPERAND", "bf16") == "bf16n":
    cols, aux = ops.prep_model(mesh.to(dev), xyz, 2)
for mode in (("argmax_unit", "argmax_bf16n") if os.environ.get("OPERAND", "bf16") == "bf16n" else ()) + ("argmax", "soft"):
    for _ in range(3):
        ops.match_fwd(rows, rinv, pad, cols, aux, None, obj, 16.0, 0, MATCH_MODES[mode])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ops.match_fwd(rows, rinv, pad, cols, aux, None, obj, 16.0, 0, MATCH_MODES[mode])
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"{mode} B={B} N={N} d={D} {regime} dbg={os.environ.get('GADM_MATCH_DBG','0')}: {ms:.4f} ms  {flop/ms/1e9:.1f} TFLOP/s  frac {flop/ms/1e9/1658.8:.3f}", flush=True)
