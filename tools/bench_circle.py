"""CircleLoss timing at the BASELINE shape: fused forward, dL/dsim kernel, whole backward (with the
two cuBLAS gradient GEMMs), next to a materialising torch-GPU version of the same formulas (what the reference's code
does on a GPU: sim [n_fg, M + 1] fp32 + ~12 elementwise passes), CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import gadm_b200  # noqa
from gadm_b200 import matching, ops, synth

dev = torch.device("cuda", 0)
B, N, M, D = int(os.environ.get("B", "8")), int(os.environ.get("N", "12800")), 8192, 128
g = torch.Generator().manual_seed(7)
rgbd, mesh, corr = synth.descriptors(B, N, M, D, n_obj=1, regime="planted", seed=2000)
xyz = synth.fibonacci_sphere(M, 0.2)[None].to(dev)
vis = (torch.rand((B, M), generator=g) < 0.6).to(dev)
labels = torch.ones((B, N), dtype=torch.long, device=dev)
match_idx = corr.to(dev)
r = 0.006
rg, me = rgbd.to(dev).requires_grad_(True), mesh.to(dev).requires_grad_(True)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def fused_fwd():
    with torch.no_grad():
        return matching.circle_match_loss(rg, me, labels, match_idx, vis, r, model_xyz=xyz)


def fused_fwd_bwd():
    rg.grad = me.grad = None
    matching.circle_match_loss(rg, me, labels, match_idx, vis, r, model_xyz=xyz, grad_gemm="fp32").backward()


def torch_materialised(backward):
    """models/geoMatch.py:102-157 + :55-83 + models/loss.py:475-490 with plain torch ops on the GPU."""
    rg.grad = me.grad = None
    pad = -torch.ones((D, 1), device=dev)
    mp = F.normalize(torch.cat([me[0], pad], dim=1), p=2, dim=0)
    tot = 0
    for i in range(B):
        sel = F.normalize(rg[i].t(), p=2, dim=1)
        sim = sel @ mp
        gt = xyz[0][match_idx[i]]
        near = torch.sqrt(((gt[:, None] - xyz[0][vis[i]][None]) ** 2).sum(2) + 1e-7) < r
        mask = torch.zeros((N, M + 1), dtype=torch.bool, device=dev)
        mask[:, :M][:, vis[i]] = near
        sd = sim.detach()
        lp = -torch.clamp_min(1.2 - sd, 0) * (sim - 0.8) * 16
        ln = torch.clamp_min(sd + 0.2, 0) * (sim - 0.2) * 16
        ninf = torch.full_like(sim, float("-inf"))
        z = torch.logsumexp(torch.where(mask, lp, ninf), 1) + torch.logsumexp(torch.where(mask, ninf, ln), 1)
        tot = tot + F.softplus(z).mean()
    tot = tot / B
    if backward:
        tot.backward()
    return tot


f = float(fused_fwd())
t = float(torch_materialised(False))
print(f"loss fused {f:.6f}  torch {t:.6f}  rel diff {abs(f - t) / abs(t):.2e}")
rows, rinv, pad_sim = ops.prep_rows(rgbd.to(dev), 0, 1)
cols, aux = ops.prep_model(mesh.to(dev), xyz, 0)
planes = torch.empty((4, B, M), device=dev)
planes[:3] = torch.where(vis[None], xyz[0].t()[:, None, :].expand(3, B, M), xyz.new_full((), 1e18))
planes[3] = r * r
fg = torch.ones((B, N), dtype=torch.uint8, device=dev)
k_fwd = timed(lambda: ops.circle_loss_fwd(rows, rinv, pad_sim, cols, aux, planes, match_idx, fg, None, 16.0, 0.2))
loss, lp_, ln_ = ops.circle_loss_fwd(rows, rinv, pad_sim, cols, aux, planes, match_idx, fg, None, 16.0, 0.2)
w = torch.full((B, N), 1.0 / (B * N), device=dev)
k_bwd = timed(lambda: ops.circle_loss_bwd(rows, rinv, pad_sim, cols, aux, planes, match_idx, None, 16.0, 0.2, lp_, ln_, w))
flop = 2.0 * N * M * D * B
print(f"circle_kernel<fwd>  {k_fwd:.3f} ms  ({flop / k_fwd / 1e9:.0f} TFLOP/s of similarity)   "
      f"circle_kernel<grad> {k_bwd:.3f} ms (writes {B * N * (M + 8) * 4 / 1e9:.2f} GB: {B * N * (M + 8) * 4 / k_bwd / 1e6:.0f} GB/s)")
print(f"fused forward (prep + kernel + reduction)   {timed(fused_fwd):.3f} ms")
print(f"fused forward + backward (fp32 grad GEMMs)  {timed(fused_fwd_bwd, 3):.3f} ms")


def fused_fwd_bwd_tf32():
    rg.grad = me.grad = None
    matching.circle_match_loss(rg, me, labels, match_idx, vis, r, model_xyz=xyz, grad_gemm="tf32").backward()


print(f"fused forward + backward (tf32 grad GEMMs)  {timed(fused_fwd_bwd_tf32, 3):.3f} ms")


def fused_fwd_bwd_split():
    rg.grad = me.grad = None
    matching.circle_match_loss(rg, me, labels, match_idx, vis, r, model_xyz=xyz, grad_gemm="bf16x2").backward()


k_split = timed(lambda: ops.circle_loss_bwd_split(rows, rinv, pad_sim, cols, aux, planes, match_idx, None, 16.0, 0.2, lp_, ln_, w))
print(f"circle_kernel<grad, split>                  {k_split:.3f} ms")
print(f"fused forward + backward (bf16x2 grad GEMMs) {timed(fused_fwd_bwd_split, 3):.3f} ms")
def fused_fwd_bwd_fused():
    rg.grad = me.grad = None
    matching.circle_match_loss(rg, me, labels, match_idx, vis, r, model_xyz=xyz, grad_gemm="fused").backward()


k_fused = timed(lambda: ops.circle_loss_bwd_fused(rows, rinv, pad_sim, cols, aux, planes, match_idx, None, 16.0, 0.2, lp_, ln_, w))
print(f"circle_df_kernel<dF> (G2 written)           {k_fused:.3f} ms")
k_full = timed(lambda: ops.circle_loss_bwd_fused(rows, rinv, pad_sim, cols, aux, planes, match_idx, None, 16.0, 0.2, lp_, ln_, w, None, True))
print(f"circle_df_kernel<dF, dM> (no G at all)      {k_full:.3f} ms")
print(f"fused forward + backward (grad_gemm='fused': dF in the kernel, dM a GEMM) {timed(fused_fwd_bwd_fused, 3):.3f} ms")


def fused_fwd_bwd_flash():
    rg.grad = me.grad = None
    matching.circle_match_loss(rg, me, labels, match_idx, vis, r, model_xyz=xyz, grad_gemm="flash").backward()


print(f"fused forward + backward (grad_gemm='flash': both products in the kernel) {timed(fused_fwd_bwd_flash, 3):.3f} ms")
# accuracy of the gradient paths against the fp32 library GEMMs
grads = {}
for mode in ("fp32", "tf32", "bf16x2", "fused", "flash"):
    rg.grad = me.grad = None
    matching.circle_match_loss(rg, me, labels, match_idx, vis, r, model_xyz=xyz, grad_gemm=mode).backward()
    grads[mode] = (rg.grad.clone(), me.grad.clone())
for mode in ("tf32", "bf16x2", "fused", "flash"):
    e = [float((grads[mode][i] - grads["fp32"][i]).abs().max() / grads["fp32"][i].abs().max()) for i in range(2)]
    print(f"max |grad - grad_fp32| / max |grad_fp32|, {mode:7s}: d rgbd {e[0]:.2e}   d mesh {e[1]:.2e}")
print(f"torch materialised forward                  {timed(lambda: torch_materialised(False), 3):.3f} ms")
print(f"torch materialised forward + backward       {timed(lambda: torch_materialised(True), 3):.3f} ms")
