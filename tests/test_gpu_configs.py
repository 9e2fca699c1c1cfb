"""-m gpu: the BASELINE.json configurations that are parity cases rather than bench lines:

  config 3  YCB-V-shaped: a 21-object model bank replicated per GPU, instances drawn with obj_id ~ U{0..20},
            N = 12800 scene points per instance, frames sharded by contiguous blocks over 2/4/8 ranks
  config 4  stress: N = 50000 scene points x 8192 vertices, d = 256 -- the score matrix (1.64 GB per frame in fp32)
            must never exist: peak device memory stays O(B N)
  config 5  geoMatch_DGCNN: feature-space kNN k = 20 on 4096 points, d = 64, batch 64 -- no [B, N, N] matrix

Full-size runs are checked through size-independent properties (planted correspondences recovered, weights in
(0, 1], self at rank 0, ...) plus the oracle on a row / batch sample; gates as in test_gpu_match.py.
"""
import pytest
import torch

from oracle import dgcnn_oracle as do
from oracle import match_oracle as mo

pytestmark = pytest.mark.gpu

TOL = 1e-3


def _check_rows(out, ref, diameter):
    idx, sim, w, sx = [o.cpu() for o in out]
    decided = ref["margin"] > TOL
    assert decided.float().mean() > 0.5
    assert torch.equal(idx[decided], ref["idx"][decided])
    assert (sim - ref["max_sim"]).abs().max() <= TOL
    assert ((w - ref["weight"]).abs() / ref["weight"]).max() <= TOL
    assert (sx - ref["soft_xyz"]).abs().max() <= TOL * diameter


def test_config3_ycbv_bank_and_frame_shards(cuda):
    from gadm_b200 import matching, sharding, synth
    n_obj, N, M, d = 21, 12800, 8192, 128
    g = torch.Generator().manual_seed(3000)
    n_inst = torch.randint(1, 7, (256,), generator=g)                 # instances per frame, U{1..6}
    assert 700 < int(n_inst.sum()) < 1100                             # ~896 instances in the whole job
    # contiguous frame blocks per rank cover the job exactly once (what every rank of bench.py --gpus G computes)
    for G in (2, 4, 8):
        blocks = [sharding.frame_range(256, r, G) for r in range(G)]
        assert blocks[0][0] == 0 and blocks[-1][1] == 256
        assert all(blocks[r][1] == blocks[r + 1][0] for r in range(G - 1))
        bal = sharding.balanced_assignment(n_inst.tolist(), G)
        loads = [sum(int(n_inst[f]) for f in b) for b in bal]
        assert sorted(f for b in bal for f in b) == list(range(256)) and max(loads) - min(loads) <= 6
    # the instances of the first frames of rank 0's block, against the replicated 21-object bank
    mesh = synth.bf16_round(torch.randn((n_obj, d, M), generator=g))
    diam = [0.10 + 0.01 * o for o in range(n_obj)]
    xyz = torch.stack([synth.fibonacci_sphere(M, diam[o]) for o in range(n_obj)])
    bank = matching.ModelBank(mesh.to(cuda), xyz.to(cuda))
    B = 10
    obj = torch.randint(0, n_obj, (B,), generator=g)
    corr = torch.randint(0, M, (B, N), generator=g)
    rgbd = torch.stack([mesh[obj[b]][:, corr[b]] + 0.5 * torch.randn((d, N), generator=g) for b in range(B)])
    rgbd = synth.bf16_round(rgbd)
    out = matching.match(rgbd.to(cuda), bank, obj_id=obj.tolist())
    idx, sim, w, sx = out
    assert (idx.cpu() == corr).float().mean() > 0.99                  # each instance matched against ITS object
    assert torch.all((w > 0) & (w <= 1 + 1e-6))
    rows = torch.arange(0, N, 25)
    for b in (0, 4, 9):
        o = int(obj[b])
        ref = mo.match_soft(rgbd[b][:, rows], mesh[o], xyz[o])
        _check_rows([t[b][rows] for t in out], ref, diam[o])


def test_config4_stress_shape_memory_is_linear(cuda):
    from gadm_b200 import matching, synth
    B, N, M, d = 2, 50000, 8192, 256
    rgbd, mesh, corr = synth.descriptors(B, N, M, d, regime="planted", seed=4000)
    diam = 0.25
    xyz = synth.fibonacci_sphere(M, diam)
    bank = matching.ModelBank(mesh.to(cuda), xyz[None].to(cuda))
    x = rgbd.to(cuda)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats(cuda)
    base = torch.cuda.memory_allocated(cuda)
    out = matching.match(x, bank)
    torch.cuda.synchronize()
    extra = torch.cuda.max_memory_allocated(cuda) - base
    # bf16 operand copy (B N d 2 = 51 MB) + outputs + norms + the per-SM stash; one fp32 score matrix is 1638 MB
    assert extra < 160e6, f"peak extra device memory {extra / 1e6:.0f} MB is not O(B N)"
    idx, sim, w, sx = out
    assert (idx.cpu() == corr).float().mean() > 0.99
    assert torch.all((w > 0) & (w <= 1 + 1e-6))
    assert torch.all(sx.norm(dim=-1) <= diam / 2 * (1 + 1e-4))
    idx2, sim2, _, _ = matching.match(x, bank, mode="argmax")
    assert torch.equal(idx, idx2) and torch.equal(sim, sim2)
    rows = torch.arange(0, N, 100)
    for b in range(B):
        ref = mo.match_soft(rgbd[b][:, rows], mesh[0], xyz)
        _check_rows([t[b][rows] for t in out], ref, diam)


def test_config5_dgcnn_full_size(cuda):
    from gadm_b200 import dgcnn, ops
    B, C, N, k = 64, 64, 4096, 20
    g = torch.Generator().manual_seed(5000)
    x = torch.randn((B, C, N), generator=g)
    xd = x.to(cuda)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats(cuda)
    base = torch.cuda.memory_allocated(cuda)
    idx = ops.knn_feat(xd, k, C)
    torch.cuda.synchronize()
    extra = torch.cuda.max_memory_allocated(cuda) - base
    # O(B N C): the int64 result (42 MB) + the tensor-core kernel's split operands [B, N, 2C] bf16 and |x|^2 (68 MB)
    assert extra < 4 * B * N * C * 2 + 8 * B * N * k + 16e6, \
        f"kNN peak extra memory {extra / 1e6:.0f} MB: a [B, N, N] matrix would be 4295 MB"
    assert idx.shape == (B, N, k) and idx.dtype == torch.int64
    assert torch.all(idx[..., 0].cpu() == torch.arange(N)[None]), "self is always rank 0"
    assert int(idx.min()) >= 0 and int(idx.max()) < N
    for b in (0, 37, 63):                                             # oracle = models/dgcnn.py:21-27 on the CPU
        ref_idx, gaps, vals = do.knn_with_gaps(x[b:b + 1], k)
        ok = gaps > 2e-4 * max(1.0, float(vals.abs().max()) / 79.0)
        assert ok.float().mean() > 0.9
        assert torch.equal(idx[b:b + 1].cpu()[ok], ref_idx[ok])
    out = dgcnn.get_graph_feature(xd, k=k, idx=idx)                   # [64, 128, 4096, 20] fp32 = 2.68 GB
    assert out.shape == (B, 2 * C, N, k)
    for b in (0, 63):
        ref = do.get_graph_feature(x[b:b + 1], k=k, idx=idx[b:b + 1].cpu())
        assert torch.equal(out[b:b + 1].cpu(), ref)


def test_frame_stream_host_buffers_match_direct_calls(cuda):
    """The end-to-end front end (pipeline.FrameStream: ONE packed pinned H2D with bf16 descriptors, ONE packed D2H of
    {int32 idx, max_sim, weight, xyz} records + kNN indices, three batches in flight) returns, bit for bit, what the
    direct calls return on the same inputs -- and bf16 host descriptors lose nothing (the matcher's operands are bf16)."""
    from gadm_b200 import matching, ops, synth
    from gadm_b200.knn import KnnPyramid
    from gadm_b200.pipeline import FrameStream
    B, N, M, d, in_size = 2, 4096, 1024, 128, 32
    xyz = synth.model_bank_xyz(2, M).to(cuda)
    bank = None
    pyr = KnnPyramid(N, {s: (in_size // s) ** 2 for s in (2, 4, 8)}, B)
    obj = torch.tensor([1, 0], dtype=torch.int32, device=cuda)
    fs, batches, inputs = None, [], []
    for i in range(4):                                         # more batches than slots: every slot is reused
        rgbd, mesh, _ = synth.descriptors(B, N, M, d, n_obj=2, regime="planted", seed=40 + i)
        if bank is None:
            bank = matching.ModelBank(mesh.to(cuda), xyz)
            fs = FrameStream(bank, pyr, B, d, N, obj_id=obj, gamma=16.0, mode="soft", depth=3)
        cld, sr = synth.frame_batch(B, in_size, N, seed=40 + i)
        inputs.append((rgbd, cld, sr))
        batches.append(fs.host_batch().fill(rgbd, cld, sr))
    assert fs.h2d_bytes == B * d * N * 2 + B * pyr.P * 12
    tickets = [fs.submit(hb) for hb in batches[:3]]
    got = [{k: v.clone() for k, v in fs.result(t).items()} for t in tickets]
    got.append({k: v.clone() for k, v in fs.result(fs.submit(batches[3])).items()})
    with pytest.raises(ValueError):
        fs.result(0)                                           # its buffers were reused by ticket 3
    for (rgbd, cld, sr), out in zip(inputs, got):
        idx, sim, w, sx = matching.match(rgbd.to(cuda), bank, obj_id=obj, gamma=16.0, mode="soft")
        assert torch.equal(out["idx"].long(), idx.cpu()) and torch.equal(out["max_sim"], sim.cpu())
        assert torch.equal(out["weight"], w.cpu()) and torch.equal(out["soft_xyz"], sx.cpu())
        knn = pyr.run_packed(pyr.pack(cld.to(cuda), {s: v.to(cuda) for s, v in sr.items()}))
        assert out["knn"].dtype == torch.uint16                      # narrowed for transport: every cloud < 65536 points
        assert torch.equal(out["knn"].to(torch.int32), knn.cpu())
        assert torch.equal(pyr.unpack(out["knn"])["cld_nei_idx0"], pyr.unpack(knn.cpu())["cld_nei_idx0"])
    # fp32 device descriptors and their bf16 copies prepare identical operands
    rows32 = ops.prep_rows(inputs[0][0].to(cuda), 0, 1)
    rows16 = ops.prep_rows(inputs[0][0].to(cuda).to(torch.bfloat16), 0, 1)
    assert all(torch.equal(a, b) for a, b in zip(rows32, rows16))


def test_step_is_cuda_graph_capturable(cuda):
    """The C ABI never allocates, synchronises or reads device memory on the host: one whole step (prep_rows, the SOFT
    matcher with its workspace memset, pack, the kNN pyramid with its grid build) is captured into a CUDA graph and
    replayed on new input values in the same buffers; every replay equals the eager result bit for bit."""
    from gadm_b200 import ops, synth, matching
    from gadm_b200._lib import MATCH_MODES
    from gadm_b200.knn import KnnPyramid
    B, N, M, d = 2, 1024, 1024, 128
    xyz = synth.model_bank_xyz(B, M).to(cuda)
    pyr = KnnPyramid(N, {s: (32 // s) ** 2 for s in (2, 4, 8)}, B)
    ws = ops._lib.load().gadm_knn3d_workspace_bytes(pyr.jobs, len(pyr.jobs), ops.KNN_ALGOS["auto"])
    pyr.workspace = torch.empty((max(ws, 16),), dtype=torch.uint8, device=cuda)
    obj = torch.arange(B, dtype=torch.int32, device=cuda)
    rgbd_buf = torch.empty((B, d, N), device=cuda)
    pts_buf = torch.empty((B * pyr.P, 3), device=cuda)
    rec = torch.empty((B, N, 6), dtype=torch.int32, device=cuda)
    knn_out = torch.empty((pyr.out_elems,), dtype=torch.int32, device=cuda)

    def make_inputs(seed):
        rgbd, mesh, _ = synth.descriptors(B, N, M, d, n_obj=B, regime="planted", seed=seed)
        cld, sr = synth.frame_batch(B, 32, N, seed=seed)
        return rgbd.to(cuda), mesh.to(cuda), pyr.pack(cld.to(cuda), {s: v.to(cuda) for s, v in sr.items()})

    rgbd0, mesh0, pts0 = make_inputs(1)
    cols, aux = ops.prep_model(mesh0, xyz, 0)

    def step():
        rows, rinv, pad = ops.prep_rows(rgbd_buf, 0, 0)
        o = ops.match_fwd(rows, rinv, pad, cols, aux, None, obj, 16.0, 0, MATCH_MODES["soft"])
        ops.pack_match_outputs(o[0], o[1], o[2], o[3], rec)
        pyr.run_packed(pts_buf, out=knn_out)

    rgbd_buf.copy_(rgbd0); pts_buf.copy_(pts0)
    side = torch.cuda.Stream(device=cuda)
    side.wait_stream(torch.cuda.current_stream(cuda))
    with torch.cuda.stream(side):
        step()                                   # warm-up outside the capture (workspace cache, lazy init)
    torch.cuda.current_stream(cuda).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    for seed in (1, 2, 3):
        rgbd_s, _, pts_s = make_inputs(seed)
        rgbd_buf.copy_(rgbd_s); pts_buf.copy_(pts_s)
        graph.replay()
        torch.cuda.synchronize()
        got_rec, got_knn = rec.clone(), knn_out.clone()
        step()
        torch.cuda.synchronize()
        assert torch.equal(got_rec, rec) and torch.equal(got_knn, knn_out), f"replay differs from eager (seed {seed})"
