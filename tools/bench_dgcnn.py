"""Secondary measurement (BASELINE.json configs[4]): the geoMatch_DGCNN graph path --
feature-space kNN k=20 on 4096 pts, batch 64, 1 x (C=3 via dim9) + 3 x (C=64) layers, + get_graph_feature.
Prints one JSON line with per-kernel times and roofline fractions (CUDA events, 3 warm-ups, inputs > L2), next to the
reference's own formulation on the same GPU: torch `matmul` + `topk` (models/dgcnn.py:21-27, materialises the
[B, N, N] distance matrix: 4.3 GB at this shape) and the gather / cat / permute of get_graph_feature (:30-56)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gadm_b200  # noqa
from gadm_b200 import ops

dev = torch.device("cuda", 0)
B, C, N, k = 64, 64, 4096, 20
g = torch.Generator().manual_seed(5000)
x = torch.randn((B, C, N), generator=g).to(dev)
x9 = torch.randn((B, 9, N), generator=g).to(dev)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}


def timed(fn, reps=5):
    t_end = time.time() + 0.4          # ramp the clocks: the box idles at ~120 MHz
    while time.time() < t_end:
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


t_knn64, idx = timed(lambda: ops.knn_feat(x, k, C))
from gadm_b200 import dgcnn
t_knn3_feat, _ = timed(lambda: ops.knn_feat(x9, k, 3))          # the feature-space kernel on three channels
t_knn3, idx3 = timed(lambda: dgcnn.knn_xyz(x9, k))              # what get_graph_feature(dim9=True) runs: the 3-D grid search
# graph_feature through the C ABI with a preallocated output and workspace (ops.graph_feature allocates 2.7 GB per call,
# which is what an earlier version of this script was really timing)
lib = ops._lib.load()
out = torch.empty((B, 2 * C, N, k), dtype=torch.float32, device=dev)
ws_bytes = lib.gadm_graph_feature_workspace_bytes(B, C, N)
ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def gf():
    rc = lib.gadm_graph_feature(x.data_ptr(), idx.data_ptr(), B, C, N, k, out.data_ptr(), ws.data_ptr(), ws_bytes, stream)
    assert rc == 0
    return out


t_gf, out = timed(gf, reps=20)


def torch_knn(x_, k_):                 # models/dgcnn.py:21-27, restated line by line
    inner = -2 * torch.matmul(x_.transpose(2, 1), x_)
    xx = torch.sum(x_ ** 2, dim=1, keepdim=True)
    pairwise_distance = -xx - inner - xx.transpose(2, 1)
    return pairwise_distance.topk(k=k_, dim=-1)[1]


def torch_graph_feature(x_, idx_):     # models/dgcnn.py:30-56 (dim9=False), restated
    batch_size, num_dims, num_points = x_.shape
    idx_base = torch.arange(0, batch_size, device=x_.device).view(-1, 1, 1) * num_points
    idx_ = (idx_ + idx_base).view(-1)
    xt = x_.transpose(2, 1).contiguous()
    feature = xt.view(batch_size * num_points, -1)[idx_, :].view(batch_size, num_points, k, num_dims)
    xt = xt.view(batch_size, num_points, 1, num_dims).repeat(1, 1, k, 1)
    return torch.cat((feature - xt, xt), dim=3).permute(0, 3, 1, 2).contiguous()


torch.backends.cuda.matmul.allow_tf32 = False
t_torch_knn, idx_t = timed(lambda: torch_knn(x, k), reps=3)
t_torch_gf, _ = timed(lambda: torch_graph_feature(x, idx), reps=3)
agree = float((idx_t == idx).float().mean())
t_gf_alloc, _ = timed(lambda: ops.graph_feature(x, idx))
gf_bytes = 4 * B * C * N + 8 * B * N * k + 4 * B * 2 * C * N * k
line = {
    "workload": "dgcnn_graph: B=64, N=4096, k=20; knn C=64, knn C=3 (dim9), get_graph_feature C=64",
    "knn_feat_c64_ms": t_knn64, "knn_feat_c64_tflops_fp32": 2.0 * B * N * N * C / (t_knn64 * 1e-3) / 1e12,
    "knn_feat_c3_ms": t_knn3, "knn_feat_c3_feature_space_kernel_ms": t_knn3_feat,
    "graph_feature_ms": t_gf, "graph_feature_with_output_allocation_ms": t_gf_alloc, "graph_feature_gbs": gf_bytes / (t_gf * 1e-3) / 1e9,
    "graph_feature_frac_of_hbm_peak": gf_bytes / (t_gf * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0),
    "layer_stack_ms (1x dim9 + 3x C=64 knn+graph)": t_knn3 + 3 * t_knn64 + 4 * t_gf,
    "torch_gpu_arm": {"knn_c64_ms (matmul fp32 + topk, models/dgcnn.py:21-27)": t_torch_knn,
                      "graph_feature_ms (gather + cat + permute, :30-56)": t_torch_gf,
                      "knn_speedup": t_torch_knn / t_knn64, "graph_feature_speedup": t_torch_gf / t_gf,
                      "knn_index_agreement": agree},
}
print(json.dumps(line))
