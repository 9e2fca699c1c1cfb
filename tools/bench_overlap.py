"""Experiment: one bench step (prep + SOFT (MATCH_MODE=argmax: ARGMAX) match + kNN pyramid) with the kNN pyramid on a second stream, so that its
CTAs (40 registers, no shared memory) share the SMs with the matcher's (1 CTA per SM, shared-memory bound)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gadm_b200  # noqa
from gadm_b200 import ops, synth
from gadm_b200._lib import MATCH_MODES
from gadm_b200.knn import KnnPyramid

MM = os.environ.get("MATCH_MODE", "soft")          # soft | argmax
for kv in os.environ.get("CONFIG", "").split(","):  # e.g. CONFIG=match.alt_cta2=0
    if kv:
        ops._lib.config_set(kv.split("=")[0], int(kv.split("=")[1]))
dev = torch.device("cuda", 0)
B, N, M, D = 8, 12800, 8192, 128
sets = []
for r in range(4):
    rgbd, mesh, _ = synth.descriptors(B, N, M, D, n_obj=8, regime="planted", seed=2000 + r)
    cld, sr = synth.frame_batch(B, 128, N, seed=2000 + r)
    sets.append((rgbd.to(dev), mesh.to(dev), cld.to(dev), {s: v.to(dev) for s, v in sr.items()}))
xyz = synth.model_bank_xyz(8, M).to(dev)
obj = torch.arange(B, dtype=torch.int32, device=dev)
pyr = KnnPyramid(N, {s: (128 // s) ** 2 for s in (2, 4, 8)}, B)
ws = ops._lib.load().gadm_knn3d_workspace_bytes(pyr.jobs, len(pyr.jobs), ops.KNN_ALGOS["auto"])
pyr.workspace = torch.empty((max(ws, 16),), dtype=torch.uint8, device=dev)
pts = [pyr.pack(s[2], s[3]) for s in sets]
prio = os.environ.get("PRIO", "0") == "1"
side = torch.cuda.Stream(device=dev, priority=0)
main_hi = torch.cuda.Stream(device=dev, priority=-1) if prio else None


mm_ev = []


def step(i, mode):
    rgbd, mesh, _, _ = sets[i % 4]
    cur = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mm_ev.append((e0, e1))
    if mode == "serial":
        cols, aux = ops.prep_model(mesh, xyz, 0)
        rows, rinv, pad = ops.prep_rows(rgbd, 0, 0)
        e0.record()
        out = ops.match_fwd(rows, rinv, pad, cols, aux, None, obj, 16.0, 0, MATCH_MODES[MM])
        e1.record()
        knn = pyr.run_packed(pts[i % 4])
        return out, knn
    cols, aux = ops.prep_model(mesh, xyz, 0)
    rows, rinv, pad = ops.prep_rows(rgbd, 0, 0)
    fork = torch.cuda.Event(); fork.record(cur)
    if mode == "match_first":
        e0.record()
        out = ops.match_fwd(rows, rinv, pad, cols, aux, None, obj, 16.0, 0, MATCH_MODES[MM])
        e1.record()
    with torch.cuda.stream(side):
        side.wait_event(fork)
        knn = pyr.run_packed(pts[i % 4])
        join = torch.cuda.Event(); join.record(side)
    if mode == "knn_first":
        e0.record()
        out = ops.match_fwd(rows, rinv, pad, cols, aux, None, obj, 16.0, 0, MATCH_MODES[MM])
        e1.record()
    cur.wait_event(join)
    return out, knn


for mode in ("serial", "match_first", "knn_first"):
    ctx = torch.cuda.stream(main_hi) if (prio and mode != "serial") else torch.cuda.stream(torch.cuda.current_stream())
    with ctx:
        for i in range(4):
            step(i, mode)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        mm_ev.clear()
        for i in range(20):
            step(i, mode)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        mm = sum(x.elapsed_time(y) for x, y in mm_ev) / len(mm_ev)
    print(f"{mode:12s} prio={int(prio)}: {ms:.4f} ms per step  {B / ms * 1e3:.0f} frames/s; match kernel {mm:.4f} ms", flush=True)
