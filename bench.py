#!/usr/bin/env python
"""bench.py -- frames/s of the scene-to-model dense correspondence path (BASELINE.json metric).

Workload (BASELINE.json configs[1]): an LM-O-shaped batch of 8 frames, one object each (8-object model bank):
per frame 12800 scene points x 8192 model vertices, d = 128, FULL matching (normalise + similarity + argmax +
softmax weight + soft coordinates) + the 22-call kNN pyramid of the reference's data layer
(datasets/lm/linemod_pbr.py:534-569; 128x128 crop).  A step = one such batch.

  python bench.py [--gpus N] [--steps K] [--warmup W]         our arm (CUDA, libgadm.so)
  python bench.py --impl reference ...                        the reference's CPU path on the host cores
  python bench.py --config ycbv [--balance] ...               BASELINE.json configs[2]: ONE 256-frame YCB-V-shaped job
                                                              (up to 6 instances per frame, 21-object bank) split over
                                                              the ranks with sharding.frame_range (strong scaling)
For N > 1 launch with torchrun (one rank per GPU); frames are sharded by rank (weak scaling: 8 frames per rank
per step), the only collective is the all_gather of the matcher outputs, overlapped on a side stream -- in the
device-resident loop AND in the end-to-end loop.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES = 8            # frames per rank per step
N_PTS, M_VERTS, D = 12800, 8192, 128
IN_SIZE = 128
N_OBJ = 8
GAMMA = 16.0
ROT = 4               # distinct resident input batches rotated through the timed steps (> L2)
WORKLOAD = "lmo_batch8: 8 frames x 1 object, 12800 pts x 8192 verts, d=128, full matching (soft) + 22-call kNN pyramid"


def shared_config():
    """The workload description both arms print verbatim (arm-specific settings go to `impl_config`)."""
    return {"workload": WORKLOAD, "frames_per_step_per_gpu": FRAMES, "n_obj": N_OBJ, "gamma": GAMMA,
            "descriptors": "bf16-representable fp32 values, planted regime (f = m_c + 0.5 eps)",
            "cloud": f"{IN_SIZE}x{IN_SIZE} depth crop, {N_PTS} points, 22 kNN calls per frame"}


def profile_metrics(kernel_substr):
    """Per-launch counters of one kernel from the committed ncu capture (profiles/kernel_metrics.json, written by
    tools/ncu_extract.py from the `ncu --set full` report of the same step); None when the kernel is not in it."""
    p = os.path.join(ROOT, "profiles", "kernel_metrics.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        table = json.load(f)
    for name, m in table.get("kernels", {}).items():
        if kernel_substr in name:
            return dict(m, source=table.get("source"))
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return j, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class near_gpu:
    """with near_gpu(torch, index): the calling THREAD runs on the CPUs of the NUMA node the GPU hangs off, so that the
    pinned host buffers allocated inside land next to it (H2D over PCIe from the far socket is what makes the end-to-end
    leg vary from run to run).  Only the calling thread is moved and its affinity is restored on exit: worker threads
    (torch's CPU pool, created beforehand) and the CPU-baseline leg keep every core.  Best effort: .node is None when
    the topology is not visible (containers often hide it)."""

    def __init__(self, torch, index):
        self.torch, self.index, self.node, self.saved = torch, index, None, None

    def __enter__(self):
        self.saved = os.sched_getaffinity(0)
        self.node = bind_to_gpu_numa_node(self.torch, self.index)
        return self

    def __exit__(self, *exc):
        try:
            os.sched_setaffinity(0, self.saved)
        except Exception:
            pass
        return False


def bind_to_gpu_numa_node(torch, index):
    try:
        prop = torch.cuda.get_device_properties(index)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)              # never leave the cgroup's own set
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_frames(n_frames, seed, with_soft=False):
    """The reference's own CPU path for n_frames frames: torch-CPU matching head (evaluator.py:89-93 restated in
    oracle/match_oracle.py) with all host threads + the 22-call nanoflann schedule (compiled reference when
    present, else the C port), one frame per worker thread like the reference's DataLoader workers.
    Returns (seconds, kind, cores)."""
    import numpy as np
    import torch
    from concurrent.futures import ThreadPoolExecutor
    sys.path.insert(0, ROOT)
    import gadm_b200  # noqa: F401  (synthetic generators only; no CUDA is touched on this path)
    from gadm_b200 import synth
    from oracle import knn_oracle as ko, match_oracle as mo

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind = "reference" if ko.have_reference() else "port"
    knn_fn = ko.knn_search_ref if kind == "reference" else (lambda s, q, k: ko.knn_port(s, q, k).astype(np.int32))
    rgbd, mesh, _ = synth.descriptors(n_frames, N_PTS, M_VERTS, D, n_obj=min(N_OBJ, n_frames), regime="planted",
                                      seed=seed)
    xyz = synth.model_bank_xyz(min(N_OBJ, n_frames), M_VERTS)
    clouds = [synth.depth_cloud(IN_SIZE, N_PTS, seed + b) for b in range(n_frames)]

    def knn_frame(b):
        cld, sr = clouds[b]
        return [knn_fn(s[None], q[None], k) for _, s, q, k in ko.schedule(cld, sr)]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=min(cores, n_frames)) as ex:
        futs = [ex.submit(knn_frame, b) for b in range(n_frames)]
        for b in range(n_frames):
            o = b % mesh.shape[0]
            if with_soft:
                mo.match_soft(rgbd[b], mesh[o], xyz[o], gamma=GAMMA)
            else:
                mo.match_ref(rgbd[b], mesh[o])       # normalize, normalize, matmul, torch.max -- nothing else
        for f in futs:
            f.result()
    return time.perf_counter() - t0, kind, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames_per_step = FRAMES
    for _ in range(args.warmup):
        cpu_reference_frames(frames_per_step, 1000)
    t = 0.0
    kind, cores = "port", 1
    for s in range(args.steps):
        dt, kind, cores = cpu_reference_frames(frames_per_step, 1000 + s)
        t += dt
    value = frames_per_step * args.steps / t
    sample = (f"{frames_per_step} frames/step: torch-CPU fp32 matching head (normalize, normalize, matmul, max; "
              f"evaluator.py:89-93) on {cores} threads + 22-call nanoflann kNN schedule per frame "
              f"({'compiled reference knn_.cxx' if kind == 'reference' else 'C port'}, one frame per worker thread)")
    line = {"impl": "reference", "metric": "frames/s", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(),
            "impl_config": {"frames_per_step": frames_per_step, "threads": cores, "knn": kind},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import gadm_b200  # noqa: F401
    from gadm_b200 import matching, ops, synth
    from gadm_b200._lib import MATCH_MODES, OPERAND_MODES, PAD_MODES
    from gadm_b200.knn import KnnPyramid

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:                                    # several ranks share the host: do not oversubscribe its cores
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    torch.randn(1 << 20).sum()                       # torch's CPU worker pool exists from here on
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version to stdout when the communicator is created; stdout must carry the one JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    pk, pk_kind = peaks()

    # ---- synthetic inputs: ROT distinct batches, host-pinned and device-resident copies
    host, res = [], []
    for r in range(ROT):
        seed = 2000 + 97 * r + 1000 * rank
        rgbd, mesh, _ = synth.descriptors(FRAMES, N_PTS, M_VERTS, D, n_obj=N_OBJ, regime="planted", seed=seed)
        cld, sr = synth.frame_batch(FRAMES, IN_SIZE, N_PTS, seed=seed)
        host.append({"rgbd": rgbd, "cld": cld, "sr": sr})   # packed into pinned HostBatches below
        res.append({"rgbd": rgbd.to(dev), "mesh": mesh.to(dev), "cld": cld.to(dev),
                    "sr": {s: sr[s].to(dev) for s in (2, 4, 8)}})
    xyz = synth.model_bank_xyz(N_OBJ, M_VERTS).to(dev)
    obj_id = torch.arange(FRAMES, dtype=torch.int32, device=dev) % N_OBJ
    pyr = KnnPyramid(N_PTS, {s: (IN_SIZE // s) ** 2 for s in (2, 4, 8)}, FRAMES)
    ws_bytes = ops._lib.load().gadm_knn3d_workspace_bytes(pyr.jobs, len(pyr.jobs), ops.KNN_ALGOS["auto"])
    pyr.workspace = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
    for r in res:
        r["pts"] = pyr.pack(r["cld"], r["sr"])
    om, pm, mm = OPERAND_MODES["bf16"], PAD_MODES["none"], MATCH_MODES["soft"]
    side = torch.cuda.Stream(device=dev)
    gathered = None

    knn_stream = torch.cuda.Stream(device=dev)

    def step(inp, ev=None):
        """One pass of the hot path over one resident batch.  Returns the outputs.
        The kNN pyramid (no data dependence on the matcher) runs on a second stream: its CTAs (40 registers, no shared
        memory) fit next to the matcher's single CTA per SM and fill the issue slots that kernel leaves idle; the
        matcher's own duration is unchanged by it (tools/bench_overlap.py: 0.403 ms alone, 0.405 ms overlapped)."""
        main = torch.cuda.current_stream()
        cols, aux = ops.prep_model(inp["mesh"], xyz, om)                      # model side (evaluator.py:90)
        rows, rinv, pad = ops.prep_rows(inp["rgbd"], om, pm)                  # scene side (evaluator.py:89)
        if args.no_overlap:
            if ev:
                ev[0].record()
            out = ops.match_fwd(rows, rinv, pad, cols, aux, None, obj_id, GAMMA, pm, mm)
            if ev:
                ev[1].record(); ev[2].record()
            knn_idx = pyr.run_packed(inp["pts"])
            if ev:
                ev[3].record()
            return out, knn_idx
        fork = torch.cuda.Event()
        fork.record(main)
        if ev:
            ev[0].record(main)
        out = ops.match_fwd(rows, rinv, pad, cols, aux, None, obj_id, GAMMA, pm, mm)   # evaluator.py:91-93 + ext.
        if ev:
            ev[1].record(main)
        with torch.cuda.stream(knn_stream):
            knn_stream.wait_event(fork)
            if ev:
                ev[2].record(knn_stream)
            knn_idx = pyr.run_packed(inp["pts"])                              # linemod_pbr.py:534-569
            if ev:
                ev[3].record(knn_stream)
            join = torch.cuda.Event()
            join.record(knn_stream)
        main.wait_event(join)
        return out, knn_idx

    ring = {"k": 0, "buf": [None, None], "ev": [None, None]}

    def gather(out):
        """N > 1: all_gather of the fixed-stride matcher outputs on a side stream (overlaps the next step).  The packed
        buffers are a persistent ring of two (no record_stream(): blocks released under a foreign stream return to
        the caching allocator at unpredictable times and a step that finds none free pays a cudaMalloc)."""
        nonlocal gathered
        if world == 1:
            return
        k = ring["k"] % 2
        ring["k"] += 1
        if ring["buf"][k] is None:
            ring["buf"][k] = torch.empty((FRAMES, N_PTS, 6), dtype=torch.int32, device=dev)
            ring["ev"][k] = torch.cuda.Event()
            gathered = gathered if gathered is not None else torch.empty((world, FRAMES, N_PTS, 6),
                                                                         dtype=torch.int32, device=dev)
        else:
            torch.cuda.current_stream().wait_event(ring["ev"][k])    # the all_gather that read this buffer two steps ago
        packed = ring["buf"][k]
        ops.pack_match_outputs(out[0], out[1], out[2], out[3], packed)   # {int32 idx, max_sim, weight, xyz} records
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(side):
            side.wait_event(done)
            dist.all_gather_into_tensor(gathered.view(-1), packed.view(-1))
            ring["ev"][k].record(side)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput (`value`)
    for w in range(args.warmup):
        gather(step(res[w % ROT])[0])
    sync_all()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    sync_all()
    e0.record()
    for s in range(args.steps):
        gather(step(res[s % ROT], evs[s])[0])
    e1.record()
    torch.cuda.current_stream().wait_stream(side)
    sync_all()
    ms_total = e0.elapsed_time(e1)
    match_ms = statistics.mean(ev[0].elapsed_time(ev[1]) for ev in evs)
    knn_ms = statistics.mean(ev[2].elapsed_time(ev[3]) for ev in evs)
    clocks = sampler.stop() if rank == 0 else None

    # ---- variants: the ARGMAX kernels on the same operands.  Each: 5 warm-up launches, then three bursts of 10 launches
    # between CUDA events; the median burst is reported (one burst right after the step loop is noisy: +-3 %).
    cols, aux = ops.prep_model(res[0]["mesh"], xyz, om)
    rows, rinv, pad = ops.prep_rows(res[0]["rgbd"], om, pm)
    cols_n, aux_n = ops.prep_model(res[0]["mesh"], xyz, OPERAND_MODES["bf16n"])
    va, vb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def burst_ms(cols_, aux_, mode):
        for _ in range(5):
            ops.match_fwd(rows, rinv, pad, cols_, aux_, None, obj_id, GAMMA, pm, MATCH_MODES[mode])
        t = []
        for _ in range(3):
            va.record()
            for _ in range(10):
                ops.match_fwd(rows, rinv, pad, cols_, aux_, None, obj_id, GAMMA, pm, MATCH_MODES[mode])
            vb.record()
            torch.cuda.synchronize()
            t.append(va.elapsed_time(vb) / 10)
        return statistics.median(t)
    # the reference's own path (hard argmax only, evaluator.py:93)
    argmax_ms = burst_ms(cols, aux, "argmax")
    # the same search on columns normalised BEFORE the bf16 rounding (operand mode bf16n), without the per-column scale
    # in the epilogue (GADM_MATCH_ARGMAX_UNIT); the winner's similarity carries its true scale
    unit_ms = burst_ms(cols_n, aux_n, "argmax_unit")
    # exact argmax on the same bf16n operands (bit-identical to ARGMAX; 32-column chunks that cannot beat a running
    # maximum are skipped, which is safe because every column scale is <= 1 + 2^-8)
    pruned_ms = burst_ms(cols_n, aux_n, "argmax_bf16n")

    # ---- variant: row compaction (evaluator.py:82-88).  A real frame's foreground is a fraction of the 12800 samples; with a
    # segmentation mask the selected rows are compacted on the device and only they are matched.  20 % foreground:
    gm = torch.Generator().manual_seed(5)
    fg = (torch.rand((FRAMES, N_PTS), generator=gm) < 0.2).to(torch.uint8).to(dev)

    def compacted(mode):
        pos, row_map, n_sel = ops.compact_rows(fg)
        r2, ri2, pd2 = ops.prep_rows_sel(res[0]["rgbd"], pos, om, pm)
        return ops.match_fwd_sel(r2, ri2, pd2, cols, aux, n_sel, row_map, obj_id, GAMMA, pm, MATCH_MODES[mode])

    def full(mode):
        r2, ri2, pd2 = ops.prep_rows(res[0]["rgbd"], om, pm)
        return ops.match_fwd(r2, ri2, pd2, cols, aux, None, obj_id, GAMMA, pm, MATCH_MODES[mode])
    fg_ms = {}
    for name, fn in (("compacted_soft", lambda: compacted("soft")), ("full_soft", lambda: full("soft")),
                     ("compacted_argmax", lambda: compacted("argmax")), ("full_argmax", lambda: full("argmax"))):
        for _ in range(3):
            fn()
        va.record()
        for _ in range(10):
            fn()
        vb.record()
        torch.cuda.synchronize()
        fg_ms[name] = va.elapsed_time(vb) / 10

    # ---- end to end through the public API with HOST buffers (H2D + D2H inside the timed region)
    # pipeline.FrameStream: pinned host inputs -> H2D -> prep + match + kNN -> D2H of every output into pinned host
    # buffers, three batches in flight (the host reads batch i - 2 while batch i uploads and batch i - 1 computes) so
    # that the H2D copies, which bound this path (54 MB per step over PCIe), run back to back.
    # ONE packed H2D (bf16 descriptors + the flat point buffer) and ONE packed D2H per batch; with several ranks every
    # batch's matcher records are also all-gathered (FrameStream does it on a side stream).
    from gadm_b200.pipeline import FrameStream
    bank = matching.ModelBank(res[0]["mesh"], xyz)
    with near_gpu(torch, local) as ng:              # pinned pages next to the GPU (when the topology is visible)
        fs = FrameStream(bank, pyr, FRAMES, D, N_PTS, obj_id=obj_id, gamma=GAMMA, mode="soft", depth=3)
        hbatch = [fs.host_batch().fill(h["rgbd"], h["cld"], h["sr"]) for h in host]
    numa = ng.node
    h2d, d2h = fs.h2d_bytes, fs.d2h_bytes

    def e2e_run(n):
        check = 0
        for s in range(n):
            tk = fs.submit(hbatch[s % ROT])
            if tk >= 2:
                out = fs.result(tk - 2)                       # host-side read of an earlier batch's results
                check += int(out["idx"][0, 0]) + int(out["knn"][0])
        for tk in range(max(0, fs.n_submitted - 2), fs.n_submitted):   # every batch's result is read on the host
            out = fs.result(tk)
            check += int(out["idx"][0, 0]) + int(out["knn"][0])
        return check

    e2e_run(max(8, args.warmup))            # every slot of the 3-deep pipeline reused at least twice before timing
    sync_all()
    e2e_steps = max(3, args.steps)
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    sync_all()
    e2e_s = time.perf_counter() - t0

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms_total, e2e_s, match_ms, knn_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s, match_ms, knn_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        frames = FRAMES * world * args.steps
        value = frames / (ms_total * 1e-3)
        flop_per_launch = 2.0 * N_PTS * M_VERTS * D * FRAMES
        match_prof, knn_prof = profile_metrics("match_pair_kernel"), profile_metrics("knn_grid_kernel")
        achieved = flop_per_launch / (match_ms * 1e-3) / 1e12
        peak = pk.get("bf16_tflops", 1590.0)
        launches = 3 + pyr_launches(pyr)
        line = {
            "metric": "frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": shared_config(),
            "impl_config": {"operand_mode": "bf16 operands, fp32 accumulate", "parallelism": f"frames sharded x{world}",
                            "l2": f"rotating {ROT} resident input batches (~{ROT * 85} MB > 126 MB L2)",
                            "streams": "prep + match on the main stream, kNN pyramid on a second stream"
                            if not args.no_overlap else "one stream",
                            "collective": ("all_gather of the matcher records on a side stream, in the device-resident "
                                           "and in the end-to-end loop") if world > 1 else "none",
                            "e2e_host_buffers": "one packed pinned buffer per batch: bf16 descriptors + fp32 points in, "
                                                "int32/fp32 records + uint16 kNN indices out",
                            "host_threads_per_rank": torch.get_num_threads(), "numa_node": numa},
            "breakdown_ms": {"match_kernel": match_ms, "knn_pyramid": knn_ms,
                             "note": ("serial: prep, match, kNN" if args.no_overlap else
                                      "the kNN pyramid runs on a second stream under the matcher; "
                                      "its duration is measured on that stream and overlaps match_kernel")},
            "roofline": {"kernel": "match_pair_kernel<soft> (tcgen05 fused similarity + softmax + argmax + soft coordinates)", "bound": "tensor",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         # the burst figure: the timed region is 20 steps (16 ms) at full clocks (see "clocks"); against
                         # the 4-second sustained figure of the same file the fraction would be higher
                         "peak_kind": f"{pk_kind} bf16 burst", "flop_per_launch": flop_per_launch,
                         "peak_sustained": pk.get("bf16_tflops_sustained"),
                         "frac_of_sustained_peak": (achieved / pk["bf16_tflops_sustained"]
                                                    if pk.get("bf16_tflops_sustained") else None),
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch of THIS kernel, read from the
                         # committed ncu capture by kernel name (null if the capture does not hold it); the
                         # algorithmic bytes are 45 MB of operands + outputs
                         "traffic": (match_prof or {}).get("dram_bytes"),
                         "traffic_source": (match_prof or {}).get("source"),
                         # the pipe that caps a SOFT kernel below the tensor peak: one fp32 MUFU.EX2 per score,
                         # 16.3 results / clk / SM measured (profiles/r2c_probe_pipe_mix.txt), at the SM clock of this run
                         "xu_pipe": xu_ceiling(match_ms, clocks, flop_per_launch, peak)},
            "variants": {"match_kernel_argmax_only_ms": argmax_ms,
                         "match_kernel_argmax_only_frac": flop_per_launch / (argmax_ms * 1e-3) / 1e12 / peak,
                         "match_kernel_argmax_bf16n_ms": pruned_ms,
                         "match_kernel_argmax_bf16n_frac": flop_per_launch / (pruned_ms * 1e-3) / 1e12 / peak,
                         "match_kernel_argmax_unit_ms": unit_ms,
                         "match_kernel_argmax_unit_frac": flop_per_launch / (unit_ms * 1e-3) / 1e12 / peak,
                         "foreground_20pct_ms": fg_ms,
                         "foreground_note": "prep + match of one 8-frame batch whose segmentation mask keeps 20 % of the "
                                            "rows: compacted = compact_rows + prep_rows_sel + match_fwd_sel (results "
                                            "scattered back), full = every row matched, mask applied at the store",
                         "note": "argmax_only = evaluator.py:89-93 exactly (match_alt_kernel); argmax_bf16n = the same, exact, on columns "
                                 "normalised before the bf16 rounding (chunk pruning); argmax_unit = same search "
                                 "on bf16n operands without column scales (index may differ only below a 2^-7 |score| "
                                 "margin); the timed step above uses the SOFT kernel"},
            "knn": {"algorithmic_bytes_per_step": pyr.algorithmic_bytes * FRAMES,
                    "achieved_gbs": pyr.algorithmic_bytes * FRAMES / (knn_ms * 1e-3) / 1e9,
                    "hbm_peak_gbs": pk.get("hbm_gbs"), "queries_per_step": pyr.n_queries * FRAMES,
                    # what actually bounds it (exact search at this size is instruction-issue bound):
                    # issue-slot and pipe utilisation of the query kernel from the committed ncu capture
                    "issue_active_pct": (knn_prof or {}).get("issue_active_pct"),
                    "alu_pipe_pct": (knn_prof or {}).get("alu_pipe_pct"),
                    "fma_pipe_pct": (knn_prof or {}).get("fma_pipe_pct"),
                    "warp_instructions": (knn_prof or {}).get("inst_executed"),
                    "profile_source": (knn_prof or {}).get("source")},
            "e2e": {"value": FRAMES * world * e2e_steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": launches * args.steps,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            n = FRAMES                                           # one frame per kNN worker thread, as the reference arm
            dt, kind, cores = cpu_reference_frames(n, 1000)      # warm
            dt, kind, cores = cpu_reference_frames(n, 1001)
            dts, _, _ = cpu_reference_frames(n, 1001, with_soft=True)
            line["cpu_baseline"] = {
                "value": n / dt, "unit": "frames/s", "cores": cores, "kind": kind,
                "value_with_soft_extension": n / dts,
                "sample": (f"{n} frames of the same workload: torch-CPU fp32 matching head (normalize, normalize, matmul, "
                           f"max; evaluator.py:89-93) on {cores} threads + the 22-call nanoflann schedule per frame "
                           f"(one frame per worker thread)")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()



# ------------------------------------------------------------------------------------------------ config 3 (YCB-V)
YCBV_FRAMES, YCBV_OBJ, YCBV_MAX_INST = 256, 21, 6
YCBV_WORKLOAD = ("ycbv_256: ONE job of 256 frames x U{1..6} instances (seed 3000), 21-object model bank, 12800 pts x "
                 "8192 verts, d=128, full matching (soft) per instance + 22-call kNN pyramid per frame")


def run_ycbv(args):
    """BASELINE.json configs[2] / SURVEY 8(d) config 3, 8(e): one fixed 256-frame job, frames split over the ranks with
    sharding.frame_range (or sharding.balanced_assignment by instance count with --balance), the 21-object bank
    replicated, every rank's matcher records gathered with one NCCL all_gather per step.  STRONG scaling: the job is the
    same for every N.  A step = the whole job.  Reports frames/s and instances/s."""
    import torch
    import torch.distributed as dist
    import gadm_b200  # noqa: F401
    from gadm_b200 import ops, sharding, synth
    from gadm_b200._lib import MATCH_MODES, OPERAND_MODES, PAD_MODES
    from gadm_b200.knn import KnnPyramid

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    pk, pk_kind = peaks()

    # ---- the job (identical on every rank: seeded)
    g = torch.Generator().manual_seed(3000)
    n_inst = torch.randint(1, YCBV_MAX_INST + 1, (YCBV_FRAMES,), generator=g).tolist()
    inst_obj = [torch.randint(0, YCBV_OBJ, (k,), generator=g).tolist() for k in n_inst]
    total_inst = sum(n_inst)
    if args.balance:
        my_frames = sharding.balanced_assignment(n_inst, world)[rank]
    else:
        lo, hi = sharding.frame_range(YCBV_FRAMES, rank, world)
        my_frames = list(range(lo, hi))
    my_obj = [o for f in my_frames for o in inst_obj[f]]
    n_local = len(my_obj)
    counts = [sum(n_inst[f] for f in (sharding.balanced_assignment(n_inst, world)[r] if args.balance
                                      else range(*sharding.frame_range(YCBV_FRAMES, r, world))))
              for r in range(world)]
    cap = max(counts)                                  # instances of the most loaded rank (all_gather pads to it)

    # ---- synthetic operands: a pool of POOL distinct instance descriptor sets (bf16, > L2) used cyclically, the bank
    CH, POOL, KB_FR = 32, 64, 8
    rgbd, mesh, _ = synth.descriptors(POOL, N_PTS, M_VERTS, D, n_obj=YCBV_OBJ, regime="planted", seed=3000 + rank)
    pool = rgbd.to(dev).to(torch.bfloat16)             # exact: the generator emits bf16-representable values
    host_pool = rgbd.to(torch.bfloat16).pin_memory()
    xyz = synth.model_bank_xyz(YCBV_OBJ, M_VERTS, diameters=synth.YCBV_DIAMETERS).to(dev)
    om, pm, mm = OPERAND_MODES["bf16"], PAD_MODES["none"], MATCH_MODES["soft"]
    cols, aux = ops.prep_model(mesh.to(dev), xyz, om)  # the bank is prepared once (replicated on every rank)
    obj_all = torch.tensor(my_obj, dtype=torch.int32, device=dev)
    pyr = KnnPyramid(N_PTS, {s: (IN_SIZE // s) ** 2 for s in (2, 4, 8)}, KB_FR)
    ws_bytes = ops._lib.load().gadm_knn3d_workspace_bytes(pyr.jobs, len(pyr.jobs), ops.KNN_ALGOS["auto"])
    pyr.workspace = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
    cld, sr = synth.frame_batch(KB_FR, IN_SIZE, N_PTS, seed=3000 + rank)
    pts = pyr.pack(cld.to(dev), {s: sr[s].to(dev) for s in (2, 4, 8)})
    host_pts = pts.cpu().pin_memory()
    n_fr_calls = (len(my_frames) + KB_FR - 1) // KB_FR
    rec = torch.zeros((cap, N_PTS, 6), dtype=torch.int32, device=dev)
    gathered = torch.empty((world, cap, N_PTS, 6), dtype=torch.int32, device=dev) if world > 1 else None
    knn_stream, side = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    coll_done = torch.cuda.Event()

    def chunks():
        for c0 in range(0, n_local, CH):
            n = min(CH, n_local - c0)
            p0 = c0 % POOL
            if p0 + n > POOL:                          # keep the pool slice contiguous (a view, no gather copy)
                p0 = 0
            yield c0, n, p0

    def job(from_host=False, slots=None):
        """One pass over this rank's share of the 256-frame job."""
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event(); fork.record(main)
        with torch.cuda.stream(knn_stream):            # the pyramids do not depend on the matcher
            knn_stream.wait_event(fork)
            for _ in range(n_fr_calls):
                if from_host:
                    slots["pts"].copy_(host_pts, non_blocking=True)
                    pyr.run_packed(slots["pts"], out=slots["knn"])
                    slots["knn_host"].copy_(slots["knn"], non_blocking=True)
                else:
                    pyr.run_packed(pts)
            join = torch.cuda.Event(); join.record(knn_stream)
        main.wait_event(coll_done)                     # the previous step's all_gather has read `rec`
        work = list(chunks())
        if from_host:                                  # chunk i + 1 crosses the bus while chunk i is matched
            def upload(i):
                _, n_, p0_ = work[i]
                buf = slots["rgbd"][i & 1]
                with torch.cuda.stream(slots["h2d"]):
                    slots["h2d"].wait_event(slots["free"][i & 1])
                    buf[:n_].copy_(host_pool[p0_:p0_ + n_], non_blocking=True)
                    slots["ready"][i & 1].record(slots["h2d"])
            upload(0)
        for i, (c0, n, p0) in enumerate(work):
            if from_host:
                if i + 1 < len(work):
                    upload(i + 1)
                main.wait_event(slots["ready"][i & 1])
                src = slots["rgbd"][i & 1][:n]
            else:
                src = pool[p0:p0 + n]
            rows, rinv, pad = ops.prep_rows(src, om, pm)
            if from_host:
                slots["free"][i & 1].record(main)      # prep_rows has consumed the staging buffer
            out = ops.match_fwd(rows, rinv, pad, cols, aux, None, obj_all[c0:c0 + n], GAMMA, pm, mm)
            ops.pack_match_outputs(out[0], out[1], out[2], out[3], rec[c0:c0 + n])
        if from_host:
            slots["rec_host"][:n_local].copy_(rec[:n_local], non_blocking=True)
        main.wait_event(join)
        if world > 1:
            done = torch.cuda.Event(); done.record(main)
            with torch.cuda.stream(side):
                side.wait_event(done)
                dist.all_gather_into_tensor(gathered.view(-1), rec.view(-1))
                coll_done.record(side)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(n_steps, **kw):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n_steps):
            job(**kw)
        e1.record()
        torch.cuda.current_stream().wait_stream(side)
        sync_all()
        return e0.elapsed_time(e1), time.perf_counter() - t0

    for _ in range(max(1, min(args.warmup, 3))):
        job()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ms_total, _ = timed(args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # end to end: every instance's descriptors and every frame's points come from pinned host memory, every result goes
    # back to pinned host memory, inside the timed region
    slots = {"rgbd": [torch.empty((CH, D, N_PTS), dtype=torch.bfloat16, device=dev) for _ in range(2)],
             "h2d": torch.cuda.Stream(device=dev), "ready": [torch.cuda.Event(), torch.cuda.Event()],
             "free": [torch.cuda.Event(), torch.cuda.Event()], "pts": torch.empty_like(pts),
             "knn": torch.empty((pyr.out_elems,), dtype=torch.int32, device=dev),
             "knn_host": torch.empty((pyr.out_elems,), dtype=torch.int32).pin_memory(),
             "rec_host": torch.empty((cap, N_PTS, 6), dtype=torch.int32).pin_memory()}
    job(from_host=True, slots=slots)
    e2e_steps = max(2, args.steps // 4)
    _, e2e_s = timed(e2e_steps, from_host=True, slots=slots)
    if world > 1:
        t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = [float(x) for x in t.tolist()]
    if rank == 0:
        sec = ms_total * 1e-3
        h2d = n_local * D * N_PTS * 2 + n_fr_calls * host_pts.numel() * 4
        d2h = n_local * N_PTS * 24 + n_fr_calls * pyr.out_elems * 4
        line = {"metric": "frames/s", "value": YCBV_FRAMES * args.steps / sec, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "instances_per_s": total_inst * args.steps / sec,
                "config": {"workload": YCBV_WORKLOAD, "frames": YCBV_FRAMES, "instances": total_inst,
                           "n_obj": YCBV_OBJ, "gamma": GAMMA},
                "impl_config": {"sharding": "balanced_assignment (greedy by instance count)" if args.balance
                                else "frame_range (contiguous blocks of frames)",
                                "instances_per_rank": counts, "chunk": CH,
                                "l2": f"pool of {POOL} distinct instance descriptor sets (~{POOL * 3.3:.0f} MB bf16)",
                                "collective": "one all_gather of the padded matcher records per step" if world > 1 else "none"},
                "roofline": {"kernel": "match_pair_kernel<soft>", "bound": "tensor", "unit": "TFLOP/s",
                             "achieved": 2.0 * N_PTS * M_VERTS * D * max(counts) * args.steps / sec / 1e12,
                             "peak": pk.get("bf16_tflops", 1590.0),
                             "frac": 2.0 * N_PTS * M_VERTS * D * max(counts) * args.steps / sec / 1e12 / pk.get("bf16_tflops", 1590.0),
                             "peak_kind": f"{pk_kind} bf16 burst", "traffic": None,
                             "note": "whole step of the most loaded rank (matcher + prep + kNN), not the kernel alone"},
                "e2e": {"value": YCBV_FRAMES * e2e_steps / e2e_s, "unit": "frames/s",
                        "instances_per_s": total_inst * e2e_steps / e2e_s,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                        "note": "per rank 0; wall clock, max over ranks"},
                "gpu_launches": (3 * ((n_local + CH - 1) // CH) + n_fr_calls * pyr_launches(pyr)) * args.steps,
                "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()

MUFU_PER_CLK_PER_SM = 16.3   # MUFU.EX2 results, tools/probes/pipe_mix_probe.cu on B200 (profiles/r2c_probe_pipe_mix.txt)


def xu_ceiling(match_ms, clocks, flop_per_launch, peak_tflops):
    """What the XU pipe alone needs for the SOFT kernel's exponentials (one per score) at this run's SM clock."""
    mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz")
    if not mhz:
        return None
    scores = float(N_PTS) * M_VERTS * FRAMES
    floor_ms = scores / (MUFU_PER_CLK_PER_SM * 148 * mhz * 1e6) * 1e3
    return {"exp2_per_launch": scores, "results_per_clk_per_sm": MUFU_PER_CLK_PER_SM, "sm_mhz": mhz,
            "floor_ms": floor_ms, "frac_of_xu_peak": floor_ms / match_ms,
            "tensor_frac_if_xu_bound": flop_per_launch / (floor_ms * 1e-3) / 1e12 / peak_tflops}


def pyr_launches(pyr):
    """Kernel launches of one gadm_knn3d call (knn3d.cu): 4 grid-build kernels when any job uses the grid
    (AUTO: n_support >= 128, g_grid_min_support), then one query kernel per algorithm present."""
    any_grid = any(j.n_support >= 128 for j in pyr.jobs)
    any_brute = any(j.n_support < 128 for j in pyr.jobs)
    return (4 + 1 if any_grid else 0) + (1 if any_brute else 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gadm", choices=["gadm", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="run the kNN pyramid after the matcher on one stream")
    ap.add_argument("--config", default="lmo", choices=["lmo", "ycbv"],
                    help="lmo: BASELINE configs[1] (the metric's config, weak scaling); ycbv: configs[2] (strong scaling)")
    ap.add_argument("--balance", action="store_true", help="ycbv: greedy balance by instance count instead of frame blocks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "gadm" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "ycbv":
        run_ycbv(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
