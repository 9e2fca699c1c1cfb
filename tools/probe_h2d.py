"""Host <-> device copy ceiling of the box, all ranks at once (not a benchmark of this library).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/probe_h2d.py

Every rank copies the end-to-end leg's own transfer sizes (bench.py: 28.0 MB in, 9.4 MB out per 8-frame batch) and a
large 256 MB buffer between pinned host memory and its GPU, first alone (ranks take turns), then all ranks together
after a barrier.  The ratio together / alone is what the host's memory system and PCIe topology leave of one GPU's
copy bandwidth when N ranks stream at once: the ceiling of the e2e scaling efficiency that `bench.py --gpus N` can reach.
Last, the step's own copy pattern: 27.96 MB in and 9.42 MB out concurrently on two streams, every rank at once -- the time
of that pair is the floor of an end-to-end step on this box whatever the kernels do ("e2e_copy_floor").
One JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def copy_gbs(dst, src, stream, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        dst.copy_(src, non_blocking=True)
        stream.synchronize()
        a.record(stream)
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        b.record(stream)
    b.synchronize()
    return src.numel() * src.element_size() * reps / a.elapsed_time(b) / 1e6


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sizes = {"in_28MB": 27_959_296, "out_9MB": 9_420_928, "big_256MB": 256 << 20}
    stream = torch.cuda.Stream(device=dev)
    res = {}
    for name, nbytes in sizes.items():
        host = torch.empty((nbytes,), dtype=torch.uint8).pin_memory()
        devb = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        reps = max(4, int(2e9 // nbytes))
        alone_h2d = alone_d2h = 0.0
        for turn in range(world):            # one rank at a time
            if world > 1:
                dist.barrier()
            if turn == rank:
                alone_h2d = copy_gbs(devb, host, stream, reps)
                alone_d2h = copy_gbs(host, devb, stream, reps)
        if world > 1:
            dist.barrier()
        tog_h2d = copy_gbs(devb, host, stream, reps)     # every rank at once
        if world > 1:
            dist.barrier()
        tog_d2h = copy_gbs(host, devb, stream, reps)
        t = torch.tensor([alone_h2d, alone_d2h, tog_h2d, tog_d2h], dtype=torch.float64, device=dev)
        if world > 1:
            allt = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
        else:
            allt = [t]
        m = torch.stack(allt).cpu()
        res[name] = {"h2d_alone_gbs_per_rank": [round(v, 1) for v in m[:, 0].tolist()],
                     "d2h_alone_gbs_per_rank": [round(v, 1) for v in m[:, 1].tolist()],
                     "h2d_together_gbs_per_rank": [round(v, 1) for v in m[:, 2].tolist()],
                     "d2h_together_gbs_per_rank": [round(v, 1) for v in m[:, 3].tolist()],
                     "h2d_together_over_alone": round(float(m[:, 2].sum() / m[:, 0].sum()), 3),
                     "d2h_together_over_alone": round(float(m[:, 3].sum() / m[:, 1].sum()), 3)}
    # the end-to-end step's own copies, both directions at once, every rank at once
    n_in, n_out = 27_959_296, 9_420_928
    h_in = torch.empty((n_in,), dtype=torch.uint8).pin_memory()
    d_in = torch.empty((n_in,), dtype=torch.uint8, device=dev)
    h_out = torch.empty((n_out,), dtype=torch.uint8).pin_memory()
    d_out = torch.empty((n_out,), dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    floors = {}
    for mode in ("alone", "together"):
        ms = 0.0
        for turn in range(world if mode == "alone" else 1):
            if world > 1:
                dist.barrier()
            if mode == "together" or turn == rank:
                reps = 40
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                s_in.wait_event(a); s_out.wait_event(a)
                for _ in range(reps):
                    with torch.cuda.stream(s_in):
                        d_in.copy_(h_in, non_blocking=True)
                    with torch.cuda.stream(s_out):
                        h_out.copy_(d_out, non_blocking=True)
                e1, e2 = torch.cuda.Event(), torch.cuda.Event()
                e1.record(s_in); e2.record(s_out)
                torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
                b.record()
                b.synchronize()
                ms = a.elapsed_time(b) / reps
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            allt = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
        else:
            allt = [t]
        floors[mode] = [round(float(x.item()), 3) for x in allt]
    worst = max(floors["together"])
    res["e2e_copy_floor"] = {"ms_per_step_alone_per_rank": floors["alone"], "ms_per_step_together_per_rank": floors["together"],
                             "frames_per_s_lockstep": round(8 * world / worst * 1e3, 1),
                             "frames_per_s_aggregate": round(sum(8 / t * 1e3 for t in floors["together"]), 1),
                             "note": "8 frames per rank and step, copies only.  lockstep = 8 * ranks / slowest rank's copy "
                                     "time (every rank waits for the slowest each step); aggregate = sum of the ranks' own "
                                     "rates (ranks that finish early leave their bandwidth to the others): a run with equal "
                                     "steps per rank, timed to the last rank, lands between the two"}
    if rank == 0:
        print(json.dumps({"probe": "pinned host <-> device copies, all ranks at once", "n_gpus": world,
                          "host_cores": os.cpu_count(), "results": res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
