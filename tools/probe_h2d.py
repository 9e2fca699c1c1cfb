"""Host <-> device copy ceiling of the box, all ranks at once (not a benchmark of this library).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/probe_h2d.py

Every rank copies the end-to-end leg's own transfer sizes (bench.py: 30.1 MB in, 16.9 MB out per 8-frame batch) and a
large 256 MB buffer between pinned host memory and its GPU, first alone (ranks take turns), then all ranks together
after a barrier.  The ratio together / alone is what the host's memory system and PCIe topology leave of one GPU's
copy bandwidth when N ranks stream at once: the ceiling of the e2e scaling efficiency that `bench.py --gpus N` can reach.
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
    sizes = {"in_30MB": 30_105_600, "out_17MB": 16_933_968, "big_256MB": 256 << 20}
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
    if rank == 0:
        print(json.dumps({"probe": "pinned host <-> device copies, all ranks at once", "n_gpus": world,
                          "host_cores": os.cpu_count(), "results": res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
