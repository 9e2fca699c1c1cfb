"""Host-buffer front end of the whole path for a stream of frame batches.

The reference moves every sample host -> device -> host synchronously (train_lm.py:160-170 uploads the dataset
dict tensor by tensor, evaluator.py:87,99 pulls indices and clouds back per frame).  FrameStream keeps `depth`
batches in flight,

    pinned host batch --ONE H2D (copy stream)--> prep_rows + match_fwd (compute stream) | kNN pyramid (its own stream)
                      --ONE D2H (second copy stream)--> pinned host results

so the PCIe copies of batch i+1 / i-1 overlap the kernels of batch i.  Every byte still crosses the bus every batch;
nothing is cached between batches.  What crosses it is packed and as small as the path's contract allows:

  in   [ descriptors bf16 [B, d, N] | points fp32 [B, P, 3] ]  -- the matcher consumes bf16-representable descriptors
       (DESIGN.md "precision contract"), so a bf16 host buffer carries exactly the operand values at half the bytes; the
       points are laid out as KnnPyramid's flat buffer (cloud, then the 1/2, 1/4, 1/8 image grids of every frame)
  out  [ records int32 [B, N, 6] = {idx, max_sim, weight, x, y, z} | kNN indices uint16 [out_elems] ]
       (uint16 while every support cloud of the pyramid has fewer than 65536 points -- the indices are narrowed on the
       device, gadm_pack_indices_u16 -- else int32; KnnPyramid.unpack() accepts either)

With several ranks (torch.distributed initialised) the matcher records of every batch are also all-gathered on a
side stream, so that each rank holds the whole job's correspondences on its device (evaluator.py:240-249 gathers on
the host); `gather=False` skips that.
"""
import torch
import torch.distributed as dist

from . import ops
from ._lib import MATCH_MODES, OPERAND_MODES, PAD_MODES


class _Slot:
    pass


class HostBatch:
    """One batch of inputs in ONE pinned host buffer; .rgbd (bf16 [B, d, N]), .cld (fp32 [B, N, 3]) and .sr[2|4|8]
    (fp32 [B, P_s, 3]) are views into it -- fill them in place."""

    def __init__(self, stream_):
        fs = stream_
        self.flat = torch.empty((fs.in_bytes,), dtype=torch.uint8).pin_memory()
        self.rgbd = self.flat[: fs.rgbd_bytes].view(torch.bfloat16).view(fs.B, fs.d, fs.N)
        pts = self.flat[fs.rgbd_bytes:].view(torch.float32).view(fs.B, fs.pyr.P, 3)
        self.cld = pts[:, : fs.N]
        self.sr = {k: pts[:, fs.pyr.off[k]: fs.pyr.off[k] + fs.pyr.grid_sizes[k]] for k in (2, 4, 8)}

    def fill(self, rgbd, cld, sr):
        self.rgbd.copy_(rgbd)            # fp32 -> bf16: exact for bf16-representable descriptors
        self.cld.copy_(cld)
        for k in (2, 4, 8):
            self.sr[k].copy_(sr[k])
        return self


class FrameStream:
    """bank: matching.ModelBank; pyramid: knn.KnnPyramid (its batch must equal B); obj_id: int32 [B] CUDA or None.

    submit(batch: HostBatch) returns a ticket; result(ticket) blocks until that batch's outputs are in host memory and
    returns {'idx' int32, 'max_sim', 'weight', 'soft_xyz', 'knn'} as views of ONE pinned host buffer ('knn' is the
    flat index buffer KnnPyramid.unpack() understands: uint16, or int32 for clouds of 65536 points and more).  The
    buffers of a ticket are reused `depth` submits later."""

    def __init__(self, bank, pyramid, B, d, N, obj_id=None, gamma=16.0, mode="soft", depth=3, gather=True):
        self.bank, self.pyr, self.B, self.d, self.N = bank, pyramid, B, d, N
        self.obj_id, self.gamma, self.mode, self.depth = obj_id, float(gamma), mode, depth
        dev = bank.device
        self.dev = dev
        self.h2d = torch.cuda.Stream(device=dev)
        self.d2h = torch.cuda.Stream(device=dev)
        self.knn = torch.cuda.Stream(device=dev)        # the pyramid has no data dependence on the matcher
        self.world = dist.get_world_size() if (gather and dist.is_available() and dist.is_initialized()) else 1
        self.coll = torch.cuda.Stream(device=dev) if self.world > 1 else None
        self.rgbd_bytes = B * d * N * 2
        self.in_bytes = self.rgbd_bytes + B * pyramid.P * 3 * 4
        self.rec_elems = B * N * 6
        self.knn_u16 = max(j.n_support for j in pyramid.jobs) < 65536
        knn_words = (pyramid.out_elems + 1) // 2 if self.knn_u16 else pyramid.out_elems     # int32 words on the bus
        self.out_elems = self.rec_elems + knn_words
        self.slots = []
        for _ in range(depth):
            s = _Slot()
            s.dev_in = torch.empty((self.in_bytes,), dtype=torch.uint8, device=dev)
            s.rgbd = s.dev_in[: self.rgbd_bytes].view(torch.bfloat16).view(B, d, N)
            s.pts = s.dev_in[self.rgbd_bytes:].view(torch.float32).view(B * pyramid.P, 3)
            s.dev_out = torch.empty((self.out_elems,), dtype=torch.int32, device=dev)
            s.host_out = torch.empty((self.out_elems,), dtype=torch.int32).pin_memory()
            rec = s.host_out[: self.rec_elems].view(B, N, 6)
            knn_host = s.host_out[self.rec_elems:]
            if self.knn_u16:
                knn_host = knn_host.view(torch.uint16)[: pyramid.out_elems]
                s.knn32 = torch.empty((pyramid.out_elems,), dtype=torch.int32, device=dev)
                s.knn16 = s.dev_out[self.rec_elems:].view(torch.uint16)[: pyramid.out_elems]
            s.out = {"idx": rec[..., 0], "max_sim": rec[..., 1].view(torch.float32),
                     "weight": rec[..., 2].view(torch.float32), "soft_xyz": rec[..., 3:6].view(torch.float32),
                     "knn": knn_host}
            s.gathered = (torch.empty((self.world, self.rec_elems), dtype=torch.int32, device=dev)
                          if self.world > 1 else None)
            s.h2d_done = torch.cuda.Event()
            s.knn_done = torch.cuda.Event()
            s.compute_done = torch.cuda.Event()
            s.d2h_done = torch.cuda.Event()
            s.coll_done = torch.cuda.Event()
            s.used = False
            self.slots.append(s)
        if pyramid.workspace is None:
            need = ops._lib.load().gadm_knn3d_workspace_bytes(pyramid.jobs, len(pyramid.jobs),
                                                              ops.KNN_ALGOS[pyramid.algo])
            pyramid.workspace = torch.empty((max(need, 16),), dtype=torch.uint8, device=dev)
        self.n_submitted = 0
        self.h2d_bytes = self.in_bytes
        self.d2h_bytes = self.out_elems * 4

    def host_batch(self):
        return HostBatch(self)

    def submit(self, batch):
        s = self.slots[self.n_submitted % self.depth]
        compute = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.h2d):
            if s.used:
                self.h2d.wait_event(s.compute_done)     # the kernels that read this slot's inputs have finished
            s.dev_in.copy_(batch.flat, non_blocking=True)
            s.h2d_done.record(self.h2d)
        compute.wait_event(s.h2d_done)
        self.knn.wait_event(s.h2d_done)
        if s.used:
            compute.wait_event(s.d2h_done)              # the copy (and the collective) that read this slot's outputs
            self.knn.wait_event(s.d2h_done)
            if self.coll is not None:
                compute.wait_event(s.coll_done)
        om, pm = OPERAND_MODES[self.bank.operand_mode], PAD_MODES["none"]
        rows, rinv, pad = ops.prep_rows(s.rgbd, om, pm)
        outs = ops.match_fwd(rows, rinv, pad, self.bank.cols, self.bank.aux, None, self.obj_id, self.gamma, pm,
                             MATCH_MODES[self.mode])
        rec = s.dev_out[: self.rec_elems]
        ops.pack_match_outputs(outs[0], outs[1], outs[2], outs[3], rec)
        with torch.cuda.stream(self.knn):
            if self.knn_u16:
                self.pyr.run_packed(s.pts, out=s.knn32)
                ops.pack_indices_u16(s.knn32, s.knn16)
            else:
                self.pyr.run_packed(s.pts, out=s.dev_out[self.rec_elems:])
            s.knn_done.record(self.knn)
        compute.wait_event(s.knn_done)
        s.compute_done.record(compute)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(s.compute_done)
            s.host_out.copy_(s.dev_out, non_blocking=True)
            s.d2h_done.record(self.d2h)
        if self.coll is not None:
            with torch.cuda.stream(self.coll):
                self.coll.wait_event(s.compute_done)
                dist.all_gather_into_tensor(s.gathered.view(-1), rec)
                s.coll_done.record(self.coll)
        s.used = True
        ticket = self.n_submitted
        self.n_submitted += 1
        return ticket

    def result(self, ticket):
        if ticket < self.n_submitted - self.depth or ticket >= self.n_submitted:
            raise ValueError("ticket expired (its buffers were reused) or not submitted yet")
        s = self.slots[ticket % self.depth]
        s.d2h_done.synchronize()
        if self.coll is not None:
            s.coll_done.synchronize()
        return s.out

    def gathered(self, ticket):
        """Device tensor [world, B, N, 6] int32: every rank's matcher records of that batch (None on one rank)."""
        s = self.slots[ticket % self.depth]
        return None if s.gathered is None else s.gathered.view(self.world, self.B, self.N, 6)
