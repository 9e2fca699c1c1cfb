"""Host-buffer front end of the whole path for a stream of frame batches.

The reference moves every sample host -> device -> host synchronously (train_lm.py:160-170 uploads the dataset
dict, evaluator.py:87,99 pulls indices and clouds back per frame).  FrameStream keeps `depth` batches in flight:

    pinned host inputs --H2D (copy stream)--> prep_rows + match_fwd + kNN pyramid (compute stream)
                       --D2H (second copy stream)--> pinned host outputs

so the PCIe copies of batch i+1 / i-1 overlap the kernels of batch i.  Every byte still crosses the bus every
batch; nothing is cached between batches.
"""
import torch

from . import ops
from ._lib import MATCH_MODES, OPERAND_MODES, PAD_MODES


class _Slot:
    pass


class FrameStream:
    """bank: matching.ModelBank; pyramid: knn.KnnPyramid (its batch must equal B); obj_id: int32 [B] CUDA or None.

    submit(rgbd, cld, sr) takes HOST tensors (pinned for real overlap): rgbd [B, d, N] fp32, cld [B, N, 3] fp32,
    sr {2|4|8: [B, P_s, 3]} fp32, and returns a ticket; result(ticket) blocks until that batch's outputs are in
    host memory and returns {'idx','max_sim','weight','soft_xyz','knn'} as pinned host tensors ('knn' is the flat
    int32 buffer KnnPyramid.unpack() understands).  The buffers of a ticket are reused `depth` submits later."""

    def __init__(self, bank, pyramid, B, d, N, obj_id=None, gamma=16.0, mode="soft", depth=2):
        self.bank, self.pyr, self.B, self.d, self.N = bank, pyramid, B, d, N
        self.obj_id, self.gamma, self.mode, self.depth = obj_id, float(gamma), mode, depth
        dev = bank.device
        self.dev = dev
        self.h2d = torch.cuda.Stream(device=dev)
        self.d2h = torch.cuda.Stream(device=dev)
        soft = mode == "soft"
        self.slots = []
        for _ in range(depth):
            s = _Slot()
            s.rgbd = torch.empty((B, d, N), dtype=torch.float32, device=dev)
            s.cld = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
            s.sr = {k: torch.empty((B, pyramid.grid_sizes[k], 3), dtype=torch.float32, device=dev)
                    for k in (2, 4, 8)}
            s.out = {"idx": torch.empty((B, N), dtype=torch.int64).pin_memory(),
                     "max_sim": torch.empty((B, N), dtype=torch.float32).pin_memory(),
                     "knn": torch.empty((pyramid.out_elems,), dtype=torch.int32).pin_memory()}
            if soft:
                s.out["weight"] = torch.empty((B, N), dtype=torch.float32).pin_memory()
                s.out["soft_xyz"] = torch.empty((B, N, 3), dtype=torch.float32).pin_memory()
            s.h2d_done = torch.cuda.Event()
            s.compute_done = torch.cuda.Event()
            s.d2h_done = torch.cuda.Event()
            s.used = False
            self.slots.append(s)
        if pyramid.workspace is None:
            need = ops._lib.load().gadm_knn3d_workspace_bytes(pyramid.jobs, len(pyramid.jobs),
                                                              ops.KNN_ALGOS[pyramid.algo])
            pyramid.workspace = torch.empty((max(need, 16),), dtype=torch.uint8, device=dev)
        self.n_submitted = 0
        self.h2d_bytes = 4 * (B * d * N + B * N * 3 + sum(B * pyramid.grid_sizes[k] * 3 for k in (2, 4, 8)))
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.slots[0].out.values())

    def submit(self, rgbd, cld, sr):
        s = self.slots[self.n_submitted % self.depth]
        compute = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.h2d):
            if s.used:
                self.h2d.wait_event(s.compute_done)     # the kernels that read this slot's inputs have finished
            s.rgbd.copy_(rgbd, non_blocking=True)
            s.cld.copy_(cld, non_blocking=True)
            for k in (2, 4, 8):
                s.sr[k].copy_(sr[k], non_blocking=True)
            s.h2d_done.record(self.h2d)
        compute.wait_event(s.h2d_done)
        if s.used:
            # the previous results of this slot are about to be released: order that (and every allocation that
            # may reuse their memory) after the copy that read them.  No record_stream(): blocks released under a
            # recorded foreign stream come back to the caching allocator late and at unpredictable times, and a
            # step that finds none free pays a cudaMalloc (measured: 6.4 k -> 0.6-2.5 k frames/s on some runs).
            compute.wait_event(s.d2h_done)
        om, pm = OPERAND_MODES[self.bank.operand_mode], PAD_MODES["none"]
        rows, rinv, pad = ops.prep_rows(s.rgbd, om, pm)
        outs = ops.match_fwd(rows, rinv, pad, self.bank.cols, self.bank.aux, None, self.obj_id, self.gamma, pm,
                             MATCH_MODES[self.mode])
        knn = self.pyr.run_packed(self.pyr.pack(s.cld, s.sr))
        s.compute_done.record(compute)
        dev_out = {"idx": outs[0], "max_sim": outs[1], "knn": knn}
        if self.mode == "soft":
            dev_out["weight"], dev_out["soft_xyz"] = outs[2], outs[3]
        s.dev_out = dev_out                           # kept alive until the slot is reused (see above)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(s.compute_done)
            for name, t in dev_out.items():
                s.out[name].copy_(t, non_blocking=True)
            s.d2h_done.record(self.d2h)
        s.used = True
        ticket = self.n_submitted
        self.n_submitted += 1
        return ticket

    def result(self, ticket):
        if ticket < self.n_submitted - self.depth or ticket >= self.n_submitted:
            raise ValueError("ticket expired (its buffers were reused) or not submitted yet")
        s = self.slots[ticket % self.depth]
        s.d2h_done.synchronize()
        return s.out
