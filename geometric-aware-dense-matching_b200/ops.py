"""PyTorch custom ops (`torch.ops.gadm.*`) over the C ABI.  CUDA only; shape inference via register_fake.

Tensors are validated here (dtype / contiguity / device); the C side validates sizes and alignment and
returns error codes that become GadmError.  All ops enqueue on torch's current stream."""
import ctypes

import torch

from . import _lib
from ._lib import KnnJob, PAD_MODES, OPERAND_MODES, MATCH_MODES, KNN_ALGOS  # noqa: F401


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need(t, dtype, name):
    if not t.is_cuda:
        raise _lib.GadmError(f"{name}: expected a CUDA tensor (libgadm has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")


def _lib_for(t):
    return _lib.ensure_init(t.device.index if t.device.index is not None else torch.cuda.current_device())


# --------------------------------------------------------------------------------------------- matching
@torch.library.custom_op("gadm::prep_rows", mutates_args=(), device_types="cuda")
def prep_rows(feat: torch.Tensor, operand_mode: int, pad_mode: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    # fp32 descriptors (the reference's end_points['rgbd']) or descriptors that are already bf16 (same outputs as for
    # the fp32 values they represent; half the bytes)
    _need(feat, torch.bfloat16 if feat.dtype == torch.bfloat16 else torch.float32, "feat")
    B, d, N = feat.shape
    lib = _lib_for(feat)
    kp = lib.gadm_operand_k(d, operand_mode)
    _lib.check(min(kp, 0), "gadm_operand_k")
    rows = torch.empty((B, N, kp), dtype=torch.bfloat16, device=feat.device)
    rinv = torch.empty((B, N), dtype=torch.float32, device=feat.device)
    pad_sim = torch.empty((B, N) if pad_mode else (0,), dtype=torch.float32, device=feat.device)
    fn, name = ((lib.gadm_prep_rows_bf16, "gadm_prep_rows_bf16") if feat.dtype == torch.bfloat16
                else (lib.gadm_prep_rows, "gadm_prep_rows"))
    with torch.cuda.device(feat.device):
        _lib.check(fn(_ptr(feat), B, d, N, operand_mode, pad_mode, _ptr(rows), _ptr(rinv),
                      _ptr(pad_sim) if pad_mode else None, _stream()), name)
    return rows, rinv, pad_sim


@prep_rows.register_fake
def _(feat, operand_mode, pad_mode):
    B, d, N = feat.shape
    kp = d * (3 if operand_mode == 1 else 1)
    return (feat.new_empty((B, N, kp), dtype=torch.bfloat16), feat.new_empty((B, N), dtype=torch.float32),
            feat.new_empty((B, N) if pad_mode else (0,), dtype=torch.float32))


@torch.library.custom_op("gadm::prep_model", mutates_args=(), device_types="cuda")
def prep_model(mesh: torch.Tensor, model_xyz: torch.Tensor, operand_mode: int) -> tuple[torch.Tensor, torch.Tensor]:
    _need(mesh, torch.float32, "mesh")
    _need(model_xyz, torch.float32, "model_xyz")
    n_obj, d, M = mesh.shape
    if tuple(model_xyz.shape) != (n_obj, M, 3):
        raise ValueError(f"model_xyz must be [{n_obj}, {M}, 3], got {tuple(model_xyz.shape)}")
    lib = _lib_for(mesh)
    kp = lib.gadm_operand_k(d, operand_mode)
    _lib.check(min(kp, 0), "gadm_operand_k")
    cols = torch.empty((n_obj, M, kp), dtype=torch.bfloat16, device=mesh.device)
    aux = torch.empty((lib.gadm_aux_floats(n_obj, M),), dtype=torch.float32, device=mesh.device)
    with torch.cuda.device(mesh.device):
        _lib.check(lib.gadm_prep_model(_ptr(mesh), _ptr(model_xyz), n_obj, d, M, operand_mode, _ptr(cols), _ptr(aux),
                                       _stream()), "gadm_prep_model")
    return cols, aux


@prep_model.register_fake
def _(mesh, model_xyz, operand_mode):
    n_obj, d, M = mesh.shape
    kp = d * (3 if operand_mode == 1 else 1)
    return (mesh.new_empty((n_obj, M, kp), dtype=torch.bfloat16),
            mesh.new_empty((n_obj * M * 7,)))     # gadm_aux_floats


@torch.library.custom_op("gadm::compact_rows", mutates_args=(), device_types="cuda")
def compact_rows(mask: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """mask [B, N] uint8 -> (pos int32 [B, N], row_map int32 [B, N], n_sel int32 [B]); evaluator.py:82-88 as a map."""
    _need(mask, torch.uint8, "mask")
    B, N = mask.shape
    pos = torch.empty((B, N), dtype=torch.int32, device=mask.device)
    row_map = torch.zeros((B, N), dtype=torch.int32, device=mask.device)
    n_sel = torch.empty((B,), dtype=torch.int32, device=mask.device)
    lib = _lib_for(mask)
    with torch.cuda.device(mask.device):
        _lib.check(lib.gadm_compact_rows(_ptr(mask), B, N, _ptr(pos), _ptr(row_map), _ptr(n_sel), _stream()),
                   "gadm_compact_rows")
    return pos, row_map, n_sel


@compact_rows.register_fake
def _(mask):
    B, N = mask.shape
    return (mask.new_empty((B, N), dtype=torch.int32), mask.new_empty((B, N), dtype=torch.int32),
            mask.new_empty((B,), dtype=torch.int32))


@torch.library.custom_op("gadm::prep_rows_sel", mutates_args=(), device_types="cuda")
def prep_rows_sel(feat: torch.Tensor, pos: torch.Tensor, operand_mode: int,
                  pad_mode: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """prep_rows with row compaction: point n of frame b becomes row pos[b, n] (or is skipped when pos < 0).  The
    outputs keep the capacity [B, N, .]; rows past the frame's count are left unwritten."""
    _need(feat, torch.bfloat16 if feat.dtype == torch.bfloat16 else torch.float32, "feat")
    _need(pos, torch.int32, "pos")
    B, d, N = feat.shape
    if tuple(pos.shape) != (B, N):
        raise ValueError("pos must be [B, N]")
    lib = _lib_for(feat)
    kp = lib.gadm_operand_k(d, operand_mode)
    _lib.check(min(kp, 0), "gadm_operand_k")
    rows = torch.zeros((B, N, kp), dtype=torch.bfloat16, device=feat.device)
    rinv = torch.zeros((B, N), dtype=torch.float32, device=feat.device)
    pad_sim = torch.zeros((B, N) if pad_mode else (0,), dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        _lib.check(lib.gadm_prep_rows_sel(_ptr(feat), int(feat.dtype == torch.bfloat16), _ptr(pos), B, d, N,
                                          operand_mode, pad_mode, _ptr(rows), _ptr(rinv),
                                          _ptr(pad_sim) if pad_mode else None, _stream()), "gadm_prep_rows_sel")
    return rows, rinv, pad_sim


@prep_rows_sel.register_fake
def _(feat, pos, operand_mode, pad_mode):
    B, d, N = feat.shape
    kp = d * (3 if operand_mode == 1 else 1)
    return (feat.new_empty((B, N, kp), dtype=torch.bfloat16), feat.new_empty((B, N), dtype=torch.float32),
            feat.new_empty((B, N) if pad_mode else (0,), dtype=torch.float32))


@torch.library.custom_op("gadm::pack_match_outputs", mutates_args=("out",), device_types="cuda")
def pack_match_outputs(idx: torch.Tensor, max_sim: torch.Tensor, weight: torch.Tensor | None,
                       soft_xyz: torch.Tensor | None, out: torch.Tensor) -> None:
    """{int32 idx, max_sim, weight, x, y, z} records of every scene point into `out` (int32 [..., 6])."""
    _need(idx, torch.int64, "idx"); _need(max_sim, torch.float32, "max_sim"); _need(out, torch.int32, "out")
    n = idx.numel()
    if out.numel() != 6 * n or max_sim.numel() != n:
        raise ValueError("pack_match_outputs: out must hold 6 words per scene point")
    if weight is not None and weight.numel() == 0:
        weight = None
    if soft_xyz is not None and soft_xyz.numel() == 0:
        soft_xyz = None
    lib = _lib_for(idx)
    with torch.cuda.device(idx.device):
        _lib.check(lib.gadm_pack_match_outputs(_ptr(idx), _ptr(max_sim), _ptr(weight), _ptr(soft_xyz), n, _ptr(out),
                                               _stream()), "gadm_pack_match_outputs")


@torch.library.custom_op("gadm::pack_indices_u16", mutates_args=("out",), device_types="cuda")
def pack_indices_u16(idx: torch.Tensor, out: torch.Tensor) -> None:
    """int32 neighbour indices -> uint16 (every support cloud < 65536 points): half the bytes on the bus."""
    _need(idx, torch.int32, "idx"); _need(out, torch.uint16, "out")
    if out.numel() != idx.numel():
        raise ValueError("pack_indices_u16: out must hold one uint16 per index")
    lib = _lib_for(idx)
    with torch.cuda.device(idx.device):
        _lib.check(lib.gadm_pack_indices_u16(_ptr(idx), idx.numel(), _ptr(out), _stream()), "gadm_pack_indices_u16")


_MATCH_WS = {}


def _match_workspace(lib, dev):
    """Scratch of the alternating / fragment-layout match kernels (argmax stash, one 64 KB slot per SM): one
    persistent buffer per (device, stream) -- launches on a stream are ordered, launches on different streams get
    different buffers, and the timed loop never touches the allocator (a per-call torch.empty next to
    record_stream()-held NCCL buffers made the caching allocator cudaMalloc inside the step)."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(dev).cuda_stream)
    ws = _MATCH_WS.get(key)
    if ws is None:
        if len(_MATCH_WS) >= 64:          # streams come and go (pools, graphs): keep the cache bounded
            _MATCH_WS.clear()
        ws = torch.empty((int(lib.gadm_match_workspace_bytes()),), dtype=torch.uint8, device=dev)
        _MATCH_WS[key] = ws
    return ws


@torch.library.custom_op("gadm::match_fwd", mutates_args=(), device_types="cuda")
def match_fwd(rows: torch.Tensor, rinv: torch.Tensor, pad_sim: torch.Tensor, cols: torch.Tensor, aux: torch.Tensor,
              mask: torch.Tensor | None, obj_id: torch.Tensor | None, gamma: float, pad_mode: int,
              mode: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    _need(rows, torch.bfloat16, "rows"); _need(cols, torch.bfloat16, "cols")
    _need(rinv, torch.float32, "rinv"); _need(aux, torch.float32, "aux")
    B, N, kp = rows.shape
    n_obj, M, kp2 = cols.shape
    if kp != kp2:
        raise ValueError(f"operand K mismatch: rows {kp} vs cols {kp2}")
    if mask is not None:
        _need(mask, torch.uint8, "mask")
        if tuple(mask.shape) != (B, N):
            raise ValueError("mask must be [B, N]")
    if obj_id is not None:
        _need(obj_id, torch.int32, "obj_id")
        if tuple(obj_id.shape) != (B,):
            raise ValueError("obj_id must be [B]")
    dev = rows.device
    idx = torch.empty((B, N), dtype=torch.int64, device=dev)
    max_sim = torch.empty((B, N), dtype=torch.float32, device=dev)
    soft = mode == 1
    weight = torch.empty((B, N) if soft else (0,), dtype=torch.float32, device=dev)
    soft_xyz = torch.empty((B, N, 3) if soft else (0,), dtype=torch.float32, device=dev)
    lib = _lib_for(rows)
    with torch.cuda.device(dev):
        ws = _match_workspace(lib, dev)
        _lib.check(lib.gadm_match_fwd(_ptr(rows), _ptr(rinv), _ptr(pad_sim) if pad_mode else None, _ptr(cols),
                                      _ptr(aux), _ptr(mask), _ptr(obj_id), B, N, M, kp, n_obj, float(gamma),
                                      pad_mode, mode, _ptr(idx), _ptr(max_sim), _ptr(weight) if soft else None,
                                      _ptr(soft_xyz) if soft else None, _ptr(ws), ws.numel(), _stream()),
                   "gadm_match_fwd")
    return idx, max_sim, weight, soft_xyz


@torch.library.custom_op("gadm::match_fwd_sel", mutates_args=(), device_types="cuda")
def match_fwd_sel(rows: torch.Tensor, rinv: torch.Tensor, pad_sim: torch.Tensor, cols: torch.Tensor, aux: torch.Tensor,
                  n_rows: torch.Tensor, row_map: torch.Tensor | None, obj_id: torch.Tensor | None, gamma: float,
                  pad_mode: int, mode: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """match_fwd over compacted frames (n_rows[b] rows each).  row_map given: results are scattered back to the rows'
    original positions (idx = -1, zeros elsewhere); row_map None: results stay in compacted order (rows past n_rows[b]:
    idx = -1, zeros) -- exactly the ordering of evaluator.py:88-93."""
    _need(rows, torch.bfloat16, "rows"); _need(cols, torch.bfloat16, "cols")
    _need(rinv, torch.float32, "rinv"); _need(aux, torch.float32, "aux"); _need(n_rows, torch.int32, "n_rows")
    B, N, kp = rows.shape
    n_obj, M, kp2 = cols.shape
    if kp != kp2:
        raise ValueError(f"operand K mismatch: rows {kp} vs cols {kp2}")
    if row_map is not None:
        _need(row_map, torch.int32, "row_map")
    if obj_id is not None:
        _need(obj_id, torch.int32, "obj_id")
    dev = rows.device
    idx = torch.full((B, N), -1, dtype=torch.int64, device=dev)
    max_sim = torch.zeros((B, N), dtype=torch.float32, device=dev)
    soft = mode == 1
    weight = torch.zeros((B, N) if soft else (0,), dtype=torch.float32, device=dev)
    soft_xyz = torch.zeros((B, N, 3) if soft else (0,), dtype=torch.float32, device=dev)
    lib = _lib_for(rows)
    with torch.cuda.device(dev):
        ws = _match_workspace(lib, dev)
        _lib.check(lib.gadm_match_fwd_sel(_ptr(rows), _ptr(rinv), _ptr(pad_sim) if pad_mode else None, _ptr(cols),
                                          _ptr(aux), _ptr(n_rows), _ptr(row_map), N, _ptr(obj_id), B, N, M, kp, n_obj,
                                          float(gamma), pad_mode, mode, _ptr(idx), _ptr(max_sim),
                                          _ptr(weight) if soft else None, _ptr(soft_xyz) if soft else None,
                                          _ptr(ws), ws.numel(), _stream()), "gadm_match_fwd_sel")
    return idx, max_sim, weight, soft_xyz


@match_fwd_sel.register_fake
def _(rows, rinv, pad_sim, cols, aux, n_rows, row_map, obj_id, gamma, pad_mode, mode):
    B, N, _ = rows.shape
    soft = mode == 1
    return (rows.new_empty((B, N), dtype=torch.int64), rows.new_empty((B, N), dtype=torch.float32),
            rows.new_empty((B, N) if soft else (0,), dtype=torch.float32),
            rows.new_empty((B, N, 3) if soft else (0,), dtype=torch.float32))


@match_fwd.register_fake
def _(rows, rinv, pad_sim, cols, aux, mask, obj_id, gamma, pad_mode, mode):
    B, N, _ = rows.shape
    soft = mode == 1
    return (rows.new_empty((B, N), dtype=torch.int64), rows.new_empty((B, N), dtype=torch.float32),
            rows.new_empty((B, N) if soft else (0,), dtype=torch.float32),
            rows.new_empty((B, N, 3) if soft else (0,), dtype=torch.float32))


@torch.library.custom_op("gadm::circle_loss_fwd", mutates_args=(), device_types="cuda")
def circle_loss_fwd(rows: torch.Tensor, rinv: torch.Tensor, pad_sim: torch.Tensor, cols: torch.Tensor,
                    aux: torch.Tensor, planes_frame: torch.Tensor, match_idx: torch.Tensor,
                    fg: torch.Tensor | None, obj_id: torch.Tensor | None, gamma: float,
                    margin: float, match_idx2: torch.Tensor | None = None) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Per-row CircleLoss of the similarity with the padded model (gadm_circle_loss_fwd): (loss, lse_p, lse_n).
    planes_frame [4, B, M]: x / y / z of the vertices (invisible ones at 1e18) and their squared positive radius."""
    _need(rows, torch.bfloat16, "rows"); _need(cols, torch.bfloat16, "cols")
    _need(rinv, torch.float32, "rinv"); _need(pad_sim, torch.float32, "pad_sim"); _need(aux, torch.float32, "aux")
    _need(planes_frame, torch.float32, "planes_frame"); _need(match_idx, torch.int64, "match_idx")
    B, N, kp = rows.shape
    n_obj, M, kp2 = cols.shape
    if kp != kp2:
        raise ValueError(f"operand K mismatch: rows {kp} vs cols {kp2}")
    if tuple(planes_frame.shape) != (4, B, M):
        raise ValueError(f"planes_frame must be [4, {B}, {M}]")
    if tuple(match_idx.shape) != (B, N) or tuple(pad_sim.shape) != (B, N):
        raise ValueError("match_idx and pad_sim must be [B, N]")
    if fg is not None:
        _need(fg, torch.uint8, "fg")
        if tuple(fg.shape) != (B, N):
            raise ValueError("fg must be [B, N]")
    if obj_id is not None:
        _need(obj_id, torch.int32, "obj_id")
    if match_idx2 is not None:                     # exact-column positives (matching_loss_sys, geoMatch.py:86-100)
        _need(match_idx2, torch.int64, "match_idx2")
        if tuple(match_idx2.shape) != (B, N):
            raise ValueError("match_idx2 must be [B, N]")
    dev = rows.device
    loss = torch.empty((B, N), dtype=torch.float32, device=dev)
    lse_p = torch.empty((B, N), dtype=torch.float32, device=dev)
    lse_n = torch.empty((B, N), dtype=torch.float32, device=dev)
    lib = _lib_for(rows)
    with torch.cuda.device(dev):
        _lib.check(lib.gadm_circle_loss_fwd(_ptr(rows), _ptr(rinv), _ptr(pad_sim), _ptr(cols), _ptr(aux),
                                            _ptr(planes_frame), _ptr(match_idx), _ptr(match_idx2), _ptr(fg), _ptr(obj_id),
                                            B, N, M, kp, n_obj, float(gamma), float(margin), _ptr(loss), _ptr(lse_p),
                                            _ptr(lse_n), _stream()), "gadm_circle_loss_fwd")
    return loss, lse_p, lse_n


@circle_loss_fwd.register_fake
def _(rows, rinv, pad_sim, cols, aux, planes_frame, match_idx, fg, obj_id, gamma, margin, match_idx2=None):
    B, N, _ = rows.shape
    return (rows.new_empty((B, N), dtype=torch.float32), rows.new_empty((B, N), dtype=torch.float32),
            rows.new_empty((B, N), dtype=torch.float32))


@torch.library.custom_op("gadm::circle_loss_bwd", mutates_args=(), device_types="cuda")
def circle_loss_bwd(rows: torch.Tensor, rinv: torch.Tensor, pad_sim: torch.Tensor, cols: torch.Tensor,
                    aux: torch.Tensor, planes_frame: torch.Tensor, match_idx: torch.Tensor,
                    obj_id: torch.Tensor | None, gamma: float, margin: float, lse_p: torch.Tensor,
                    lse_n: torch.Tensor, w: torch.Tensor, match_idx2: torch.Tensor | None = None) -> torch.Tensor:
    """dL/dsim [B, N, M + 8] (column M = pad column, the rest of the padding 0) for per-row upstream gradients w
    (gadm_circle_loss_bwd)."""
    _need(rows, torch.bfloat16, "rows"); _need(cols, torch.bfloat16, "cols")
    for t, n in ((rinv, "rinv"), (pad_sim, "pad_sim"), (aux, "aux"), (planes_frame, "planes_frame"), (lse_p, "lse_p"),
                 (lse_n, "lse_n"), (w, "w")):
        _need(t, torch.float32, n)
    _need(match_idx, torch.int64, "match_idx")
    B, N, kp = rows.shape
    n_obj, M, _ = cols.shape
    Mp = M + 8
    G = torch.empty((B, N, Mp), dtype=torch.float32, device=rows.device)
    lib = _lib_for(rows)
    with torch.cuda.device(rows.device):
        _lib.check(lib.gadm_circle_loss_bwd(_ptr(rows), _ptr(rinv), _ptr(pad_sim), _ptr(cols), _ptr(aux),
                                            _ptr(planes_frame), _ptr(match_idx), _ptr(match_idx2), _ptr(obj_id), B, N, M,
                                            kp, n_obj, float(gamma), float(margin), _ptr(lse_p), _ptr(lse_n),
                                            _ptr(w), _ptr(G), Mp, _stream()), "gadm_circle_loss_bwd")
    return G


@circle_loss_bwd.register_fake
def _(rows, rinv, pad_sim, cols, aux, planes_frame, match_idx, obj_id, gamma, margin, lse_p, lse_n, w,
      match_idx2=None):
    B, N, _ = rows.shape
    return rows.new_empty((B, N, cols.shape[1] + 8), dtype=torch.float32)


@torch.library.custom_op("gadm::circle_loss_bwd_split", mutates_args=(), device_types="cuda")
def circle_loss_bwd_split(rows: torch.Tensor, rinv: torch.Tensor, pad_sim: torch.Tensor, cols: torch.Tensor,
                          aux: torch.Tensor, planes_frame: torch.Tensor, match_idx: torch.Tensor,
                          obj_id: torch.Tensor | None, gamma: float, margin: float, lse_p: torch.Tensor,
                          lse_n: torch.Tensor, w: torch.Tensor,
                          match_idx2: torch.Tensor | None = None) -> tuple[torch.Tensor, torch.Tensor]:
    """(G2 [B, N, 2 (M + 8)] bf16, g_pad [B, N] fp32): dL/dsim with both norms folded in, as hi + lo bf16 parts
    interleaved in groups of 8 columns, and the pad column's gradient (gadm_circle_loss_bwd_split)."""
    _need(rows, torch.bfloat16, "rows"); _need(cols, torch.bfloat16, "cols")
    for t, n in ((rinv, "rinv"), (pad_sim, "pad_sim"), (aux, "aux"), (planes_frame, "planes_frame"), (lse_p, "lse_p"),
                 (lse_n, "lse_n"), (w, "w")):
        _need(t, torch.float32, n)
    _need(match_idx, torch.int64, "match_idx")
    B, N, kp = rows.shape
    n_obj, M, _ = cols.shape
    Mp = M + 8
    G2 = torch.empty((B, N, 2 * Mp), dtype=torch.bfloat16, device=rows.device)
    g_pad = torch.empty((B, N), dtype=torch.float32, device=rows.device)
    lib = _lib_for(rows)
    with torch.cuda.device(rows.device):
        _lib.check(lib.gadm_circle_loss_bwd_split(_ptr(rows), _ptr(rinv), _ptr(pad_sim), _ptr(cols), _ptr(aux),
                                                  _ptr(planes_frame), _ptr(match_idx), _ptr(match_idx2), _ptr(obj_id),
                                                  B, N, M, kp, n_obj, float(gamma), float(margin), _ptr(lse_p),
                                                  _ptr(lse_n), _ptr(w), _ptr(G2), Mp, _ptr(g_pad), _stream()),
                   "gadm_circle_loss_bwd_split")
    return G2, g_pad


@circle_loss_bwd_split.register_fake
def _(rows, rinv, pad_sim, cols, aux, planes_frame, match_idx, obj_id, gamma, margin, lse_p, lse_n, w,
      match_idx2=None):
    B, N, _ = rows.shape
    return (rows.new_empty((B, N, 2 * (cols.shape[1] + 8)), dtype=torch.bfloat16),
            rows.new_empty((B, N), dtype=torch.float32))


@torch.library.custom_op("gadm::circle_loss_bwd_fused", mutates_args=(), device_types="cuda")
def circle_loss_bwd_fused(rows: torch.Tensor, rinv: torch.Tensor, pad_sim: torch.Tensor, cols: torch.Tensor,
                          aux: torch.Tensor, planes_frame: torch.Tensor, match_idx: torch.Tensor,
                          obj_id: torch.Tensor | None, gamma: float, margin: float, lse_p: torch.Tensor,
                          lse_n: torch.Tensor, w: torch.Tensor, match_idx2: torch.Tensor | None = None,
                          model_side: bool = False) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(G2, g_pad, dF, dM) of gadm_circle_loss_bwd_fused (K' <= 128): dF [B, N, K'] fp32 = sum_j G''_ij cols_j, accumulated
    in tensor memory by a second MMA inside the kernel.  model_side=False: G2 as circle_loss_bwd_split, dM empty.
    model_side=True: dM [B, M + 8, K'] fp32 = sum_i G''_ij rows_i formed in the kernel as well (third MMA per model tile,
    fp32 reductions into global memory); G2 is not written (empty)."""
    _need(rows, torch.bfloat16, "rows"); _need(cols, torch.bfloat16, "cols")
    for t, n in ((rinv, "rinv"), (pad_sim, "pad_sim"), (aux, "aux"), (planes_frame, "planes_frame"), (lse_p, "lse_p"),
                 (lse_n, "lse_n"), (w, "w")):
        _need(t, torch.float32, n)
    _need(match_idx, torch.int64, "match_idx")
    B, N, kp = rows.shape
    n_obj, M, _ = cols.shape
    Mp = M + 8
    dev = rows.device
    G2 = torch.empty((0,) if model_side else (B, N, 2 * Mp), dtype=torch.bfloat16, device=dev)
    dM = torch.zeros((B, Mp, kp), dtype=torch.float32, device=dev) if model_side else \
        torch.empty((0,), dtype=torch.float32, device=dev)
    g_pad = torch.empty((B, N), dtype=torch.float32, device=dev)
    dF = torch.empty((B, N, kp), dtype=torch.float32, device=dev)
    lib = _lib_for(rows)
    with torch.cuda.device(dev):
        _lib.check(lib.gadm_circle_loss_bwd_fused(_ptr(rows), _ptr(rinv), _ptr(pad_sim), _ptr(cols), _ptr(aux),
                                                  _ptr(planes_frame), _ptr(match_idx), _ptr(match_idx2), _ptr(obj_id),
                                                  B, N, M, kp, n_obj, float(gamma), float(margin), _ptr(lse_p),
                                                  _ptr(lse_n), _ptr(w), None if model_side else _ptr(G2), Mp,
                                                  _ptr(g_pad), _ptr(dF), _ptr(dM) if model_side else None, _stream()),
                   "gadm_circle_loss_bwd_fused")
    return G2, g_pad, dF, dM


@circle_loss_bwd_fused.register_fake
def _(rows, rinv, pad_sim, cols, aux, planes_frame, match_idx, obj_id, gamma, margin, lse_p, lse_n, w,
      match_idx2=None, model_side=False):
    B, N, kp = rows.shape
    Mp = cols.shape[1] + 8
    return (rows.new_empty((0,) if model_side else (B, N, 2 * Mp), dtype=torch.bfloat16),
            rows.new_empty((B, N), dtype=torch.float32), rows.new_empty((B, N, kp), dtype=torch.float32),
            rows.new_empty((B, Mp, kp) if model_side else (0,), dtype=torch.float32))


@torch.library.custom_op("gadm::kabsch_moments", mutates_args=(), device_types="cuda")
def kabsch_moments(idx: torch.Tensor, mask: torch.Tensor | None, cloud: torch.Tensor, aux: torch.Tensor,
                   obj_id: torch.Tensor | None, M: int, n_obj: int) -> torch.Tensor:
    _need(idx, torch.int64, "idx"); _need(cloud, torch.float32, "cloud"); _need(aux, torch.float32, "aux")
    B, N = idx.shape
    if tuple(cloud.shape) != (B, N, 3):
        raise ValueError("cloud must be [B, N, 3]")
    out = torch.empty((B, 16), dtype=torch.float64, device=idx.device)
    lib = _lib_for(idx)
    with torch.cuda.device(idx.device):
        _lib.check(lib.gadm_kabsch_moments(_ptr(idx), _ptr(mask), _ptr(cloud), _ptr(aux), _ptr(obj_id), B, N, M,
                                           n_obj, _ptr(out), _stream()), "gadm_kabsch_moments")
    return out


@kabsch_moments.register_fake
def _(idx, mask, cloud, aux, obj_id, M, n_obj):
    return idx.new_empty((idx.shape[0], 16), dtype=torch.float64)


@torch.library.custom_op("gadm::kabsch_moments_w", mutates_args=(), device_types="cuda")
def kabsch_moments_w(idx: torch.Tensor, mask: torch.Tensor | None, weight: torch.Tensor, cloud: torch.Tensor,
                     aux: torch.Tensor, obj_id: torch.Tensor | None, M: int, n_obj: int) -> torch.Tensor:
    """Weighted moments {sum w, sum w A, sum w B, sum w A B^T} (weighted Procrustes with the matcher's weights)."""
    _need(idx, torch.int64, "idx"); _need(cloud, torch.float32, "cloud"); _need(aux, torch.float32, "aux")
    _need(weight, torch.float32, "weight")
    B, N = idx.shape
    if tuple(cloud.shape) != (B, N, 3) or tuple(weight.shape) != (B, N):
        raise ValueError("cloud must be [B, N, 3] and weight [B, N]")
    out = torch.empty((B, 16), dtype=torch.float64, device=idx.device)
    lib = _lib_for(idx)
    with torch.cuda.device(idx.device):
        _lib.check(lib.gadm_kabsch_moments_w(_ptr(idx), _ptr(mask), _ptr(weight), _ptr(cloud), _ptr(aux), _ptr(obj_id),
                                             B, N, M, n_obj, _ptr(out), _stream()), "gadm_kabsch_moments_w")
    return out


@kabsch_moments_w.register_fake
def _(idx, mask, weight, cloud, aux, obj_id, M, n_obj):
    return idx.new_empty((idx.shape[0], 16), dtype=torch.float64)


@torch.library.custom_op("gadm::kabsch_poses", mutates_args=(), device_types="cuda")
def kabsch_poses(moments: torch.Tensor, count_moments: torch.Tensor | None, det: torch.Tensor | None,
                 min_pts: int) -> torch.Tensor:
    """moments [B, 16] fp64 -> poses [B, 3, 4] fp32 on the device (best_fit_transform + the evaluator's sentinel)."""
    _need(moments, torch.float64, "moments")
    if count_moments is not None:
        _need(count_moments, torch.float64, "count_moments")
    if det is not None:
        _need(det, torch.uint8, "det")
    B = moments.shape[0]
    poses = torch.empty((B, 3, 4), dtype=torch.float32, device=moments.device)
    lib = _lib_for(moments)
    with torch.cuda.device(moments.device):
        _lib.check(lib.gadm_kabsch_poses(_ptr(moments), _ptr(count_moments), _ptr(det), B, int(min_pts), _ptr(poses),
                                         _stream()), "gadm_kabsch_poses")
    return poses


@kabsch_poses.register_fake
def _(moments, count_moments, det, min_pts):
    return moments.new_empty((moments.shape[0], 3, 4), dtype=torch.float32)


# --------------------------------------------------------------------------------------------- kNN 3-D
def make_jobs(job_list):
    """[(support_off, query_off, out_off, s_bstride, q_bstride, o_bstride, n_support, n_query, k, batch)]."""
    arr = (KnnJob * len(job_list))()
    for a, j in zip(arr, job_list):
        (a.support_off, a.query_off, a.out_off, a.support_bstride, a.query_bstride, a.out_bstride,
         a.n_support, a.n_query, a.k, a.batch) = j
    return arr


def knn3d_jobs(support, query, jobs, out_elems, algo="auto", return_dist=False, workspace=None, out=None):
    """Run a job table over flat point buffers.  support/query: [P, 3] fp32 CUDA; returns int32 [out_elems]
    (`out`: write the indices there instead of allocating)."""
    _need(support, torch.float32, "support"); _need(query, torch.float32, "query")
    lib = _lib_for(support)
    a = KNN_ALGOS[algo] if isinstance(algo, str) else int(algo)
    need = lib.gadm_knn3d_workspace_bytes(jobs, len(jobs), a)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty((max(need, 16),), dtype=torch.uint8, device=support.device)
    if out is not None:
        _need(out, torch.int32, "out")
        if out.numel() != out_elems:
            raise ValueError(f"out must hold {out_elems} int32")
    idx = out if out is not None else torch.empty((out_elems,), dtype=torch.int32, device=support.device)
    d2 = torch.empty((out_elems,), dtype=torch.float32, device=support.device) if return_dist else None
    with torch.cuda.device(support.device):
        _lib.check(lib.gadm_knn3d(_ptr(support), _ptr(query), jobs, len(jobs), a, _ptr(idx), _ptr(d2),
                                  _ptr(workspace), workspace.numel(), _stream()), "gadm_knn3d")
    return (idx, d2) if return_dist else idx


@torch.library.custom_op("gadm::knn3d", mutates_args=(), device_types="cuda")
def knn3d(support: torch.Tensor, query: torch.Tensor, k: int, algo: int) -> torch.Tensor:
    """support [B, N1, 3], query [B, N2, 3] -> int32 [B, N2, k]."""
    B, n1, _ = support.shape
    n2 = query.shape[1]
    jobs = make_jobs([(0, 0, 0, n1, n2, n2 * k, n1, n2, k, B)])
    return knn3d_jobs(support.view(-1, 3), query.view(-1, 3), jobs, B * n2 * k, algo).view(B, n2, k)


@knn3d.register_fake
def _(support, query, k, algo):
    return support.new_empty((support.shape[0], query.shape[1], k), dtype=torch.int32)


# --------------------------------------------------------------------------------------------- DGCNN
_KNN_FEAT_TC = True     # tests switch the tensor-core path off to compare it with the fp32 SIMT kernel


@torch.library.custom_op("gadm::knn_feat", mutates_args=(), device_types="cuda")
def knn_feat(x: torch.Tensor, k: int, kdim: int) -> torch.Tensor:
    """models/dgcnn.py:21-27.  Shapes the tensor-core kernel takes (kdim == C, C % 64 == 0, k <= 20, N % 4 == 0,
    N >= 256) run there (bf16x3 split dot products, fp32 accumulate); everything else on the fp32 SIMT kernel."""
    _need(x, torch.float32, "x")
    B, C, N = x.shape
    idx = torch.empty((B, N, k), dtype=torch.int64, device=x.device)
    lib = _lib_for(x)
    ws_bytes = lib.gadm_knn_feat_tc_workspace_bytes(B, C, N, kdim, k) if _KNN_FEAT_TC else 0
    with torch.cuda.device(x.device):
        if ws_bytes:
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
            _lib.check(lib.gadm_knn_feat_tc(_ptr(x), B, C, N, k, _ptr(idx), _ptr(ws), ws_bytes, _stream()),
                       "gadm_knn_feat_tc")
        else:
            _lib.check(lib.gadm_knn_feat(_ptr(x), B, C, N, kdim, k, _ptr(idx), _stream()), "gadm_knn_feat")
    return idx


@knn_feat.register_fake
def _(x, k, kdim):
    return x.new_empty((x.shape[0], x.shape[2], k), dtype=torch.int64)


@torch.library.custom_op("gadm::graph_feature", mutates_args=(), device_types="cuda")
def graph_feature(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _need(x, torch.float32, "x"); _need(idx, torch.int64, "idx")
    B, C, N = x.shape
    k = idx.shape[2]
    out = torch.empty((B, 2 * C, N, k), dtype=torch.float32, device=x.device)
    lib = _lib_for(x)
    ws_bytes = lib.gadm_graph_feature_workspace_bytes(B, C, N)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device) if ws_bytes else None
    with torch.cuda.device(x.device):
        _lib.check(lib.gadm_graph_feature(_ptr(x), _ptr(idx), B, C, N, k, _ptr(out), _ptr(ws), ws_bytes, _stream()),
                   "gadm_graph_feature")
    return out


@graph_feature.register_fake
def _(x, idx):
    B, C, N = x.shape
    return x.new_empty((B, 2 * C, N, idx.shape[2]))


@torch.library.custom_op("gadm::graph_feature_bwd", mutates_args=(), device_types="cuda")
def graph_feature_bwd(grad_out: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """grad_out [B, 2C, N, k] -> grad_x [B, C, N] (models/dgcnn.py:49-54 differentiated)."""
    _need(grad_out, torch.float32, "grad_out"); _need(idx, torch.int64, "idx")
    B, C2, N, k = grad_out.shape
    gx = torch.empty((B, C2 // 2, N), dtype=torch.float32, device=grad_out.device)
    lib = _lib_for(grad_out)
    with torch.cuda.device(grad_out.device):
        _lib.check(lib.gadm_graph_feature_bwd(_ptr(grad_out), _ptr(idx), B, C2 // 2, N, k, _ptr(gx), _stream()),
                   "gadm_graph_feature_bwd")
    return gx


@graph_feature_bwd.register_fake
def _(grad_out, idx):
    B, C2, N, k = grad_out.shape
    return grad_out.new_empty((B, C2 // 2, N))


def _gf_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[1])


def _gf_backward(ctx, grad):
    (idx,) = ctx.saved_tensors
    return graph_feature_bwd(grad.contiguous(), idx), None


graph_feature.register_autograd(_gf_backward, setup_context=_gf_setup)


# --------------------------------------------------------------------------------------------- grouping
@torch.library.custom_op("gadm::group_fwd", mutates_args=(), device_types="cuda")
def group_fwd(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _need(features, torch.float32, "features"); _need(idx, torch.int32, "idx")
    b, c, n = features.shape
    _, m, s = idx.shape
    out = torch.empty((b, c, m, s), dtype=torch.float32, device=features.device)
    lib = _lib_for(features)
    with torch.cuda.device(features.device):
        _lib.check(lib.gadm_group_fwd(_ptr(features), _ptr(idx), b, c, n, m, s, _ptr(out), _stream()),
                   "gadm_group_fwd")
    return out


@group_fwd.register_fake
def _(features, idx):
    return features.new_empty((features.shape[0], features.shape[1], idx.shape[1], idx.shape[2]))


@torch.library.custom_op("gadm::group_bwd", mutates_args=(), device_types="cuda")
def group_bwd(grad_out: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    _need(grad_out, torch.float32, "grad_out"); _need(idx, torch.int32, "idx")
    b, c, m, s = grad_out.shape
    gf = torch.empty((b, c, n), dtype=torch.float32, device=grad_out.device)
    lib = _lib_for(grad_out)
    with torch.cuda.device(grad_out.device):
        _lib.check(lib.gadm_group_bwd(_ptr(grad_out), _ptr(idx), b, c, n, m, s, _ptr(gf), _stream()),
                   "gadm_group_bwd")
    return gf


@group_bwd.register_fake
def _(grad_out, idx, n):
    return grad_out.new_empty((grad_out.shape[0], grad_out.shape[1], n))


def _group_setup(ctx, inputs, output):
    features, idx = inputs
    ctx.save_for_backward(idx)
    ctx.n = features.shape[2]


def _group_backward(ctx, grad):
    (idx,) = ctx.saved_tensors
    return group_bwd(grad.contiguous(), idx, ctx.n), None


group_fwd.register_autograd(_group_backward, setup_context=_group_setup)


@torch.library.custom_op("gadm::gather_neighbour", mutates_args=(), device_types="cuda")
def gather_neighbour(pc: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _need(pc, torch.float32, "pc"); _need(idx, torch.int64, "idx")
    B, N, C = pc.shape
    _, M, K = idx.shape
    out = torch.empty((B, M, K, C), dtype=torch.float32, device=pc.device)
    lib = _lib_for(pc)
    with torch.cuda.device(pc.device):
        _lib.check(lib.gadm_gather_neighbour(_ptr(pc), _ptr(idx), B, N, C, M, K, _ptr(out), _stream()),
                   "gadm_gather_neighbour")
    return out


@gather_neighbour.register_fake
def _(pc, idx):
    return pc.new_empty((pc.shape[0], idx.shape[1], idx.shape[2], pc.shape[2]))


@torch.library.custom_op("gadm::gather_neighbour_bwd", mutates_args=(), device_types="cuda")
def gather_neighbour_bwd(grad_out: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    """grad_out [B, M, K, C] -> grad_pc [B, N, C] (RandLANet.py:729-738 differentiated)."""
    _need(grad_out, torch.float32, "grad_out"); _need(idx, torch.int64, "idx")
    B, M, K, C = grad_out.shape
    g = torch.empty((B, n, C), dtype=torch.float32, device=grad_out.device)
    lib = _lib_for(grad_out)
    with torch.cuda.device(grad_out.device):
        _lib.check(lib.gadm_gather_neighbour_bwd(_ptr(grad_out), _ptr(idx), B, n, C, M, K, _ptr(g), _stream()),
                   "gadm_gather_neighbour_bwd")
    return g


@gather_neighbour_bwd.register_fake
def _(grad_out, idx, n):
    return grad_out.new_empty((grad_out.shape[0], n, grad_out.shape[3]))


def _gn_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[1])
    ctx.n = inputs[0].shape[1]


def _gn_backward(ctx, grad):
    (idx,) = ctx.saved_tensors
    return gather_neighbour_bwd(grad.contiguous(), idx, ctx.n), None


gather_neighbour.register_autograd(_gn_backward, setup_context=_gn_setup)


# --------------------------------------------------------------------------------------------- RandLA consumers
@torch.library.custom_op("gadm::gather_max", mutates_args=(), device_types="cuda")
def gather_max(feature: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """feature [B, C, N] fp32, idx [B, M, K] int64 -> [B, C, M] = max_k feature[b, c, idx[b, m, k]]."""
    _need(feature, torch.float32, "feature"); _need(idx, torch.int64, "idx")
    B, C, N = feature.shape
    _, M, K = idx.shape
    out = torch.empty((B, C, M), dtype=torch.float32, device=feature.device)
    lib = _lib_for(feature)
    with torch.cuda.device(feature.device):
        _lib.check(lib.gadm_gather_max(_ptr(feature), _ptr(idx), B, C, N, M, K, _ptr(out), _stream()),
                   "gadm_gather_max")
    return out


@gather_max.register_fake
def _(feature, idx):
    return feature.new_empty((feature.shape[0], feature.shape[1], idx.shape[1]))


@torch.library.custom_op("gadm::gather_max_bwd", mutates_args=(), device_types="cuda")
def gather_max_bwd(feature: torch.Tensor, idx: torch.Tensor, grad_out: torch.Tensor) -> torch.Tensor:
    """grad_out [B, C, M] -> grad_feature [B, C, N]: to the first neighbour attaining each maximum
    (RandLANet.py:90-120 differentiated)."""
    _need(feature, torch.float32, "feature"); _need(idx, torch.int64, "idx"); _need(grad_out, torch.float32, "grad_out")
    B, C, N = feature.shape
    _, M, K = idx.shape
    g = torch.empty((B, C, N), dtype=torch.float32, device=feature.device)
    lib = _lib_for(feature)
    with torch.cuda.device(feature.device):
        _lib.check(lib.gadm_gather_max_bwd(_ptr(feature), _ptr(idx), _ptr(grad_out), B, C, N, M, K, _ptr(g), _stream()),
                   "gadm_gather_max_bwd")
    return g


@gather_max_bwd.register_fake
def _(feature, idx, grad_out):
    return feature.new_empty(feature.shape)


def _gm_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])


def _gm_backward(ctx, grad):
    feature, idx = ctx.saved_tensors
    return gather_max_bwd(feature, idx, grad.contiguous()), None


gather_max.register_autograd(_gm_backward, setup_context=_gm_setup)


def _rpe_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])


def _rpe_backward(ctx, grad):
    """[dis, rel, tile, nbr] with rel = tile - nbr, dis = |rel| (RandLANet.py:720-727): the gradient reaches xyz through
    the tiled centre (summed over K) and through the gathered neighbour (scatter-add, gather_neighbour_bwd)."""
    xyz, idx = ctx.saved_tensors
    enc = relative_pos_encoding(xyz, idx)
    dis, rel = enc[..., 0:1], enc[..., 1:4]
    g_rel = grad[..., 1:4] + grad[..., 0:1] * rel / dis.clamp_min(1e-30)
    g_tile = grad[..., 4:7] + g_rel
    g_nbr = grad[..., 7:10] - g_rel
    return g_tile.sum(dim=2) + gather_neighbour_bwd(g_nbr.contiguous(), idx, xyz.shape[1]), None


@torch.library.custom_op("gadm::relative_pos_encoding", mutates_args=(), device_types="cuda")
def relative_pos_encoding(xyz: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """xyz [B, N, 3] fp32, idx [B, N, K] int64 -> [B, N, K, 10] = [|p-q|, p-q, p, q]."""
    _need(xyz, torch.float32, "xyz"); _need(idx, torch.int64, "idx")
    B, N, _ = xyz.shape
    K = idx.shape[2]
    out = torch.empty((B, N, K, 10), dtype=torch.float32, device=xyz.device)
    lib = _lib_for(xyz)
    with torch.cuda.device(xyz.device):
        _lib.check(lib.gadm_relative_pos_encoding(_ptr(xyz), _ptr(idx), B, N, K, _ptr(out), _stream()),
                   "gadm_relative_pos_encoding")
    return out


@relative_pos_encoding.register_fake
def _(xyz, idx):
    return xyz.new_empty((xyz.shape[0], xyz.shape[1], idx.shape[2], 10))


relative_pos_encoding.register_autograd(_rpe_backward, setup_context=_rpe_setup)


@torch.library.custom_op("gadm::seg_mask", mutates_args=(), device_types="cuda")
def seg_mask(seg: torch.Tensor) -> torch.Tensor:
    """seg [B, 2, N] fp32 -> uint8 [B, N] = (argmax over dim 1 == 1)   (evaluator.py:78,82)."""
    _need(seg, torch.float32, "seg")
    B, two, N = seg.shape
    if two != 2:
        raise ValueError("seg must be [B, 2, N]")
    out = torch.empty((B, N), dtype=torch.uint8, device=seg.device)
    lib = _lib_for(seg)
    with torch.cuda.device(seg.device):
        _lib.check(lib.gadm_seg_mask(_ptr(seg), B, N, _ptr(out), _stream()), "gadm_seg_mask")
    return out


@seg_mask.register_fake
def _(seg):
    return seg.new_empty((seg.shape[0], seg.shape[2]), dtype=torch.uint8)
