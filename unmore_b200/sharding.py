"""Image sharding across the GPUs of one node and the single end-of-run collective.

The reference shards by hand: several processes with disjoint ``--start_idx/--end_idx``
ranges, one GPU each, no communication (datasets.py:432-435, object_reasoning.py:99-100).
Images are independent on this path (the loop at object_reasoning.py:617 carries no state),
so the same partition is kept — one process per GPU, zero traffic while reasoning — and the
per-image detections are exchanged once at the end with ONE ``all_gather_into_tensor`` of a
fixed-capacity row buffer (NCCL over NVLink on GPUs; gloo in the CPU tests).

Row buffer (``ops.detection_rows`` / ``unmore_pack_detections``): ``[max_rows + 1, 6]`` fp64, row 0 is the header
``(count, overflow, 0, 0, 0, 0)``, row ``1 + r`` is ``(image_id, x, y, w, h, score)``.  The scoring stage
appends to it on the device, so there is no staging copy, no per-chunk boolean indexing and no host
synchronisation between the last kernel and the collective.  Image ids travel as doubles (exact to 2^53)."""
from __future__ import annotations

import hashlib
from typing import List, Tuple

import torch
import torch.distributed as dist

ROW_WIDTH = 6


def shard_range(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, end) of rank ``rank`` — the reference's --start_idx/--end_idx."""
    base, rem = divmod(n_images, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_indices(n_images: int, rank: int, world: int, interleave: bool = False) -> List[int]:
    if interleave:  # i mod G balances proposal-count variance across ranks
        return list(range(rank, n_images, world))
    s, e = shard_range(n_images, rank, world)
    return list(range(s, e))


def pack_rows_host(image_index: torch.Tensor, boxes: torch.Tensor, counts: torch.Tensor, scores=None,
                   max_rows: int = 0) -> torch.Tensor:
    """Host/torch restatement of ``unmore_pack_detections`` (CPU tests and tooling): ragged per-image results
    -> the row buffer described in the module docstring.  boxes [B, cap, 4] are written as they are (the
    caller decides between xyxy and COCO xywh); scores [B, cap] or None (-> 1.0)."""
    B, cap = boxes.shape[0], boxes.shape[1]
    valid = torch.arange(cap, device=boxes.device)[None, :] < counts[:, None].to(torch.long)
    idx = image_index.to(boxes.device, torch.float64)[:, None].expand(B, cap)
    sc = scores if scores is not None else torch.ones((B, cap), dtype=torch.float64, device=boxes.device)
    rows = torch.cat([idx[..., None], boxes.to(torch.float64), sc.to(torch.float64)[..., None]], dim=2)[valid]
    n = rows.shape[0]
    max_rows = max(max_rows, n)
    buf = torch.zeros((max_rows + 1, ROW_WIDTH), dtype=torch.float64, device=boxes.device)
    buf[0, 0] = n
    buf[1:1 + n] = rows
    return buf


def gather_rows(buf: torch.Tensor) -> torch.Tensor:
    """ONE collective: every rank receives every rank's row buffer -> [world, max_rows + 1, 6].
    All ranks must use the same capacity.  No host synchronisation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return buf[None]
    world = dist.get_world_size()
    out = torch.empty((world * buf.shape[0], buf.shape[1]), dtype=buf.dtype, device=buf.device)   # dim-0 concatenation
    dist.all_gather_into_tensor(out, buf.contiguous())
    return out.view(world, buf.shape[0], buf.shape[1])


def merge_rows(gathered: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[world, max_rows + 1, 6] -> (rows [world * max_rows, 6] with the valid rows FIRST, stably sorted by
    image id (rank order, then arrival order, inside an image — an image lives on one rank, so this equals
    a single-process run over all images), total [1] int64).  Device-side, no host synchronisation: the
    caller slices ``rows[:int(total)]`` when it reads the result."""
    world, m1, w = gathered.shape
    counts = gathered[:, 0, 0].to(torch.int64)
    rows = gathered[:, 1:, :].reshape(world * (m1 - 1), w)
    pos = torch.arange(m1 - 1, device=gathered.device)[None, :].expand(world, m1 - 1)
    valid = (pos < counts[:, None]).reshape(-1)
    key = torch.where(valid, rows[:, 0], torch.full_like(rows[:, 0], float("inf")))
    order = torch.sort(key, stable=True).indices
    return rows[order], counts.sum().reshape(1)


def overflowed(gathered: torch.Tensor) -> torch.Tensor:
    """True (device bool) if any rank ran out of row capacity."""
    return (gathered[:, 0, 1] != 0).any()


def gather_detections(buf: torch.Tensor) -> torch.Tensor:
    """Convenience form that reads the result: all ranks' detections, image-sorted ([K, 6] fp64).
    Synchronises (one ``.item()``) — use gather_rows / merge_rows inside timed regions."""
    g = gather_rows(buf)
    if bool(overflowed(g)):
        raise RuntimeError("detection row buffer overflowed on some rank: raise max_rows")
    rows, total = merge_rows(g)
    return rows[: int(total.item())]


def rows_digest(rows: torch.Tensor) -> str:
    """sha256 over the bytes of the (already image-sorted) detection rows: equal digests <=> bit-identical
    detections, which is how an N-GPU run is compared with the 1-GPU run of the same images."""
    return hashlib.sha256(rows.detach().cpu().contiguous().numpy().tobytes()).hexdigest()
