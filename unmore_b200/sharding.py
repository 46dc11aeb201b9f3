"""Image sharding across the GPUs of one node and the single end-of-run collective.

The reference shards by hand: several processes with disjoint ``--start_idx/--end_idx``
ranges, one GPU each, no communication (datasets.py:432-435, object_reasoning.py:99-100).
Images are independent on this path (the loop at object_reasoning.py:617 carries no state),
so the same partition is kept — one process per GPU, zero traffic while reasoning — and the
per-image detections are exchanged once at the end with an all-gather (NCCL over NVLink on
GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


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


def pack_detections(image_index: torch.Tensor, boxes: torch.Tensor, counts: torch.Tensor, scores=None) -> torch.Tensor:
    """Ragged per-image results -> flat rows (image_idx, x1, y1, x2, y2, score), image-major, order kept.

    image_index [B] (global image ids), boxes [B, cap, 4], counts [B]; scores [B, cap] or None (-> 1.0)."""
    B, cap = boxes.shape[0], boxes.shape[1]
    valid = torch.arange(cap, device=boxes.device)[None, :] < counts[:, None].to(torch.long)
    idx = image_index.to(boxes.device, torch.float32)[:, None].expand(B, cap)
    sc = scores if scores is not None else torch.ones((B, cap), dtype=torch.float32, device=boxes.device)
    rows = torch.cat([idx[..., None], boxes.to(torch.float32), sc.to(torch.float32)[..., None]], dim=2)
    return rows[valid]


def gather_detections(rows: torch.Tensor) -> torch.Tensor:
    """All ranks receive every rank's rows, concatenated in rank order then stably sorted by image
    index, so the result equals a single-process run over all images.  Two collectives: counts,
    then rows padded to the largest count."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rows
    world = dist.get_world_size()
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts + [1])
    padded = torch.zeros((n_max, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    padded[: rows.shape[0]] = rows
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    merged = torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)
    order = torch.sort(merged[:, 0], stable=True).indices
    return merged[order]
