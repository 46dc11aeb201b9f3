"""Tensor-level wrappers over the C ABI: pointer extraction, output allocation, stream.

PyTorch is only the allocator / stream provider here; all arithmetic runs in
``libunmore_b200.so``.  Inputs must be CUDA tensors; there is no CPU path."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from .synth import CH_CCOL, CH_CROW, CH_EXIST, CH_SDF

CROP = 128


@dataclass(frozen=True)
class Channels:
    sdf: int = CH_SDF
    center_row: int = CH_CROW
    center_col: int = CH_CCOL
    exist: int = CH_EXIST


DEFAULT_CHANNELS = Channels()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _check_fields(fields: torch.Tensor) -> Tuple[int, int, int, int]:
    if not fields.is_cuda:
        raise _lib.UnmoreError("fields must live on a CUDA device (no CPU path)")
    if fields.dtype != torch.float32 or fields.dim() != 4 or not fields.is_contiguous():
        raise _lib.UnmoreError("fields must be contiguous fp32 [n_img, C, H, W]")
    return tuple(fields.shape)  # type: ignore[return-value]


def _check_boxes(boxes: torch.Tensor, n_img: int) -> Tuple[int, int]:
    if boxes.dim() != 3 or boxes.shape[0] != n_img or boxes.shape[2] != 4 or not boxes.is_contiguous():
        raise _lib.UnmoreError("boxes must be contiguous [n_img, cap, 4]")
    if boxes.dtype not in (torch.float32, torch.float64):
        raise _lib.UnmoreError("boxes must be fp32 or fp64")
    return boxes.shape[1], int(boxes.dtype == torch.float64)


def _check_counts(counts: Optional[torch.Tensor], n_img: int):
    if counts is not None and (counts.dtype != torch.int32 or counts.numel() != n_img or not counts.is_cuda):
        raise _lib.UnmoreError("counts must be a CUDA int32 tensor of n_img entries")


def workspace(n_img: int, device) -> torch.Tensor:
    n = _lib.load().unmore_workspace_bytes(n_img)
    return torch.empty((n + 3) // 4, dtype=torch.int32, device=device)


def existence_scores(fields, boxes, counts=None, ch: Channels = DEFAULT_CHANNELS, ws=None, out=None):
    n_img, C, H, W = _check_fields(fields)
    cap, f64 = _check_boxes(boxes, n_img)
    _check_counts(counts, n_img)
    ws = workspace(n_img, fields.device) if ws is None else ws
    out = torch.zeros((n_img, cap), dtype=torch.float32, device=fields.device) if out is None else out
    _lib.call("unmore_existence_scores", fields.data_ptr(), n_img, C, H, W, ch.exist, boxes.data_ptr(), f64,
              _ptr(counts), cap, out.data_ptr(), ws.data_ptr(), _stream())
    return out


def center_reasoning(fields, boxes, counts=None, thr: float = 0.009, ch: Channels = DEFAULT_CHANNELS, ws=None,
                     want_splits: bool = True):
    n_img, C, H, W = _check_fields(fields)
    cap, f64 = _check_boxes(boxes, n_img)
    _check_counts(counts, n_img)
    dev = fields.device
    ws = workspace(n_img, dev) if ws is None else ws
    maxv = torch.zeros((n_img, cap), dtype=torch.float64, device=dev)
    argmax = torch.full((n_img, cap), -1, dtype=torch.int32, device=dev)
    splits = torch.zeros((n_img, cap, 4, 4), dtype=torch.float64, device=dev) if want_splits else None
    _lib.call("unmore_center_reasoning", fields.data_ptr(), n_img, C, H, W, ch.sdf, ch.center_row, ch.center_col,
              boxes.data_ptr(), f64, _ptr(counts), cap, float(thr), maxv.data_ptr(), argmax.data_ptr(),
              _ptr(splits), ws.data_ptr(), _stream())
    return maxv, argmax, splits


def boundary_refine(fields, boxes, counts=None, n_round: int = 50, apply_small_filter: bool = True,
                    early_exit: bool = True, proposal_area_thres: float = 50.0, max_sdf_thres: float = 0.5,
                    max_shrink_threshold: float = 16.0, delta_ratio: float = 0.5, ch: Channels = DEFAULT_CHANNELS,
                    ws=None, want_rounds: bool = True):
    n_img, C, H, W = _check_fields(fields)
    cap, f64 = _check_boxes(boxes, n_img)
    _check_counts(counts, n_img)
    dev = fields.device
    ws = workspace(n_img, dev) if ws is None else ws
    out = torch.zeros((n_img, cap, 4), dtype=torch.float32, device=dev)
    labels = torch.full((n_img, cap), -2.0, dtype=torch.float32, device=dev)
    rounds = torch.zeros((n_img, cap), dtype=torch.int32, device=dev) if want_rounds else None
    _lib.call("unmore_boundary_refine", fields.data_ptr(), n_img, C, H, W, ch.sdf, boxes.data_ptr(), f64,
              _ptr(counts), cap, int(n_round), int(apply_small_filter), int(early_exit), float(proposal_area_thres),
              float(max_sdf_thres), float(max_shrink_threshold), float(delta_ratio), out.data_ptr(),
              labels.data_ptr(), _ptr(rounds), ws.data_ptr(), _stream())
    return out, labels, rounds


def update_bbox_from_tiles(tiles: torch.Tensor):
    if not tiles.is_cuda or tiles.dtype != torch.float32 or tiles.dim() != 3 or tiles.shape[1:] != (CROP, CROP):
        raise _lib.UnmoreError("tiles must be CUDA fp32 [M, 128, 128]")
    tiles = tiles.contiguous()
    m = tiles.shape[0]
    deltas = torch.zeros((m, 4), dtype=torch.float32, device=tiles.device)
    mx = torch.zeros((m,), dtype=torch.float32, device=tiles.device)
    _lib.call("unmore_update_bbox_from_tiles", tiles.data_ptr(), m, deltas.data_ptr(), mx.data_ptr(), _stream())
    return deltas, mx


MODE_FLAGS, MODE_SCORE_GE, MODE_LABEL_EQ, MODE_ARGMAX_GE0, MODE_ARGMAX_LT0 = range(5)


def compact_boxes(inp, counts_in, mode, pred, thr=0.0, group=1, out=None, counts_out=None, cap_out=None,
                  out_dtype=None, append=False, want_index=False):
    """Stable per-image selection; returns (out [n_img, cap_out, 4], counts_out [n_img], index or None)."""
    dev = inp.device
    if group == 1:
        n_img, cap_in = inp.shape[0], inp.shape[1]
    else:
        n_img, cap_in = inp.shape[0], inp.shape[1]
        assert inp.shape[2] == group
    out_dtype = out_dtype or inp.dtype
    if cap_out is None:
        cap_out = out.shape[1] if out is not None else cap_in * group
    if out is None:
        out = torch.zeros((n_img, cap_out, 4), dtype=out_dtype, device=dev)
    if counts_out is None:
        counts_out = torch.zeros((n_img,), dtype=torch.int32, device=dev)
    index = torch.full((n_img, cap_out), -1, dtype=torch.int32, device=dev) if want_index else None
    _lib.call("unmore_compact_boxes", inp.data_ptr(), int(inp.dtype == torch.float64), _ptr(counts_in), cap_in,
              group, mode, pred.data_ptr(), float(thr), out.data_ptr(), int(out.dtype == torch.float64), cap_out,
              counts_out.data_ptr(), int(append), _ptr(index), n_img, _stream())
    return out, counts_out, index


def box_nms(boxes, scores=None, counts=None, iou_threshold: float = 0.5, want_boxes: bool = True):
    """boxes [n_img, cap, 4] fp32 -> (keep [n_img, cap] int32, keep_counts [n_img], kept boxes or None)."""
    if not boxes.is_cuda or boxes.dtype != torch.float32 or boxes.dim() != 3 or not boxes.is_contiguous():
        raise _lib.UnmoreError("boxes must be contiguous CUDA fp32 [n_img, cap, 4]")
    n_img, cap = boxes.shape[0], boxes.shape[1]
    dev = boxes.device
    keep = torch.full((n_img, cap), -1, dtype=torch.int32, device=dev)
    kc = torch.zeros((n_img,), dtype=torch.int32, device=dev)
    order = torch.empty((n_img, max(cap, 1)), dtype=torch.int32, device=dev)
    kb = torch.zeros((n_img, cap, 4), dtype=torch.float32, device=dev) if want_boxes else None
    if scores is not None:
        scores = scores.contiguous().to(torch.float32)
    _lib.call("unmore_box_nms", boxes.data_ptr(), _ptr(scores), _ptr(counts), cap, n_img, float(iou_threshold),
              keep.data_ptr(), kc.data_ptr(), _ptr(kb), order.data_ptr(), _stream())
    return keep, kc, kb
