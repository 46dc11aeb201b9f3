"""COCO run-length encoding of binary masks and the two JSON result files of the reference.

``binary_mask_to_rle`` (object_scoring.py:167-170) calls pycocotools' ``mask.encode``; pycocotools
2.0.7 is not available here, so the codec is restated from its published C source (maskApi.c:
rleEncode, rleToString, rleFrString, rleDecode).  Run lengths come from the GPU
(``ops.mask_rle_counts`` on the bit-packed masks the scorer already holds); only the ASCII
compression of a few hundred integers per mask runs on the host.
"""
from __future__ import annotations

import json
from typing import Dict, Iterable, List

import numpy as np
import torch

from . import ops


def counts_to_string(counts: Iterable[int]) -> str:
    """maskApi.c rleToString: like LEB128 with 6 bits per char (ASCII 48..111); from the third count
    on the difference to the count two places earlier is stored."""
    cnts = [int(c) for c in counts]
    out = []
    for i, c in enumerate(cnts):
        x = c - cnts[i - 2] if i > 2 else c
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(chr(ch + 48))
    return "".join(out)


def string_to_counts(s: str) -> List[int]:
    """maskApi.c rleFrString."""
    cnts: List[int] = []
    p = 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = ord(s[p]) - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(cnts) > 2:
            x += cnts[-2]
        cnts.append(x)
    return cnts


def decode(rle: dict) -> np.ndarray:
    """{'size': [H, W], 'counts': str} -> uint8 [H, W] (maskApi.c rleDecode, column-major)."""
    h, w = rle["size"]
    cnts = string_to_counts(rle["counts"]) if isinstance(rle["counts"], str) else list(rle["counts"])
    flat = np.zeros(h * w, dtype=np.uint8)
    pos, v = 0, 0
    for c in cnts:
        if v:
            flat[pos:pos + c] = 1
        pos += c
        v ^= 1
    return flat.reshape(w, h).T.copy()


def encode_packed(packed: torch.Tensor, width: int, max_runs: int = 4096) -> List[dict]:
    """Bit-packed masks [K, H, ceil(W/32)] (device) -> [{'size': [H, W], 'counts': str}, ...] like
    ``binary_mask_to_rle`` after its ``.decode('ascii')``."""
    K, H, _ = packed.shape
    if K == 0:
        return []
    counts, n_runs = ops.mask_rle_counts(packed, width, max_runs)
    n = n_runs.cpu().numpy()
    limit = 8192   # bound of the device buffer; masks with more runs (pathological) are encoded on the host below
    if max_runs < limit and int(n.max()) > max_runs:
        return encode_packed(packed, width, max_runs=limit)
    c = counts.cpu().numpy().view(np.uint32)
    out = []
    for k in range(K):
        if n[k] <= min(max_runs, limit):
            out.append({"size": [H, width], "counts": counts_to_string(c[k, : n[k]])})
        else:  # pathological mask (> 8192 runs): run lengths on the host from the unpacked bits
            bits = np.unpackbits(packed[k].cpu().numpy().view(np.uint8), axis=-1, bitorder="little")[:, :width]
            m = bits.T.reshape(-1)
            edges = np.concatenate([[0], np.nonzero(np.diff(np.concatenate([[0], m])))[0], [m.size]])
            out.append({"size": [H, width], "counts": counts_to_string(np.diff(edges))})
    return out


def discovery_results_json(results: Dict[int, np.ndarray]) -> str:
    """discovery_results.json (object_reasoning.py:662-665): {image_id: [[x1, y1, x2, y2], ...]}."""
    return json.dumps({str(k): np.asarray(v, dtype=np.float64).tolist() for k, v in results.items()}, indent=2)


def scored_annotations_json(annotations: List[dict]) -> str:
    """object_discovery_with_scores.json (object_scoring.py:257-272); numpy scalars become floats
    the way the reference's NpEncoder makes them."""
    def clean(v):
        if isinstance(v, (np.floating, np.integer)):
            return v.item()
        if isinstance(v, np.ndarray):
            return v.tolist()
        if isinstance(v, (list, tuple)):
            return [clean(x) for x in v]
        if isinstance(v, dict):
            return {k: clean(x) for k, x in v.items()}
        return v
    return json.dumps([clean(a) for a in annotations], indent=2)
