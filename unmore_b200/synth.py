"""Synthetic objectness-field scenes and proposal sets.

The reference has no per-image fields: its nets run on every crop
(object_reasoning.py:398-417).  Under the field-stub bridge the
``image`` handed to the reasoning code is the ``[4, H, W]`` field stack
``[sdf, center_row, center_col, existence]`` and the nets are channel selectors.
This module builds such stacks deterministically from an image index so the CPU
oracle, the parity tests and ``bench.py`` all see the same inputs.

Channel conventions follow the reference's training targets:
  * center field = unit vector from the owning object's centre to the pixel,
    (row, col) order, zero outside objects (datasets.py:200-212);
  * boundary-distance field ("sdf") positive inside, negative outside
    (datasets.py:187-195 with ``use_bg_sdf``), squashed by tanh like the
    ``sdf_activation='tanh'`` head (models/objectness_net.py:128-135).
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np
import torch

FIELD_CHANNELS = 4  # [sdf, center_row, center_col, existence]
CH_SDF, CH_CROW, CH_CCOL, CH_EXIST = 0, 1, 2, 3


def scene_params(index: int, height: int, width: int) -> torch.Tensor:
    """Object list for image ``index``: ``[K, 4]`` rows ``(cy, cx, ry, rx)`` (float32, CPU).

    K ~ U{3..12}; centres uniform in the image; radii U[30, 90] px (ry, rx drawn
    independently, so objects are axis-aligned ellipses).
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(int(index))
    k = int(torch.randint(3, 13, (1,), generator=g).item())
    u = torch.rand((k, 4), generator=g, dtype=torch.float32)
    cy = u[:, 0] * height
    cx = u[:, 1] * width
    ry = 30.0 + 60.0 * u[:, 2]
    rx = 30.0 + 60.0 * u[:, 3]
    return torch.stack([cy, cx, ry, rx], dim=1)


def render_fields(params: torch.Tensor, height: int, width: int, device=None) -> torch.Tensor:
    """Render one ``[4, H, W]`` float32 field stack from ``scene_params`` rows.

    d_k = r_eff_k * (1 - rho_k) with rho_k the normalised elliptical radius, i.e. a
    signed "distance-like" value that is r_eff at the centre, 0 on the outline and
    negative outside; d = max_k d_k picks the owning object.
    """
    device = torch.device(device) if device is not None else params.device
    p = params.to(device=device, dtype=torch.float32)
    ys = torch.arange(height, device=device, dtype=torch.float32).view(1, height, 1)
    xs = torch.arange(width, device=device, dtype=torch.float32).view(1, 1, width)
    cy, cx, ry, rx = (p[:, i].view(-1, 1, 1) for i in range(4))
    dy = ys - cy
    dx = xs - cx
    rho = torch.sqrt((dy / ry) ** 2 + (dx / rx) ** 2)
    reff = torch.minimum(ry, rx)
    dk = reff * (1.0 - rho)                      # [K, H, W]
    d, owner = dk.max(dim=0)                     # [H, W]
    inside = d > 0
    oy = torch.gather(dy.expand(-1, -1, width), 0, owner.unsqueeze(0))[0]
    ox = torch.gather(dx.expand(-1, height, -1), 0, owner.unsqueeze(0))[0]
    nrm = torch.sqrt(oy * oy + ox * ox).clamp_min(1e-12)
    crow = torch.where(inside, oy / nrm, torch.zeros_like(oy))
    ccol = torch.where(inside, ox / nrm, torch.zeros_like(ox))
    sdf = torch.tanh(d / 40.0)
    exist = torch.sigmoid(d / 10.0)
    return torch.stack([sdf, crow, ccol, exist], dim=0).contiguous()


def make_fields(index: int, height: int = 480, width: int = 640, device="cpu") -> torch.Tensor:
    """Field stack of image ``index`` (seed = index)."""
    return render_fields(scene_params(index, height, width), height, width, device=device)


def make_field_batch(indices, height: int = 480, width: int = 640, device="cpu") -> torch.Tensor:
    """``[B, 4, H, W]`` float32 batch for the given image indices."""
    return torch.stack([make_fields(int(i), height, width, device=device) for i in indices], dim=0)


def anchor_proposals(height: int, width: int) -> np.ndarray:
    """Anchor grid with the semantics of ``Object_Discovery.generate_random_proposal``
    (object_reasoning.py:110-137): grid sizes 32..512, three shapes per centre
    (2g x 2g, g x 2g, 2g x g), clipped to the image, full-image box appended.
    Returns float64 ``[N, 4]`` xyxy; N = 1225 at 480x640, 4093 at 1024x1024.
    """
    chunks = []
    for g in (32, 64, 128, 256, 512):
        cys = np.arange(0, height, g, dtype=np.int64)
        cxs = np.arange(0, width, g, dtype=np.int64)
        cx, cy = np.meshgrid(cxs, cys)
        ctr = np.stack([cx.ravel(), cy.ravel(), cx.ravel(), cy.ravel()], axis=1).astype(np.float64)
        shapes = np.array([[-g, -g, g, g], [-g / 2, -g, g / 2, g], [-g, -g / 2, g, g / 2]], dtype=np.float64)
        chunks.append((ctr[:, None, :] + shapes[None, :, :]).reshape(-1, 4))
    out = np.concatenate(chunks, axis=0)
    out[:, 0] = np.where(out[:, 0] < 0, 0.0, out[:, 0])
    out[:, 1] = np.where(out[:, 1] < 0, 0.0, out[:, 1])
    out[:, 2] = np.where(out[:, 2] >= width, float(width), out[:, 2])
    out[:, 3] = np.where(out[:, 3] >= height, float(height), out[:, 3])
    return np.concatenate([out, np.array([[0.0, 0.0, float(width), float(height)]])], axis=0)


def random_proposals(index: int, count: int, height: int, width: int) -> np.ndarray:
    """``count`` extra boxes: log-uniform side in [16, 512] per axis, uniform centres,
    clipped to the image.  float64 ``[count, 4]`` xyxy."""
    if count <= 0:
        return np.zeros((0, 4), dtype=np.float64)
    g = torch.Generator(device="cpu")
    g.manual_seed(1_000_003 + int(index))
    u = torch.rand((count, 4), generator=g, dtype=torch.float64).numpy()
    lo, hi = math.log(16.0), math.log(512.0)
    w = np.exp(lo + (hi - lo) * u[:, 0])
    h = np.exp(lo + (hi - lo) * u[:, 1])
    cx = u[:, 2] * width
    cy = u[:, 3] * height
    box = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1)
    box[:, 0] = np.clip(box[:, 0], 0.0, float(width))
    box[:, 1] = np.clip(box[:, 1], 0.0, float(height))
    box[:, 2] = np.clip(box[:, 2], 0.0, float(width))
    box[:, 3] = np.clip(box[:, 3], 0.0, float(height))
    return box


def make_proposals(index: int, count: int, height: int = 480, width: int = 640) -> np.ndarray:
    """Exactly ``count`` float64 proposals for image ``index``: the anchor grid, then
    random boxes as padding (a strided anchor subset plus random boxes if ``count`` is
    smaller, always keeping the full-image box last like the reference does)."""
    anchors = anchor_proposals(height, width)
    n = anchors.shape[0]
    if count == n:
        return anchors
    if count < n:
        # three quarters strided anchors (all grid scales stay represented), one quarter
        # random fractional boxes, full-image box last
        n_anchor = max(1, (count * 3) // 4 - 1)
        sel = np.unique(np.linspace(0, n - 2, n_anchor).round().astype(np.int64))
        extra = random_proposals(index, count - 1 - sel.shape[0], height, width)
        return np.concatenate([anchors[sel], extra, anchors[-1:]], axis=0)
    return np.concatenate([anchors, random_proposals(index, count - n, height, width)], axis=0)


def shapes_for(config: str) -> Tuple[int, int, int]:
    """(H, W, proposals/image) of the BASELINE.json configs."""
    table = {
        "config0": (480, 640, 512),
        "config1": (480, 640, 4096),
        "config4": (1024, 1024, 4093),
    }
    return table[config]
