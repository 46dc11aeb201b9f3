"""Mirror of the reference's ``utils/misc.py`` hot-path helper."""
from __future__ import annotations

import torch

from .. import ops


def batch_erode(binary_masks: torch.Tensor, kernel_size: int = 9, num_round: int = 3) -> torch.Tensor:
    """utils/misc.py:10-20 — [B,128,128] integer masks -> int64 {0,1}, eroded ``num_round`` times with a
    ``kernel_size`` ones kernel and zero border (bit-parallel on the GPU instead of a float64 conv)."""
    out = ops.batch_erode((binary_masks != 0).to(torch.uint8), kernel_size, num_round)
    return out.to(torch.int64)
