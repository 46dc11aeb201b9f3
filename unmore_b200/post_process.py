"""Mirror of the reference's ``post_process.py`` filter (post_process.py:59-74).

The predicate is three comparisons per annotation; for device-resident batches it is
evaluated inside ``unmore_final_scores`` (the ``selected`` flags).  This host form keeps the
reference's list-of-dicts interface."""
from __future__ import annotations

from typing import List


def select_annotations(pred_annotations: List[dict], existence_score_thres: float = 0.5,
                       center_score_thres: float = 0.8, boundary_score_thres: float = 0.75) -> List[dict]:
    selected = []
    for ann in pred_annotations:
        if ann["existence_score"] < existence_score_thres:
            continue
        if ann["center_score"] < center_score_thres:
            continue
        if ann["boundary_score"] < boundary_score_thres:
            continue
        ann = dict(ann)
        ann["id"] = len(selected)
        ann["score"] = ann["area_score"]
        selected.append(ann)
    return selected


def convert_pred_annotations_to_training_format(selected_annotations_list, image_info_list):
    """post_process.py:11-32 without the file I/O."""
    return {"categories": {"is_crowd": 0, "id": 1}, "images": image_info_list, "annotations": selected_annotations_list}
