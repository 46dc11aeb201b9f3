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


CATEGORIES = {"is_crowd": 0, "id": 1}   # post_process.py:6-9


def convert_pred_annotations_to_training_format(selected_annotations_list, gt_annotation_path, out_fname_training=None):
    """post_process.py:11-32: attach the ground-truth ``images`` list and dump the training dict to
    ``out_fname_training``.  ``gt_annotation_path`` may also be the ``images`` list itself (no GT file on this
    box), and ``out_fname_training=None`` skips the dump; the dict is returned either way."""
    import json
    if isinstance(gt_annotation_path, (str, bytes)):
        with open(gt_annotation_path) as f:
            image_info_list = json.load(f)["images"]
    else:
        image_info_list = gt_annotation_path
    training_annotations = {"categories": CATEGORIES, "images": image_info_list, "annotations": selected_annotations_list}
    if out_fname_training is not None:
        with open(out_fname_training, "w") as f:
            json.dump(training_annotations, f, indent=2)
    return training_annotations


def main(argv=None):
    """post_process.py ``__main__`` (:35-77): filter ``--pred_annotations_path`` with the three thresholds and
    write ``selected_training_annotations.json`` next to it.  ``--gt_annotation_path`` replaces the reference's
    hard-coded COCO paths (:50-55); without it the ``images`` list is empty."""
    import argparse, json, os
    ap = argparse.ArgumentParser()
    ap.add_argument("--pred_annotations_path", type=str, default=None)
    ap.add_argument("--existence_score_thres", type=float, default=0.5)
    ap.add_argument("--center_score_thres", type=float, default=0.8)
    ap.add_argument("--boundary_score_thres", type=float, default=0.75)
    ap.add_argument("--dataset", type=str, default="COCO")
    ap.add_argument("--split", type=str, default="test")
    ap.add_argument("--gt_annotation_path", type=str, default=None)
    args = ap.parse_args(argv)
    result_folder = os.path.dirname(args.pred_annotations_path)
    with open(os.path.join(result_folder, "configs_post_process.json"), "w") as f:
        json.dump(args.__dict__, f, indent=2)
    with open(args.pred_annotations_path) as f:
        pred = json.load(f)
    sel = select_annotations(pred, args.existence_score_thres, args.center_score_thres, args.boundary_score_thres)
    out = os.path.join(result_folder, "selected_training_annotations.json")
    return convert_pred_annotations_to_training_format(sel, args.gt_annotation_path or [], out)


if __name__ == "__main__":
    main()
