"""Known-answer tests of the COCO AP / AR restatement (unmore_b200/coco_eval.py).

pycocotools / detectron2 are not available, so every expected value here is derived by hand from the
published COCOeval protocol (101 recall thresholds, IoU .50:.05:.95, area ranges 32^2 / 96^2, maxDets
1/10/100) — parity unpinned, see the module docstring."""
import math

import numpy as np
import pytest

from unmore_b200 import coco_eval as C


def _gt(anns, h=200, w=200, n_img=1):
    return {"images": [{"id": i + 1, "height": h, "width": w} for i in range(n_img)],
            "annotations": [dict(a, id=k + 1, category_id=1, iscrowd=a.get("iscrowd", 0)) for k, a in enumerate(anns)]}


def _box_ann(img, x, y, w, h, **kw):
    return dict(image_id=img, bbox=[x, y, w, h], area=float(w * h), **kw)


def test_perfect_predictions_score_100():
    gt = _gt([_box_ann(1, 10, 10, 40, 40), _box_ann(1, 100, 100, 20, 20), _box_ann(1, 20, 80, 120, 110)])
    ev = C.COCOEvaluator(gt, tasks=("bbox",))
    ev.process(1, [dict(image_id=1, bbox=a["bbox"], score=0.9 - 0.1 * k) for k, a in enumerate(gt["annotations"])])
    r = ev.evaluate()["bbox"]
    for m in ("AP", "AP50", "AP75", "APs", "APm", "APl", "AR10", "AR100", "ARs", "ARm", "ARl"):
        assert r[m] == pytest.approx(100.0, abs=1e-9), m
    assert r["AR1"] == pytest.approx(100.0 / 3)   # one detection per image can recall one of three objects


def test_one_hit_one_miss_one_false_positive():
    """TP (0.9) on the medium object, FP (0.8), the small object missed: recall stops at 0.5, precision 1 up
    to there -> AP = 51/101 (thresholds 0.00 .. 0.50)."""
    gt = _gt([_box_ann(1, 10, 10, 40, 40), _box_ann(1, 60, 60, 20, 20)])
    ev = C.COCOEvaluator(gt, tasks=("bbox",))
    ev.process(1, [dict(image_id=1, bbox=[10, 10, 40, 40], score=0.9), dict(image_id=1, bbox=[0, 150, 5, 5], score=0.8)])
    r = ev.evaluate()["bbox"]
    assert r["AP"] == pytest.approx(100 * 51 / 101)
    assert r["AP50"] == pytest.approx(100 * 51 / 101) and r["AP75"] == pytest.approx(100 * 51 / 101)
    assert r["AR100"] == pytest.approx(50.0) and r["AR1"] == pytest.approx(50.0)
    assert r["APm"] == pytest.approx(100.0) and r["ARm"] == pytest.approx(100.0)   # 40x40 is medium
    assert r["APs"] == pytest.approx(0.0) and r["ARs"] == pytest.approx(0.0)       # the 20x20 object is never found
    assert math.isnan(r["APl"]) and math.isnan(r["ARl"])                           # no large object at all


def test_false_positive_ranked_first_halves_precision():
    gt = _gt([_box_ann(1, 10, 10, 40, 40)])
    ev = C.COCOEvaluator(gt, tasks=("bbox",))
    ev.process(1, [dict(image_id=1, bbox=[100, 100, 40, 40], score=0.9), dict(image_id=1, bbox=[10, 10, 40, 40], score=0.8)])
    r = ev.evaluate()["bbox"]
    assert r["AP"] == pytest.approx(50.0) and r["AR100"] == pytest.approx(100.0)
    assert r["AR1"] == pytest.approx(0.0)          # the single allowed detection is the false positive


def test_iou_threshold_sweep():
    """IoU 0.62 -> a match at thresholds 0.50, 0.55, 0.60 only: AP = 3/10, AP50 = 1, AP75 = 0."""
    gt = _gt([_box_ann(1, 0, 0, 100, 62)])
    ev = C.COCOEvaluator(gt, tasks=("bbox",))
    ev.process(1, [dict(image_id=1, bbox=[0, 0, 100, 100], score=1.0)])
    r = ev.evaluate()["bbox"]
    assert r["AP"] == pytest.approx(30.0) and r["AP50"] == pytest.approx(100.0) and r["AP75"] == pytest.approx(0.0)
    assert r["AR100"] == pytest.approx(30.0)


def test_crowd_regions_are_ignored_not_penalised():
    """A detection inside a crowd region is neither TP nor FP; the crowd itself is not a recall target."""
    gt = _gt([_box_ann(1, 10, 10, 40, 40), _box_ann(1, 100, 100, 80, 80, iscrowd=1)])
    ev = C.COCOEvaluator(gt, tasks=("bbox",))
    ev.process(1, [dict(image_id=1, bbox=[10, 10, 40, 40], score=0.9),
                   dict(image_id=1, bbox=[110, 110, 20, 20], score=0.95),     # inside the crowd: IoU = inter / det area = 1
                   dict(image_id=1, bbox=[120, 140, 20, 20], score=0.85)])    # crowds can absorb several detections
    r = ev.evaluate()["bbox"]
    assert r["AP"] == pytest.approx(100.0) and r["AR100"] == pytest.approx(100.0)


def test_max_detections_and_multiple_images():
    """Two images; the second image's only correct detection is its 2nd-ranked one, so AR1 sees one of the two objects."""
    gt = _gt([_box_ann(1, 10, 10, 50, 50), _box_ann(2, 30, 30, 50, 50)], n_img=2)
    ev = C.COCOEvaluator(gt, tasks=("bbox",))
    ev.process(1, [dict(image_id=1, bbox=[10, 10, 50, 50], score=0.9)])
    ev.process(2, [dict(image_id=2, bbox=[120, 120, 50, 50], score=0.8), dict(image_id=2, bbox=[30, 30, 50, 50], score=0.7)])
    r = ev.evaluate()["bbox"]
    assert r["AR1"] == pytest.approx(50.0) and r["AR10"] == pytest.approx(100.0)
    # ranked: TP(.9) FP(.8) TP(.7): precision 1 up to recall .5, then 2/3 -> AP = (51 + 50 * 2/3) / 101
    assert r["AP"] == pytest.approx(100 * (51 + 50 * 2 / 3) / 101)
    with pytest.raises(ValueError):
        C.COCOEvaluator(gt, max_dets_per_image=[100])


def test_mask_task_uses_mask_iou_and_mask_area():
    h = w = 64
    a = np.zeros((h, w), np.uint8); a[8:40, 8:40] = 1            # 32x32 = 1024 px -> medium (>= 32^2)
    b = np.zeros((h, w), np.uint8); b[8:40, 8:24] = 1            # half of it: IoU 0.5
    gt = {"images": [{"id": 1, "height": h, "width": w}],
          "annotations": [dict(id=1, image_id=1, category_id=1, iscrowd=0, area=1024.0, bbox=[8, 8, 32, 32],
                               segmentation=[[8, 8, 40, 8, 40, 40, 8, 40]])]}
    ev = C.COCOEvaluator(gt, tasks=("segm", "bbox"))
    ev.process(1, [dict(image_id=1, score=0.5, bbox=[8, 8, 32, 32], segmentation={"mask": b})])
    r = ev.evaluate()
    assert r["bbox"]["AP"] == pytest.approx(100.0)
    assert r["segm"]["AP50"] == pytest.approx(100.0) and r["segm"]["AP75"] == pytest.approx(0.0)
    assert r["segm"]["AP"] == pytest.approx(10.0)                # only the 0.50 threshold matches
    assert np.array_equal(C.segmentation_to_mask(gt["annotations"][0]["segmentation"], h, w), a)


def test_polygon_rasterisation_and_rle_round_trip():
    from unmore_b200 import rle
    m = C.poly_to_mask([1, 1, 4, 1, 4, 4, 1, 4], 6, 6)           # integer corners cover [1,4) x [1,4)
    assert m.sum() == 9 and m[1:4, 1:4].all()
    tri = C.poly_to_mask([10, 10, 50, 12, 30, 40], 64, 64)
    assert abs(int(tri.sum()) - 580) <= 8                        # polygon area 580 px^2
    mask = np.zeros((37, 53), np.uint8); mask[5:20, 7:30] = 1; mask[25:30, 40:50] = 1
    runs, v, prev = [], 0, 0
    flat = mask.T.reshape(-1)
    edges = np.concatenate([[0], np.nonzero(np.diff(np.concatenate([[0], flat])))[0], [flat.size]])
    enc = {"size": [37, 53], "counts": rle.counts_to_string(np.diff(edges))}
    assert np.array_equal(C.segmentation_to_mask(enc, 37, 53), mask)
    del runs, v, prev


def test_evaluate_ap_files(tmp_path):
    import json
    gt = _gt([_box_ann(1, 10, 10, 40, 40)])
    gt["annotations"][0]["segmentation"] = [[10, 10, 50, 10, 50, 50, 10, 50]]
    preds = [dict(image_id=1, bbox=[10, 10, 40, 40], weight=0.7, segmentation=[[10, 10, 50, 10, 50, 50, 10, 50]])]
    (tmp_path / "gt.json").write_text(json.dumps(gt))
    (tmp_path / "pred.json").write_text(json.dumps(preds))
    out = C.evaluate_ap(str(tmp_path / "gt.json"), str(tmp_path / "pred.json"))
    assert out["bbox"]["AP"] == pytest.approx(100.0) and out["segm"]["AP"] == pytest.approx(100.0)
    assert out["number_of_images"] == 1 and out["number_of_annotations"] == 1
