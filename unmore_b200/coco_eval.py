"""Class-agnostic COCO AP / AR for the path's outputs.

The reference scores ``object_discovery_with_scores.json`` / detector outputs with
``COCO_evaluator/main.py:24-70`` -> ``COCOEvaluator`` (COCO_evaluator/coco_evaluation.py:37-410), which is
pycocotools' ``COCOeval`` driven through detectron2's C++ ``COCOeval_opt`` (fast_eval_api.py).  Neither
pycocotools 2.0.7 nor detectron2 exists here, so this module restates the *published* COCOeval protocol
(pycocotools/cocoeval.py: evaluateImg / accumulate / summarize, and maskApi.c: rleFrPoly, rleIou, bbIou)
in numpy — **parity unpinned** (no copy of the dependency to run; pinned by hand-derived known answers in
tests/test_coco_eval.py).  It is host-side metric code, deliberately outside the CUDA path.

Interface mirrors the reference's evaluator: ``COCOEvaluator(gt, tasks=("bbox", "segm"))``, ``reset()``,
``process(image_id, coco_instances)``, ``evaluate() -> {"bbox": {...}, "segm": {...}}`` with the twelve
numbers of coco_evaluation.py:349-353 (AP, AP50, AP75, APs, APm, APl, AR1, AR10, AR100, ARs, ARm, ARl),
scaled by 100, NaN where undefined.
"""
from __future__ import annotations

import json
import math
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import rle as _rle

METRICS = ["AP", "AP50", "AP75", "APs", "APm", "APl", "AR1", "AR10", "AR100", "ARs", "ARm", "ARl"]
IOU_THRS = np.linspace(0.5, 0.95, int(np.round((0.95 - 0.5) / 0.05)) + 1, endpoint=True)
REC_THRS = np.linspace(0.0, 1.00, int(np.round((1.00 - 0.0) / 0.01)) + 1, endpoint=True)
AREA_RNG = [[0 ** 2, 1e5 ** 2], [0 ** 2, 32 ** 2], [32 ** 2, 96 ** 2], [96 ** 2, 1e5 ** 2]]
AREA_LBL = ["all", "small", "medium", "large"]


# ---------------------------------------------------------------------------------------------
# masks: polygon / RLE -> dense, areas, IoU (maskApi.c semantics)
# ---------------------------------------------------------------------------------------------
def poly_to_mask(xy: Sequence[float], h: int, w: int) -> np.ndarray:
    """maskApi.c rleFrPoly restated: the polygon is traced on a 5x up-sampled grid, the column crossings of
    its boundary are collected, down-sampled and turned into column-major runs.  Returns uint8 [h, w]."""
    k = len(xy) // 2
    scale = 5.0
    x = [int(scale * xy[2 * j] + 0.5) for j in range(k)]
    y = [int(scale * xy[2 * j + 1] + 0.5) for j in range(k)]
    x.append(x[0]); y.append(y[0])
    u: List[int] = []
    v: List[int] = []
    for j in range(k):
        xs, xe, ys, ye = x[j], x[j + 1], y[j], y[j + 1]
        dx, dy = abs(xe - xs), abs(ys - ye)
        flip = (dx >= dy and xs > xe) or (dx < dy and ys > ye)
        if flip:
            xs, xe, ys, ye = xe, xs, ye, ys
        s = (ye - ys) / dx if dx >= dy and dx else ((xe - xs) / dy if dy else 0.0)
        if dx >= dy:
            for d in range(dx + 1):
                t = dx - d if flip else d
                u.append(t + xs); v.append(int(ys + s * t + 0.5))
        else:
            for d in range(dy + 1):
                t = dy - d if flip else d
                v.append(t + ys); u.append(int(xs + s * t + 0.5))
    pts = []
    for j in range(1, len(u)):
        if u[j] != u[j - 1]:
            xd = float(u[j] if u[j] < u[j - 1] else u[j] - 1)
            xd = (xd + 0.5) / scale - 0.5
            if math.floor(xd) != xd or xd < 0 or xd > w - 1:
                continue
            yd = float(v[j] if v[j] < v[j - 1] else v[j - 1])
            yd = (yd + 0.5) / scale - 0.5
            yd = 0.0 if yd < 0 else (float(h) if yd > h else yd)
            pts.append(int(xd) * h + int(math.ceil(yd)))
    a = sorted(pts + [h * w])
    diffs = np.diff(np.asarray([0] + a, dtype=np.int64))
    runs: List[int] = []
    j = 0
    runs.append(int(diffs[0])); j = 1
    while j < len(diffs):
        if diffs[j] > 0:
            runs.append(int(diffs[j])); j += 1
        else:
            j += 1
            if j < len(diffs):
                runs[-1] += int(diffs[j]); j += 1
    return _rle.decode({"size": [h, w], "counts": runs})


def segmentation_to_mask(seg, h: int, w: int) -> np.ndarray:
    """COCO 'segmentation' (polygon list, uncompressed or compressed RLE, or a dense array) -> uint8 [h, w]."""
    if isinstance(seg, np.ndarray):
        return (seg != 0).astype(np.uint8)
    if isinstance(seg, dict):
        if "mask" in seg:   # this repo's in-memory annotations (Object_Scoring.score_image)
            return (np.asarray(seg["mask"]) != 0).astype(np.uint8)
        cnt = seg["counts"]
        if isinstance(cnt, bytes):
            cnt = cnt.decode("ascii")
        return _rle.decode({"size": seg["size"], "counts": cnt})
    out = np.zeros((h, w), dtype=np.uint8)   # list of polygons: union (pycocotools merges them)
    for poly in seg:
        out |= poly_to_mask(poly, h, w)
    return out


def bbox_iou(dt: np.ndarray, gt: np.ndarray, iscrowd: np.ndarray) -> np.ndarray:
    """maskApi.c bbIou on xywh boxes: [D, G]; for a crowd gt the union is the detection's area."""
    if len(dt) == 0 or len(gt) == 0:
        return np.zeros((len(dt), len(gt)))
    d = np.asarray(dt, dtype=np.float64)[:, None, :]
    g = np.asarray(gt, dtype=np.float64)[None, :, :]
    iw = np.minimum(d[..., 0] + d[..., 2], g[..., 0] + g[..., 2]) - np.maximum(d[..., 0], g[..., 0])
    ih = np.minimum(d[..., 1] + d[..., 3], g[..., 1] + g[..., 3]) - np.maximum(d[..., 1], g[..., 1])
    inter = np.clip(iw, 0, None) * np.clip(ih, 0, None)
    da, ga = d[..., 2] * d[..., 3], g[..., 2] * g[..., 3]
    union = np.where(np.asarray(iscrowd, dtype=bool)[None, :], da, da + ga - inter)
    return np.where(union > 0, inter / np.where(union > 0, union, 1), 0.0)


def mask_iou(dt: np.ndarray, gt: np.ndarray, iscrowd: np.ndarray) -> np.ndarray:
    """maskApi.c rleIou on dense masks [D, H, W], [G, H, W] -> [D, G]."""
    if len(dt) == 0 or len(gt) == 0:
        return np.zeros((len(dt), len(gt)))
    d = dt.reshape(len(dt), -1).astype(np.float32)
    g = gt.reshape(len(gt), -1).astype(np.float32)
    inter = (d @ g.T).astype(np.float64)
    da, ga = d.sum(1, dtype=np.float64)[:, None], g.sum(1, dtype=np.float64)[None, :]
    union = np.where(np.asarray(iscrowd, dtype=bool)[None, :], da, da + ga - inter)
    return np.where(union > 0, inter / np.where(union > 0, union, 1), 0.0)


# ---------------------------------------------------------------------------------------------
# COCOeval protocol
# ---------------------------------------------------------------------------------------------
def _evaluate_img(ious: np.ndarray, gt_area, gt_crowd, gt_ignore_flag, dt_score, dt_area, a_rng, max_det):
    """cocoeval.py evaluateImg for one image, one area range.  `ious` is [D, G] for detections sorted by
    descending score (top max_det).  Returns (dtMatches [T, D] bool, dtIgnore [T, D] bool, gtIgnore [G])."""
    G, D = len(gt_area), len(dt_score)
    g_ig = np.array([bool(gt_ignore_flag[i]) or gt_area[i] < a_rng[0] or gt_area[i] > a_rng[1] for i in range(G)], dtype=bool)
    gtind = np.argsort(g_ig, kind="mergesort")   # not-ignored first, order kept otherwise
    g_ig = g_ig[gtind]
    crowd = np.asarray(gt_crowd, dtype=bool)[gtind] if G else np.zeros(0, dtype=bool)
    ious = ious[:max_det][:, gtind] if ious.size else ious
    D = min(D, max_det)
    T = len(IOU_THRS)
    gtm = -np.ones((T, G), dtype=np.int64)
    dtm = -np.ones((T, D), dtype=np.int64)
    dt_ig = np.zeros((T, D), dtype=bool)
    for ti, t in enumerate(IOU_THRS):
        for di in range(D):
            iou = min(t, 1 - 1e-10)
            m = -1
            for gi in range(G):
                if gtm[ti, gi] >= 0 and not crowd[gi]:
                    continue                      # already matched, and not a crowd
                if m > -1 and not g_ig[m] and g_ig[gi]:
                    break                         # matched a regular gt, only ignored ones follow
                if ious[di, gi] < iou:
                    continue
                iou = ious[di, gi]
                m = gi
            if m == -1:
                continue
            dt_ig[ti, di] = g_ig[m]
            dtm[ti, di] = m
            gtm[ti, m] = di
    out_of_range = np.array([dt_area[i] < a_rng[0] or dt_area[i] > a_rng[1] for i in range(D)], dtype=bool)
    dt_ig = dt_ig | ((dtm < 0) & out_of_range[None, :])
    return dtm >= 0, dt_ig, g_ig


def _accumulate(per_image, max_dets):
    """cocoeval.py accumulate for one category.  per_image[a][i] = (scores [D], dtMatched [T,D], dtIgnore [T,D],
    gtIgnore [G]) of image i in area range a, detections sorted by score, capped at max(max_dets)."""
    T, R, A, M = len(IOU_THRS), len(REC_THRS), len(AREA_RNG), len(max_dets)
    precision = -np.ones((T, R, A, M))
    recall = -np.ones((T, A, M))
    for a in range(A):
        E = per_image[a]
        if not E:
            continue
        for mi, max_det in enumerate(max_dets):
            scores = np.concatenate([e[0][:max_det] for e in E])
            inds = np.argsort(-scores, kind="mergesort")
            dtm = np.concatenate([e[1][:, :max_det] for e in E], axis=1)[:, inds]
            dt_ig = np.concatenate([e[2][:, :max_det] for e in E], axis=1)[:, inds]
            npig = int(sum(np.count_nonzero(~e[3]) for e in E))
            if npig == 0:
                continue
            tps = np.cumsum(dtm & ~dt_ig, axis=1, dtype=np.float64)
            fps = np.cumsum(~dtm & ~dt_ig, axis=1, dtype=np.float64)
            for t in range(T):
                tp, fp = tps[t], fps[t]
                nd = len(tp)
                rc = tp / npig
                pr = tp / (fp + tp + np.spacing(1))
                recall[t, a, mi] = rc[-1] if nd else 0
                pr = pr.tolist()
                for i in range(nd - 1, 0, -1):     # precision envelope
                    if pr[i] > pr[i - 1]:
                        pr[i - 1] = pr[i]
                q = np.zeros(R)
                idx = np.searchsorted(rc, REC_THRS, side="left")
                for ri, pi in enumerate(idx):
                    if pi < nd:
                        q[ri] = pr[pi]
                precision[t, :, a, mi] = q
    return precision, recall


def _summarize(precision, recall, max_dets) -> List[float]:
    def stat(ap, iou_thr=None, area="all", max_det=None):
        a = AREA_LBL.index(area)
        m = max_dets.index(max_det)
        s = precision[:, :, a, m] if ap else recall[:, a, m]
        if iou_thr is not None:
            s = s[np.where(np.isclose(IOU_THRS, iou_thr))[0]]
        s = s[s > -1]
        return float(np.mean(s)) if s.size else -1.0
    md = max_dets
    return [stat(1, max_det=md[2]), stat(1, 0.5, max_det=md[2]), stat(1, 0.75, max_det=md[2]),
            stat(1, area="small", max_det=md[2]), stat(1, area="medium", max_det=md[2]), stat(1, area="large", max_det=md[2]),
            stat(0, max_det=md[0]), stat(0, max_det=md[1]), stat(0, max_det=md[2]),
            stat(0, area="small", max_det=md[2]), stat(0, area="medium", max_det=md[2]), stat(0, area="large", max_det=md[2])]


class COCOEvaluator:
    """Class-agnostic (every annotation is category 1, as in the reference's *_cls_agnostic.json ground truth)
    COCO evaluation of box and mask predictions."""

    def __init__(self, gt, tasks: Iterable[str] = ("bbox", "segm"), max_dets_per_image: Optional[Sequence[int]] = None):
        if isinstance(gt, str):
            with open(gt) as f:
                gt = json.load(f)
        self.tasks = tuple(tasks)
        self.max_dets = list(max_dets_per_image) if max_dets_per_image is not None else [1, 10, 100]
        if len(self.max_dets) < 3:
            raise ValueError("COCOeval requires maxDets to have length at least 3")   # coco_evaluation.py:613-616
        self.images = {im["id"]: im for im in gt.get("images", [])}
        self.gts: Dict[int, List[dict]] = {}
        for ann in gt.get("annotations", []):
            self.gts.setdefault(ann["image_id"], []).append(ann)
        self.reset()

    def reset(self):
        self._pred: Dict[int, List[dict]] = {}

    def process(self, image_id, coco_instances: List[dict]):
        """COCO_evaluator/coco_evaluation.py:182-187: the instances of one image, COCO result dicts."""
        self._pred.setdefault(image_id, []).extend(a for a in coco_instances if a is not None)

    # -- per image ----------------------------------------------------------------------------
    def _image_hw(self, image_id, anns):
        im = self.images.get(image_id)
        if im is not None and "height" in im:
            return int(im["height"]), int(im["width"])
        for a in anns:
            seg = a.get("segmentation")
            if isinstance(seg, dict) and "size" in seg:
                return int(seg["size"][0]), int(seg["size"][1])
        raise ValueError(f"image {image_id}: no size in the ground truth and no RLE to take it from")

    def _eval_task(self, task: str, img_ids) -> Dict[str, float]:
        per_image = [[] for _ in AREA_RNG]
        cap = self.max_dets[-1]
        for image_id in img_ids:
            gts = self.gts.get(image_id, [])
            dts = self._pred.get(image_id, [])
            if not gts and not dts:
                continue
            order = np.argsort([-float(d.get("score", 1.0)) for d in dts], kind="mergesort")[:cap]
            dts = [dts[i] for i in order]
            crowd = np.array([int(g.get("iscrowd", 0)) for g in gts], dtype=np.int64)
            ignore = np.array([int(g.get("ignore", 0)) or int(g.get("iscrowd", 0)) for g in gts], dtype=np.int64)
            if task == "bbox":
                ious = bbox_iou(np.array([d["bbox"] for d in dts]).reshape(-1, 4), np.array([g["bbox"] for g in gts]).reshape(-1, 4), crowd)
                d_area = [float(d["bbox"][2] * d["bbox"][3]) if "area" not in d else float(d["area"]) for d in dts]
                g_area = [float(g["area"]) if "area" in g else float(g["bbox"][2] * g["bbox"][3]) for g in gts]
            else:
                h, w = self._image_hw(image_id, dts + gts)
                dm = np.stack([segmentation_to_mask(d["segmentation"], h, w) for d in dts]) if dts else np.zeros((0, h, w), np.uint8)
                gm = np.stack([segmentation_to_mask(g["segmentation"], h, w) for g in gts]) if gts else np.zeros((0, h, w), np.uint8)
                ious = mask_iou(dm, gm, crowd)
                # mask AP uses the MASK area of a detection (coco_evaluation.py:600-606 drops 'bbox' for this)
                d_area = [float(m.sum()) for m in dm]
                g_area = [float(g["area"]) if "area" in g else float(m.sum()) for g, m in zip(gts, gm)]
            scores = np.array([float(d.get("score", 1.0)) for d in dts])
            for a, rng in enumerate(AREA_RNG):
                matched, dt_ig, g_ig = _evaluate_img(ious, g_area, crowd, ignore, scores, d_area, rng, cap)
                per_image[a].append((scores, matched, dt_ig, g_ig))
        precision, recall = _accumulate(per_image, self.max_dets)
        stats = _summarize(precision, recall, self.max_dets)
        return {m: (float(s * 100) if s >= 0 else float("nan")) for m, s in zip(METRICS, stats)}

    def evaluate(self, img_ids=None) -> Dict[str, Dict[str, float]]:
        """COCO_evaluator/coco_evaluation.py:189-220.  Images = those of the ground truth (or `img_ids`)."""
        if img_ids is None:
            img_ids = sorted(set(self.images) | set(self.gts)) or sorted(self._pred)
        return {task: self._eval_task(task, list(img_ids)) for task in self.tasks}


def evaluate_ap(gt_annotation_path: str, pred_annotation_path: str, tasks=("bbox", "segm")) -> dict:
    """COCO_evaluator/main.py:24-70 without the plotting: JSON files in, the ap_score.json dict out."""
    with open(pred_annotation_path) as f:
        preds = json.load(f)
    ev = COCOEvaluator(gt_annotation_path, tasks=tasks)
    by_image: Dict[int, List[dict]] = {}
    for i, ann in enumerate(preds):
        if ann is None:
            continue
        ann.setdefault("id", i)
        if "score" not in ann:
            ann["score"] = ann.get("weight", 1)   # main.py:55-59
        by_image.setdefault(ann["image_id"], []).append(ann)
    for image_id, instances in by_image.items():
        ev.process(image_id, instances)
    results = ev.evaluate()
    results.update(pred_annotation_path=pred_annotation_path, gt_annotation_path=gt_annotation_path,
                   number_of_images=len(by_image), number_of_annotations=len(preds))
    return results
