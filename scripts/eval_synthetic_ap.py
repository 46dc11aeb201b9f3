#!/usr/bin/env python
"""End-task check: class-agnostic COCO AP of the GPU path's detections on
synthetic scenes whose ground truth is known — the ellipses that generated the fields.

    python scripts/eval_synthetic_ap.py [n_images] [n_proposals]

Discovery + scoring run through the public mirrors (Object_Discovery / Object_Scoring) on cuda:0; the
evaluator is unmore_b200/coco_eval.py.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from unmore_b200 import synth
from unmore_b200.coco_eval import COCOEvaluator
from unmore_b200.object_reasoning import Object_Discovery
from unmore_b200.object_scoring import Object_Scoring

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n_prop = int(sys.argv[2]) if len(sys.argv) > 2 else 1225
H, W = 480, 640
dev = torch.device("cuda:0")
od, osc = Object_Discovery(device=dev), Object_Scoring(device=dev)
images, anns = [], []
ys = np.arange(H, dtype=np.float32)[None, :, None]
xs = np.arange(W, dtype=np.float32)[None, None, :]
for i in range(n_img):
    p = synth.scene_params(i, H, W).numpy()
    cy, cx, ry, rx = (p[:, k][:, None, None] for k in range(4))
    rho = np.sqrt(((ys - cy) / ry) ** 2 + ((xs - cx) / rx) ** 2)
    dk = np.minimum(ry, rx) * (1 - rho)
    owner, inside = dk.argmax(0), dk.max(0) > 0
    images.append({"id": i, "height": H, "width": W})
    for k in range(len(p)):
        m = (owner == k) & inside            # visible part of object k
        if m.sum() == 0:
            continue
        yy, xx = np.nonzero(m)
        anns.append({"id": len(anns) + 1, "image_id": i, "category_id": 1, "iscrowd": 0, "area": float(m.sum()),
                     "bbox": [float(xx.min()), float(yy.min()), float(xx.max() - xx.min() + 1), float(yy.max() - yy.min() + 1)],
                     "segmentation": m.astype(np.uint8)})
ev = COCOEvaluator({"images": images, "annotations": anns})
t0 = time.time()
n_det = 0
for i in range(n_img):
    f = synth.make_fields(i, H, W).to(dev)
    det = od.discover_image(f, synth.make_proposals(i, n_prop, H, W))
    a = osc.score_image(f, det.astype(np.float64).tolist(), image_id=i) if len(det) else []
    n_det += len(a)
    ev.process(i, a)
t1 = time.time()
res = ev.evaluate()
print(json.dumps({"images": n_img, "proposals_per_image": n_prop, "gt_objects": len(anns), "detections": n_det,
                  "gpu_s": round(t1 - t0, 2), "eval_s": round(time.time() - t1, 2),
                  "bbox": {k: round(v, 2) for k, v in res["bbox"].items()},
                  "segm": {k: round(v, 2) for k, v in res["segm"].items()}}))
