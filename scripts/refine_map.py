"""Cost map of one boundary-reasoning round (and of the existence / center kernels) over the crop-window size:
ns per proposal(-round) for windows of h x w source pixels, proposals at random positions of synthetic scenes.
usage: python scripts/refine_map.py [kernel=refine|exist|center] [sizes "h1xw1,h2xw2,..."] [n_img] [per_img]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from unmore_b200 import synth, ops

kernel = sys.argv[1] if len(sys.argv) > 1 else "refine"
sizes = sys.argv[2] if len(sys.argv) > 2 else "grid"
n_img = int(sys.argv[3]) if len(sys.argv) > 3 else 32
per = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
dev = torch.device("cuda:0")
H, W = 480, 640
fields = torch.stack([synth.render_fields(synth.scene_params(i, H, W), H, W, device=dev) for i in range(n_img)])
ws = ops.workspace(n_img, dev)
if sizes == "grid":
    hs = [32, 64, 100, 128, 160, 200, 256, 257, 320, 400, 480]
    wds = [32, 64, 100, 128, 129, 200, 256, 320, 400, 512, 640]
    todo = [(h, w) for h in hs for w in wds]
else:
    todo = [tuple(int(v) for v in s.split("x")) for s in sizes.split(",")]
rng = np.random.default_rng(0)

def run(h, w):
    x1 = rng.integers(0, W - w + 1, size=(n_img, per)).astype(np.float64)
    y1 = rng.integers(0, H - h + 1, size=(n_img, per)).astype(np.float64)
    fx = rng.random((n_img, per)) * 0.9 + 0.05       # fractional edges inside the same snapped window
    fy = rng.random((n_img, per)) * 0.9 + 0.05
    b = np.stack([x1 + fx * (w > 1), y1 + fy * (h > 1), x1 + w - fx * (w > 1) * 0.5, y1 + h - fy * (h > 1) * 0.5], axis=2)
    b[..., 0] = np.where(w > 1, b[..., 0] - fx + np.minimum(fx, 0.5), b[..., 0])
    boxes = torch.tensor(np.clip(b, 0, [W, H, W, H]), device=dev).float().contiguous()
    if kernel == "refine":
        fn = lambda: ops.boundary_refine(fields, boxes, None, n_round=1, apply_small_filter=False, early_exit=False, ws=ws, want_rounds=False)
    elif kernel == "exist":
        fn = lambda: ops.existence_scores(fields, boxes, None, ws=ws)
    else:
        fn = lambda: ops.center_reasoning(fields, boxes, None, ws=ws, want_splits=False)
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
    return min(ts) * 1e6 / (n_img * per)

if sizes == "grid":
    print(f"{kernel}: ns per proposal(-round); rows = window height, columns = window width")
    print("h\\w   " + "".join(f"{w:7d}" for w in wds))
    for h in hs:
        print(f"{h:5d} " + "".join(f"{run(h, w):7.1f}" for w in wds))
else:
    for h, w in todo:
        print(f"{kernel} {h}x{w}: {run(h, w):8.1f} ns per proposal(-round)")
