import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from unmore_b200 import synth, ops
from unmore_b200.object_reasoning import Object_Discovery
dev = torch.device("cuda:0"); H, W, N = 480, 640, 4096; n_img = 32
fields = torch.stack([synth.render_fields(synth.scene_params(i, H, W), H, W, device=dev) for i in range(n_img)])
anch = synth.anchor_proposals(H, W)
props = torch.from_numpy(np.stack([np.concatenate([anch, synth.random_proposals(i, N - len(anch), H, W)]) for i in range(n_img)])).to(dev)
od = Object_Discovery(device=dev); st = {}
od.discover_batch(fields, props, stats=st)
rin, rc = st["refine_in_boxes"], st["refine_in"]
valid = torch.arange(rin.shape[1], device=dev)[None] < rc[:, None]
out = {}
for nr in (50, 49, 48, 47, 46, 44):
    b, l, r = ops.boundary_refine(fields, rin, rc, n_round=nr)
    out[nr] = (b[valid], l[valid], r[valid])
r50 = out[50][2].cpu().numpy(); l50 = out[50][1].cpu().numpy()
print("proposals", len(r50), "total rounds", r50.sum())
hist = np.bincount(r50, minlength=51)
print("rounds histogram (count):", {i: int(c) for i, c in enumerate(hist) if c})
print("share of rounds spent by proposals running all 50:", (r50 == 50).sum() * 50 / r50.sum(), "n=", int((r50 == 50).sum()))
print("labels at end among 50-rounders:", {v: int(((l50 == v) & (r50 == 50)).sum()) for v in (-2, -1, 0, 1)})
full = r50 == 50
for a, b in ((50, 48), (50, 49), (50, 47), (50, 46), (50, 44)):
    same = (out[a][0] == out[b][0]).all(1).cpu().numpy()
    print(f"boxes identical between n_round {a} and {b} among 50-rounders: {int((same & full).sum())} of {int(full.sum())}")
