import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from unmore_b200 import ops
g = np.load("tests/golden/units.npz")
m = torch.tensor(g["a5_masks"], device="cuda")
e = ops.batch_erode(m, 9, 3).cpu().numpy()
ref = g["a5_out"]
print("mismatch", int((e != ref).sum()), "ref ones", int(ref.sum()), "got ones", int(e.sum()))
ys, xs = np.nonzero(ref[0]); print("ref0 bbox", ys.min(), ys.max(), xs.min(), xs.max())
if e[0].any():
    ys, xs = np.nonzero(e[0]); print("got0 bbox", ys.min(), ys.max(), xs.min(), xs.max())
for k, r in [(1, 1), (3, 1), (9, 1), (9, 2), (9, 3)]:
    e = ops.batch_erode(m, k, r).cpu().numpy(); print(k, r, int(e.sum()))
