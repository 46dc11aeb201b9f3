"""Repeat every stage on the golden scene many times; report any run that differs from the first."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from unmore_b200 import synth, ops
from unmore_b200.object_reasoning import Object_Discovery
from unmore_b200.object_scoring import Object_Scoring
dev = torch.device("cuda:0")
g = np.load("tests/golden/scene_a.npz")
fields = synth.make_fields(0).to(dev)[None].contiguous()
props = torch.tensor(synth.make_proposals(0, 512), device=dev)[None].contiguous()
p1 = torch.tensor(g["pass1"], device=dev)[None].contiguous()
rin = torch.tensor(g["refine_in"], device=dev)[None].contiguous()
od = Object_Discovery(device=dev); sc = Object_Scoring(device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ref = None
bad = {}
for it in range(n):
    ex = ops.existence_scores(fields, props)
    mv, am, sp, _ = ops.center_reasoning(fields, props)
    rb, lab, _ = ops.boundary_refine(fields, rin)
    kb, kc = od.discover_batch(fields, props)
    r = sc.score_batch(fields, kb[:, :16].contiguous(), kc)
    cur = dict(ex=ex, mv=mv, am=am, sp=sp, rb=rb, lab=lab, kb=kb, kc=kc, out=r["out"], masks=r["masks"], keep=r["keep"])
    cur = {k: v.clone() for k, v in cur.items()}
    if ref is None:
        ref = cur
    else:
        for k in cur:
            if not torch.equal(cur[k], ref[k]):
                bad[k] = bad.get(k, 0) + 1
                if bad[k] <= 2:
                    d = (cur[k] != ref[k]).nonzero()
                    print("iter", it, "stage", k, "differs at", d[:5].tolist(), cur[k][tuple(d[0])].item(), ref[k][tuple(d[0])].item())
print("iterations", n, "mismatching stages:", bad)
