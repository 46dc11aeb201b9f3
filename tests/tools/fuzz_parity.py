"""Seeded fuzz: full discovery + scoring on the GPU against the CPU oracle, many scenes.
Reports every mismatch with its size so forks can be attributed."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import oracle as O
from unmore_b200 import synth
from unmore_b200.object_reasoning import Object_Discovery, default_args
from unmore_b200.object_scoring import Object_Scoring

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_prop = int(sys.argv[2]) if len(sys.argv) > 2 else 160
first = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
cc = len(sys.argv) > 4 and sys.argv[4] == "cc"
dev = torch.device("cuda:0")
od = Object_Discovery(default_args(analyze_cc=cc), device=dev)
sc = Object_Scoring(device=dev)
args = O.make_args(analyze_cc=cc)
stats = dict(scenes=0, det=0, list_mismatch=0, box_bad=0, worst_rel=0.0, mask_bad=0, score_bad=0, refine_in_mismatch=0)
t0 = time.time()
for s in range(first, first + n_seeds):
    img = synth.make_fields(s)
    props = synth.make_proposals(s, n_prop)
    dbg = {}
    try:
        ref = O.discover_image(img, props, args, debug=dbg)
    except Exception as e:
        print("seed", s, "oracle raised", type(e).__name__, e); continue
    st = {}
    kb, kc = od.discover_batch(img.to(dev)[None].contiguous(), torch.tensor(props, device=dev)[None].contiguous(), stats=st)
    det = kb[0, : int(kc[0])].cpu().numpy()
    stats["scenes"] += 1; stats["det"] += len(ref)
    n_in = int(st["refine_in"][0])
    if "refine_in" in dbg and not np.array_equal(st["refine_in_boxes"][0, :n_in].cpu().numpy(), dbg["refine_in"].numpy()):
        stats["refine_in_mismatch"] += 1; print("seed", s, "refine_in differs", n_in, len(dbg["refine_in"]))
    if det.shape != ref.shape:
        stats["list_mismatch"] += 1; print("seed", s, "detection count", det.shape, ref.shape); continue
    if len(ref):
        side = np.maximum(ref[:, 2] - ref[:, 0], ref[:, 3] - ref[:, 1])[:, None]
        rel = np.abs(det - ref) / np.maximum(np.abs(ref), side)
        stats["worst_rel"] = max(stats["worst_rel"], float(rel.max()))
        if (rel > 1e-5).any():
            stats["box_bad"] += int((rel > 1e-5).any(1).sum()); print("seed", s, "boxes off: worst rel", rel.max())
        s_ref = O.score_image(img, ref.tolist(), args)
        anns = sc.score_image(img.to(dev), ref.astype(np.float64).tolist())
        if len(anns) != len(s_ref["score"]):
            stats["score_bad"] += 1; print("seed", s, "annotation count", len(anns), len(s_ref["score"])); continue
        masks = np.stack([a["segmentation"]["mask"] for a in anns])
        if not np.array_equal(masks, s_ref["masks"]):
            stats["mask_bad"] += 1; print("seed", s, "mask bits differ:", int((masks != s_ref["masks"]).sum()))
        got = np.array([a["score"] for a in anns])
        if not np.allclose(got, s_ref["score"], rtol=1e-5, atol=0):
            stats["score_bad"] += 1; print("seed", s, "scores off", np.abs(got - s_ref["score"]).max())
print("fuzz", stats, f"{time.time()-t0:.0f}s")
