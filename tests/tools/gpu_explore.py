"""Exploratory GPU-vs-oracle error report (not a test)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import oracle as O
from unmore_b200 import synth, ops
from unmore_b200.object_reasoning import Object_Discovery

dev = torch.device("cuda:0")
G = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden")
od = Object_Discovery(device=dev)
args = O.make_args()

def relerr(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)

u = np.load(os.path.join(G, "units.npz"))
d, mx = ops.update_bbox_from_tiles(torch.tensor(u["a10_tiles"], device=dev))
print("a10 deltas max relerr", relerr(d.cpu().numpy(), u["a10_deltas"]).max())
for tag in "ab":
    g = np.load(os.path.join(G, f"scene_{tag}.npz"))
    idx, n_prop = int(g["index"]), int(g["n_prop"])
    img = synth.make_fields(idx)
    fields = img.to(dev)
    props = torch.tensor(synth.make_proposals(idx, n_prop))
    ex = od.existence_checking(fields, props)["existence_scores"].numpy()
    print(tag, "existence max relerr", relerr(ex, g["existence_scores"]).max(), "keep equal",
          np.array_equal(ex >= 0.1, g["existence_scores"] >= 0.1))
    p1 = props[torch.tensor(g["existence_scores"]) >= 0.1]
    cr = od.center_reasoning(fields, p1)
    print(tag, "center pass equal", np.array_equal(cr["proposals_pass_singularity"].cpu().numpy(), g["pass1"]),
          "split equal", np.array_equal(cr["splited_new_proposals"].cpu().numpy().reshape(-1, 4), g["split"].reshape(-1, 4)))
    n = int(g["n_trace"])
    for r in [0, 1, 2, 5, n - 1]:
        pin = torch.tensor(g[f"r{r}_in"])
        out = od.optimize_one_image_single_round(fields, pin)
        lab = out["labels"].cpu().numpy(); bb = out["updated_bboxes"].cpu().numpy()
        le = np.array_equal(lab, g[f"r{r}_labels"])
        ok = g[f"r{r}_labels"] >= 0
        err = np.abs(bb - g[f"r{r}_out"])
        print(tag, f"round {r}: n={len(pin)} labels equal {le} (mismatch {int((lab != g[f'r{r}_labels']).sum())})",
              "max abs err", err[ok].max() if ok.any() else 0, "max rel err", relerr(bb, g[f"r{r}_out"])[ok][g[f'r{r}_out'][ok] != 0].max() if ok.any() else 0)
    t = time.time()
    br = od.boundary_reasoning(fields, torch.tensor(g["refine_in"]))
    torch.cuda.synchronize()
    fp = br["proposals"].cpu().numpy(); fl = br["labels"].cpu().numpy()
    print(tag, "trajectory: list sizes", fp.shape, g["final_proposals"].shape, "time", time.time() - t)
    if fp.shape == g["final_proposals"].shape:
        print(tag, " labels equal", np.array_equal(fl, g["final_labels"]), "max abs err", np.abs(fp - g["final_proposals"]).max(),
              "rows > 1e-3", int((np.abs(fp - g["final_proposals"]).max(1) > 1e-3).sum()))
    det = od.discover_image(fields, props)
    print(tag, "discovered", det.shape, g["discovered"].shape)
    if det.shape == g["discovered"].shape:
        print(tag, " max abs err", np.abs(det - g["discovered"]).max(), "max rel", relerr(det, g["discovered"])[g["discovered"] != 0].max())
    # timing of the batched pipeline
    B = 8
    fb = torch.stack([synth.make_fields(i) for i in range(B)]).to(dev)
    pb = torch.tensor(np.stack([synth.make_proposals(i, 4096) for i in range(B)])).to(dev)
    for it in range(3):
        torch.cuda.synchronize(); t = time.time()
        st = {}
        kb, kc = od.discover_batch(fb, pb, stats=st)
        torch.cuda.synchronize(); dt = time.time() - t
    print("batch of", B, "x4096:", dt * 1e3, "ms; kept", kc.tolist(), "pass1", st["pass1"].tolist(), "refine_in", st["refine_in"].tolist(),
          "rounds total", int(st["refine_rounds"].sum()))
    break
