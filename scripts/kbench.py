"""Per-kernel micro-benchmark on realistic stage inputs (device time, CUDA events).
usage: python scripts/kbench.py [n_images] [kernel ...]   kernels: exist center refine score sat all"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from unmore_b200 import synth, ops
from unmore_b200.object_reasoning import Object_Discovery

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 64
which = set(sys.argv[2:]) or {"all"}
dev = torch.device("cuda:0")
H, W, N = 480, 640, 4096
fields = torch.stack([synth.render_fields(synth.scene_params(i, H, W), H, W, device=dev) for i in range(n_img)])
anch = synth.anchor_proposals(H, W)
props = torch.from_numpy(np.stack([np.concatenate([anch, synth.random_proposals(i, N - len(anch), H, W)]) for i in range(n_img)])).to(dev)
od = Object_Discovery(device=dev)
st = {}
if which & {"all", "center", "refine", "score"}:
    kb, kc = od.discover_batch(fields, props, stats=st)
torch.cuda.synchronize()

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)

ws = ops.workspace(n_img, dev)
if which & {"all", "exist"}:
    mn, av = timeit(lambda: ops.existence_scores(fields, props, None, ws=ws))
    print(f"exist   : {mn:8.3f} ms min {av:8.3f} avg  {n_img*N/mn/1e3:8.2f} Mprop/s")
    ex = ops.existence_scores(fields, props, None, ws=ws).cpu()
    if os.environ.get("KB_SAVE"):
        torch.save(ex, os.environ["KB_SAVE"] + ".exist")
    if os.environ.get("KB_CMP") and os.path.exists(os.environ["KB_CMP"] + ".exist"):
        print("   vs saved: existence scores bit-equal", bool(torch.equal(torch.load(os.environ["KB_CMP"] + ".exist"), ex)))
if which & {"all", "center"}:
    p1, c1 = st["pass1_boxes"], st["pass1"]
    n = int(c1.sum())
    mn, av = timeit(lambda: ops.center_reasoning(fields, p1, c1, ws=ws))
    print(f"center  : {mn:8.3f} ms min {av:8.3f} avg  {n/mn/1e3:8.2f} Mprop/s  ({n} proposals)")
    co = ops.center_reasoning(fields, p1, c1, ws=ws)
    co = {"maxv": co[0].cpu(), "argmax": co[1].cpu(), "splits": co[2].cpu()}
    if os.environ.get("KB_SAVE"):
        torch.save(co, os.environ["KB_SAVE"] + ".center")
    if os.environ.get("KB_CMP") and os.path.exists(os.environ["KB_CMP"] + ".center"):
        ref = torch.load(os.environ["KB_CMP"] + ".center")
        print("   vs saved: center outputs bit-equal", all(bool(torch.equal(ref[k], co[k])) for k in co))
if which & {"all", "center"}:   # second pass: the split proposals that survive the existence re-check
    p2, c2 = st["pass2_boxes"], st["pass2"]
    n2 = int(c2.sum())
    mn, av = timeit(lambda: ops.center_reasoning(fields, p2, c2, ws=ws, want_splits=False))
    print(f"center2 : {mn:8.3f} ms min {av:8.3f} avg  {n2/mn/1e3:8.2f} Mprop/s  ({n2} split proposals)")
if which & {"all", "refine"}:
    rin, rc = st["refine_in_boxes"], st["refine_in"]
    rounds = int(st["refine_rounds"].sum())
    mn, av = timeit(lambda: ops.boundary_refine(fields, rin, rc, ws=ws, want_rounds=False))
    print(f"refine  : {mn:8.3f} ms min {av:8.3f} avg  {rounds/mn/1e3:8.2f} Mprop-rounds/s  ({int(rc.sum())} proposals, {rounds} rounds)")
    out1 = ops.boundary_refine(fields, rin, rc, ws=ws)
    if os.environ.get("KB_SAVE"):
        torch.save({"boxes": out1[0].cpu(), "labels": out1[1].cpu()}, os.environ["KB_SAVE"])
    if os.environ.get("KB_CMP") and os.path.exists(os.environ["KB_CMP"]):
        ref = torch.load(os.environ["KB_CMP"])
        print("   vs saved: labels equal", bool(torch.equal(ref["labels"], out1[1].cpu())), "max abs box diff",
              float((ref["boxes"] - out1[0].cpu()).abs().max()))
if which & {"all", "score"}:
    cap = 32
    det = kb[:, :cap].contiguous()
    mn, av = timeit(lambda: ops.score_and_rasterise(fields, det, kc))
    print(f"score   : {mn:8.3f} ms min {av:8.3f} avg  ({int(kc.sum())} boxes)")
if which & {"pack"}:
    K = 2048
    dense = (torch.rand((K, H, W), device=dev) < 0.3).to(torch.uint8)
    mn, av = timeit(lambda: ops.mask_pack(dense))
    b = K * H * W + K * H * ((W + 31) // 32) * 4
    print(f"pack    : {mn:8.3f} ms min {av:8.3f} avg  {b/mn/1e6:8.1f} GB/s ({b/1e6:.0f} MB)")
if which & {"all", "sat"}:
    out_buf = torch.empty((n_img, 2, H + 1, W + 1), dtype=torch.float64, device=dev)
    mn, av = timeit(lambda: ops.sat_build_fields(fields, [3, 0], out=out_buf))
    b = n_img * 2 * (H * W * 4 + (H + 1) * (W + 1) * 8)
    print(f"sat     : {mn:8.3f} ms min {av:8.3f} avg  {b/mn/1e6:8.1f} GB/s ({b/1e6:.0f} MB)")
