"""GPU parity tests: the CUDA path (through the C ABI) against the reference's golden vectors
and against the CPU oracle on seeded inputs.  Run on the B200 box: pytest -m gpu.

Tolerances (BASELINE.json north_star): scores and box coordinates within 1e-5 relative in
fp32 — for a coordinate the scale is max(|coordinate|, box side), because a round moves an
edge by delta * side / 128 (object_reasoning.py:185-194); NMS keep-sets, labels, list
membership / order and binary masks bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from unmore_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def od(dev):
    from unmore_b200.object_reasoning import Object_Discovery
    return Object_Discovery(device=dev)


@pytest.fixture(scope="module")
def ops():
    from unmore_b200 import ops as _ops
    return _ops


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def assert_boxes_close(got, ref, what=""):
    got = np.asarray(got, np.float64).reshape(-1, 4)
    ref = np.asarray(ref, np.float64).reshape(-1, 4)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if not len(ref):
        return
    side = np.maximum(ref[:, 2] - ref[:, 0], ref[:, 3] - ref[:, 1])[:, None]
    scale = np.maximum(np.abs(ref), side)
    err = np.abs(got - ref)
    bad = err > RTOL * scale
    assert not bad.any(), f"{what}: {int(bad.sum())} coords off, worst rel {np.max(err / np.maximum(scale, 1e-30)):.3e}"


def assert_rel(got, ref, what="", rtol=RTOL):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = np.abs(got - ref)
    assert (err <= rtol * np.abs(ref)).all(), f"{what}: worst rel {np.max(err / np.maximum(np.abs(ref), 1e-30)):.3e}"


# ------------------------------------------------------------------------------------------
# unit ops against reference goldens
# ------------------------------------------------------------------------------------------
def test_update_bbox_from_tiles(golden_dir, dev, ops):
    g = _load(golden_dir, "units.npz")
    d, mx = ops.update_bbox_from_tiles(torch.tensor(g["a10_tiles"], device=dev))
    assert_rel(d.cpu().numpy(), g["a10_deltas"], "a10 deltas")
    assert np.array_equal(mx.cpu().numpy(), g["a10_tiles"].max(axis=(1, 2)))


def test_crop_resize_bit_exact(golden_dir, dev, ops):
    """a2 directly: the resized 128x128 crops against the reference's own Resize output, bit for bit
    (up- and down-sampling, fractional boxes, image-edge and full-image windows), and through both
    mirrors of get_prediction_with_proposals."""
    from unmore_b200.object_reasoning import Object_Discovery
    from unmore_b200.object_scoring import Object_Scoring
    g = _load(golden_dir, "units.npz")
    fields = synth.make_fields(3).to(dev)[None].contiguous()
    boxes = torch.tensor(g["crop_boxes"], dtype=torch.float64, device=dev)[None].contiguous()
    got = ops.crop_resize(fields, boxes, [0, 1, 2, 3])[0].cpu().numpy()
    assert np.array_equal(got.view(np.int32), g["crop_out"].view(np.int32))
    sdf, cen = Object_Discovery(device=dev).get_prediction_with_proposals(g["crop_boxes"], fields[0])
    assert np.array_equal(sdf.cpu().numpy(), g["crop_out"][:, 0]) and np.array_equal(cen.cpu().numpy(), g["crop_out"][:, 1:3])
    pred = Object_Scoring(device=dev).get_prediction_with_proposals(fields[0], g["crop_boxes"].tolist())
    assert set(pred) == {"pred_boundary_fields", "pred_center_fields", "pred_existence_scores"}
    assert np.array_equal(pred["pred_boundary_fields"].cpu().numpy(), g["crop_out"][:, 0])
    assert_rel(pred["pred_existence_scores"].cpu().numpy(), g["crop_out"][:, 3].mean(axis=(1, 2)), "existence")
    # fp32 boxes and ragged counts
    counts = torch.tensor([3], dtype=torch.int32, device=dev)
    got32 = ops.crop_resize(fields, boxes.float().contiguous(), [0], counts)[0].cpu().numpy()
    b32 = g["crop_boxes"].astype(np.float32).astype(np.float64)
    for k in range(3):
        ref = O.crop_resize(synth.make_fields(3), b32[k])[0].numpy()
        assert np.array_equal(got32[k, 0], ref)
    assert not got32[3:].any()


def test_mask_resize_round_half_even_golden(golden_dir, dev, ops):
    """N2: Resize of int64 masks back to a box (object_scoring.py:206-207) against the reference's
    output, sizes on both sides of ATen's h+w <= 128 kernel switch, dyadic sizes with exact 0.5 ties."""
    g = _load(golden_dir, "units.npz")
    masks = torch.tensor(g["n2_masks"], device=dev)
    for k, (h, w) in enumerate(g["n2_sizes"]):
        got = ops.mask_resize(masks, int(h), int(w)).cpu().numpy()
        assert np.array_equal(got, g[f"n2_out_{k}"]), (h, w)
    # many more sizes against the oracle (same ATen call the reference makes)
    rng = np.random.default_rng(11)
    for _ in range(25):
        h, w = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        got = ops.mask_resize(masks, h, w).cpu().numpy()
        for i in range(masks.shape[0]):
            ref = O.resize_mask_to_box(torch.tensor(g["n2_masks"][i]).long(), h, w).numpy()
            assert np.array_equal(got[i], ref), (h, w, i)


def test_batch_erode_and_anti_center(golden_dir, dev, ops):
    from unmore_b200.utils.misc import batch_erode
    g = _load(golden_dir, "units.npz")
    e = batch_erode(torch.tensor(g["a5_masks"], device=dev).long(), kernel_size=9, num_round=3)
    assert e.dtype == torch.int64 and np.array_equal(e.cpu().numpy().astype(np.uint8), g["a5_out"])
    a = ops.anti_center_map(torch.tensor(g["a6_in"], device=dev))
    assert np.allclose(a.cpu().numpy(), g["a6_out"], rtol=0, atol=1e-14)


def test_box_nms_golden(golden_dir, dev, ops):
    g = _load(golden_dir, "units.npz")
    b = torch.tensor(g["a14_boxes"], device=dev)[None].contiguous()
    s = torch.tensor(g["a14_scores"], device=dev)[None].contiguous()
    keep, kc, kb = ops.box_nms(b, s)
    assert np.array_equal(keep[0, : int(kc[0])].cpu().numpy(), g["a14_keep"])
    assert np.array_equal(kb[0, : int(kc[0])].cpu().numpy(), g["a14_boxes"][g["a14_keep"]])
    keep, kc, _ = ops.box_nms(b, None)
    assert np.array_equal(keep[0, : int(kc[0])].cpu().numpy(), g["a14_keep_allones"])
    # matrix variant: same keep-set
    assert np.array_equal(ops.box_nms_matrix(b[0], s[0]).cpu().numpy(), g["a14_keep"])
    assert np.array_equal(ops.box_nms_matrix(b[0], None).cpu().numpy(), g["a14_keep_allones"])


# ------------------------------------------------------------------------------------------
# stage-by-stage against the reference's run on a synthetic scene (configs[0]: 512 proposals)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b"])
def test_scene_existence_and_center(golden_dir, od, dev, tag):
    g = _load(golden_dir, f"scene_{tag}.npz")
    idx, n_prop = int(g["index"]), int(g["n_prop"])
    fields = synth.make_fields(idx).to(dev)
    props = torch.tensor(synth.make_proposals(idx, n_prop))
    ex = od.existence_checking(fields, props)["existence_scores"]
    assert ex.dtype == torch.float32 and ex.device.type == "cpu"
    assert_rel(ex.numpy(), g["existence_scores"], "existence")
    assert np.array_equal(ex.numpy() >= 0.1, g["existence_scores"] >= 0.1)
    p1 = props[torch.tensor(g["existence_scores"]) >= 0.1]
    cr = od.center_reasoning(fields, p1)
    assert cr["proposals_pass_singularity"].dtype == torch.float64
    assert np.array_equal(cr["proposals_pass_singularity"].cpu().numpy(), g["pass1"])
    assert np.array_equal(cr["splited_new_proposals"].cpu().numpy().reshape(-1, 4), g["split"].reshape(-1, 4))
    if len(g["split"]):
        sp = torch.tensor(g["split"])
        ex2 = od.existence_checking(fields, sp)["existence_scores"].numpy()
        assert_rel(ex2, g["split_existence"], "split existence")
        cr2 = od.center_reasoning(fields, sp[torch.tensor(g["split_existence"]) >= 0.1])
        assert np.array_equal(cr2["proposals_pass_singularity"].cpu().numpy(), g["pass2"])


def test_center_reasoning_plateaus_and_ties(od, dev, ops):
    """Center reasoning where the screening pass cannot separate the maximum: constant center fields (every
    surviving pixel ties -> more exact-pass candidates than the queue holds) and a two-valued field whose seam
    gives long runs of exactly equal maxima (argmax must be the first in raster order).  Against the oracle:
    max values bit-equal, pass / split lists equal."""
    H, W = 480, 640
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    props = torch.tensor(synth.make_proposals(5, 96, H, W))
    args = O.make_args()
    cases = []
    f = torch.zeros((4, H, W)); f[0] = 1.0; f[1] = 0.6; f[2] = -0.3; f[3] = 1.0
    cases.append(("constant", f))
    f = torch.zeros((4, H, W)); f[0] = 1.0; f[3] = 1.0
    f[1] = torch.where(yy < 240, 1.0, -1.0); f[2] = torch.where(xx < 320, 0.5, -0.5)   # vectors converge on the seams
    cases.append(("seams", f))
    f = torch.zeros((4, H, W)); f[0] = 1.0; f[3] = 1.0                                  # zero center field: all scores 0
    cases.append(("zero", f))
    for name, fields in cases:
        ref = O.center_reasoning(fields, props, args, return_debug=True)
        got = od.center_reasoning(fields.to(dev), props)
        assert np.array_equal(got["proposals_pass_singularity"].cpu().numpy(), ref["proposals_pass_singularity"].numpy()), name
        assert np.array_equal(got["splited_new_proposals"].cpu().numpy().reshape(-1, 4),
                              ref["splited_new_proposals"].numpy().reshape(-1, 4)), name
        mv, am, _, _ = ops.center_reasoning(fields.to(dev)[None].contiguous(), props.to(dev)[None].contiguous())[:4]
        assert np.array_equal(mv[0].cpu().numpy(), ref["max_values"].numpy()), name
    assert len(O.center_reasoning(cases[1][1], props, args)["splited_new_proposals"]) > 0   # the seam case does split


def test_center_reasoning_survivor_bands_vs_oracle(od, dev, ops):
    """The center kernel resamples only the rows that can hold survivors of the 25x25 erosion (decided from eight
    lattice rows).  Union masks made of rectangles whose edges sweep across the lattice rows, the 12-pixel erosion
    frame and the 25-pixel run length — one or two bands per 128x128 crop, crops taken 1:1 and scaled — must give the
    reference's maxima (1e-14: fp64 summation order) and exactly the same pass / split lists."""
    H, W = 480, 640
    rng = np.random.default_rng(314)
    args = O.make_args()
    edges = [0, 1, 3, 4, 7, 8, 9, 11, 12, 13, 19, 20, 21, 24, 36, 37]
    sizes = [23, 24, 25, 26, 27, 33, 40, 41, 49, 57, 64, 90, 128]
    for img in range(6):
        f = torch.zeros((4, H, W))
        f[0] = -1.0; f[3] = 1.0
        f[1:3] = torch.tensor(rng.normal(0, 0.12, (2, H, W)).astype(np.float32))
        props = []
        for cy in range(3):
            for cx in range(5):
                oy, ox = 128 * cy + (32 if cy else 0), 128 * cx
                for _ in range(int(rng.integers(1, 3))):          # one or two rectangles in the cell
                    a, c = int(rng.choice(edges)), int(rng.choice(edges))
                    if rng.random() < 0.4:
                        a = 127 - a - int(rng.choice(sizes[:6]))  # hug the bottom edge instead
                    h, w = int(rng.choice(sizes)), int(rng.choice(sizes))
                    a = max(a, 0)
                    f[0, oy + a: min(oy + a + h, oy + 128), ox + c: min(ox + c + w, ox + 128)] = 1.0
                props.append([ox, oy, ox + 128, oy + 128])                      # 1:1 crop
                props.append([ox + 3, oy + 5, ox + 128 - 9, oy + 128 - 2])      # scaled crop of the same content
        props = torch.tensor(props, dtype=torch.float64)
        props[:, [1, 3]] = props[:, [1, 3]].clamp(0, H)
        ref = O.center_reasoning(f, props, args, return_debug=True)
        got = od.center_reasoning(f.to(dev), props)
        assert np.array_equal(got["proposals_pass_singularity"].cpu().numpy(), ref["proposals_pass_singularity"].numpy()), img
        assert np.array_equal(got["splited_new_proposals"].cpu().numpy().reshape(-1, 4),
                              ref["splited_new_proposals"].numpy().reshape(-1, 4)), img
        mv = ops.center_reasoning(f.to(dev)[None].contiguous(), props.to(dev)[None].contiguous())[0]
        got_mv, ref_mv = mv[0].cpu().numpy(), ref["max_values"].numpy()
        # the fp64 correlation is a library convolution in the reference (summation order unspecified): 1e-14, as
        # test_batch_erode_and_anti_center; "no survivor" (exactly 0) must agree exactly
        assert np.array_equal(got_mv == 0, ref_mv == 0) and np.allclose(got_mv, ref_mv, rtol=0, atol=1e-14), img
        n_empty = int((ref["eroded"].reshape(len(props), -1).sum(1) == 0).sum())
        assert 0 < n_empty < len(props)        # both the early exit and the partial pass are exercised


@pytest.mark.parametrize("tag", ["a", "b"])
def test_scene_single_rounds_teacher_forced(golden_dir, od, dev, tag):
    """optimize_one_image_single_round with inputs forced from the reference's own trajectory:
    every recorded round, fp64 inputs on round 0 and fp32 afterwards."""
    g = _load(golden_dir, f"scene_{tag}.npz")
    fields = synth.make_fields(int(g["index"])).to(dev)
    for r in range(int(g["n_trace"])):
        pin = torch.tensor(g[f"r{r}_in"])
        out = od.optimize_one_image_single_round(fields, pin)
        assert np.array_equal(out["labels"].cpu().numpy(), g[f"r{r}_labels"]), f"labels, round {r}"
        assert_boxes_close(out["updated_bboxes"].cpu().numpy(), g[f"r{r}_out"], f"round {r}")


@pytest.mark.parametrize("tag", ["a", "b"])
def test_scene_full_trajectory_and_discovery(golden_dir, od, dev, tag):
    g = _load(golden_dir, f"scene_{tag}.npz")
    idx = int(g["index"])
    fields = synth.make_fields(idx).to(dev)
    br = od.boundary_reasoning(fields, torch.tensor(g["refine_in"]))
    assert np.array_equal(br["labels"].cpu().numpy(), g["final_labels"])
    assert_boxes_close(br["proposals"].cpu().numpy(), g["final_proposals"], "final proposals")
    det = od.discover_image(fields, synth.make_proposals(idx, int(g["n_prop"])))
    assert_boxes_close(det, g["discovered"], "discovered")


@pytest.mark.parametrize("tag", ["a", "b"])
def test_scene_scoring(golden_dir, dev, tag):
    from unmore_b200.object_scoring import Object_Scoring
    g = _load(golden_dir, f"scene_{tag}.npz")
    fields = synth.make_fields(int(g["index"])).to(dev)
    sc = Object_Scoring(device=dev)
    anns = sc.score_image(fields, g["discovered"].astype(np.float64).tolist(), image_id=int(g["index"]))
    assert len(anns) == len(g["score_score"])
    assert np.array_equal(np.array([a["bbox"] for a in anns], np.float32), g["score_bbox"])
    for key in ("score", "existence_score", "center_score", "boundary_score", "area_score"):
        assert_rel([a[key] for a in anns], g["score_" + key], key)
    masks = np.stack([a["segmentation"]["mask"] for a in anns])
    packed = np.packbits(masks.reshape(len(anns), -1), axis=1, bitorder="little")
    assert np.array_equal(packed, g["score_masks_packed"]), "binary masks must be bit-exact"


def test_analyze_cc(golden_dir, dev):
    """a8 + :561-572 — connected components inside center reasoning, against the reference's run."""
    from scipy.ndimage import label
    from unmore_b200.object_reasoning import Object_Discovery, default_args
    g = _load(golden_dir, "scene_cc.npz")
    odc = Object_Discovery(default_args(analyze_cc=True), device=dev)
    for index in g["indices"]:
        index = int(index)
        fields = synth.make_fields(index).to(dev)
        props = torch.tensor(synth.make_proposals(index, int(g[f"i{index}_n_prop"])))
        ex = odc.existence_checking(fields, props)["existence_scores"]
        cr = odc.center_reasoning(fields, props[ex >= 0.1])
        assert np.array_equal(cr["proposals_pass_singularity"].cpu().numpy(), g[f"i{index}_pass1"])
        assert np.array_equal(cr["splited_new_proposals"].cpu().numpy(), g[f"i{index}_split"])
    idx, n_prop = int(g["disc_index"]), int(g["disc_n_prop"])
    det = odc.discover_image(synth.make_fields(idx).to(dev), synth.make_proposals(idx, n_prop))
    assert_boxes_close(det, g["disc"], "discovery with analyze_cc")
    # stand-alone labelling op against scipy on noisy masks (many small components, diagonal links)
    rng = np.random.default_rng(3)
    masks = (rng.random((6, 128, 128)) < 0.0005).astype(np.uint8)   # a few isolated pixels each
    masks[0, 20:60, 30:90] = 1; masks[0, 70:100, 10:40] = 1
    masks[1] = 0                                                     # empty: neither single nor multi
    masks[2] = np.eye(128, dtype=np.uint8); masks[2, 5, 100] = 1     # diagonal chain is ONE 8-connected component
    masks[3, 64, :] = 1; masks[3, :, 64] = 1                         # cross: single
    for k in range(10):                                              # snake: long propagation path
        masks[4, 10 + 6 * k, 5:120] = 1
        masks[4, 10 + 6 * k:16 + 6 * k, 119 if k % 2 == 0 else 5] = 1
    assert all(label(m, np.ones((3, 3), int))[1] <= 16 for m in masks)
    comb, ind = Object_Discovery.separate_connected_components(torch.tensor(masks, device=dev))
    ref, ref_ind = O.separate_connected_components(torch.tensor(masks))
    assert ind == ref_ind and comb["single"] == ref["single"] and comb["multi"] == ref["multi"]
    assert Object_Discovery.enlarge_proposals(ref["multi"], (480, 640), 1.5) == O.enlarge_proposals(ref["multi"], (480, 640), 1.5)


def test_main_loop_and_post_process(golden_dir, dev, od):
    from unmore_b200.object_scoring import Object_Scoring
    from unmore_b200.post_process import select_annotations
    g = _load(golden_dir, "main_loop.npz")
    H, W = int(g["H"]), int(g["W"])
    ids = [int(i) for i in g["ids"]]
    images = [synth.make_fields(i, H, W).to(dev) for i in ids]
    res = od.main_object_discovery(images, ids)
    for i in ids:
        assert_boxes_close(res.get(i, np.zeros((0, 4))), g[f"disc_{i}"], f"image {i}")
    raw = {str(i): g[f"disc_{i}"].astype(np.float64).tolist() for i in ids if len(g[f"disc_{i}"])}
    anns = Object_Scoring(device=dev, raw_annotations=raw).main_object_scoring(images, ids)
    assert np.array_equal(np.array([a["image_id"] for a in anns]), g["ann_image_id"])
    assert np.array_equal(np.array([a["bbox"] for a in anns], np.float32).reshape(-1, 4), g["ann_bbox"])
    for key in ("score", "existence_score", "center_score", "boundary_score", "area_score"):
        assert_rel([a[key] for a in anns], g["ann_" + key], key)
    sel = select_annotations(anns)
    assert np.array_equal(np.array([a["id"] for a in sel]), g["pp_ids"])
    assert np.array_equal(np.array([a["image_id"] for a in sel]), g["pp_image_id"])
    assert_rel([a["score"] for a in sel], g["pp_score"], "post-process score")


# ------------------------------------------------------------------------------------------
# seeded inputs against the oracle (no golden): fractional boxes, edges, ragged batches
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [21, 22])
def test_discovery_and_scoring_vs_oracle(dev, od, seed):
    from unmore_b200.object_scoring import Object_Scoring
    args = O.make_args()
    img = synth.make_fields(seed)
    props = synth.make_proposals(seed, 192)
    dbg = {}
    ref = O.discover_image(img, props, args, debug=dbg)
    stats = {}
    boxes = torch.tensor(props, device=dev)[None].contiguous()
    kb, kc = od.discover_batch(img.to(dev)[None].contiguous(), boxes, stats=stats)
    n_in = int(stats["refine_in"][0])
    assert np.array_equal(stats["refine_in_boxes"][0, :n_in].cpu().numpy(), dbg["refine_in"].numpy())
    assert_boxes_close(kb[0, : int(kc[0])].cpu().numpy(), ref, "discovered vs oracle")
    if len(ref):
        s_ref = O.score_image(img, ref.tolist(), args)
        anns = Object_Scoring(device=dev).score_image(img.to(dev), ref.astype(np.float64).tolist())
        assert np.array_equal(np.array([a["bbox"] for a in anns], np.float32), s_ref["bbox"])
        assert_rel([a["score"] for a in anns], s_ref["score"], "score")
        masks = np.stack([a["segmentation"]["mask"] for a in anns])
        assert np.array_equal(masks, s_ref["masks"])


@pytest.mark.parametrize("hw", [(640, 480), (320, 640), (960, 320), (200, 360)])
def test_kernel_specialisations_on_other_shapes(dev, hw):
    """The hot kernels have compile-time specialisations keyed on the field geometry (center: plane size
    480*640 with consecutive channels; refine: row pitch 640).  Shapes that hit one specialisation but not the
    other — 640x480 (same plane size, other pitch), 320x640 and 960x320 (same pitch or same area with another
    layout) — and one that hits neither must all agree with the oracle."""
    from unmore_b200.object_reasoning import Object_Discovery
    from unmore_b200.object_scoring import Object_Scoring
    H, W = hw
    args = O.make_args()
    img = synth.make_fields(77, H, W)
    props = synth.make_proposals(77, 128, H, W)
    ref = O.discover_image(img, props, args)
    od = Object_Discovery(device=dev)
    od.height, od.width = H, W
    det = od.discover_image(img.to(dev), props)
    assert_boxes_close(det, ref, f"discovered vs oracle at {H}x{W}")
    if len(ref):
        s_ref = O.score_image(img, ref.tolist(), args)
        anns = Object_Scoring(device=dev).score_image(img.to(dev), ref.astype(np.float64).tolist())
        assert_rel([a["score"] for a in anns], s_ref["score"], "score")
        assert np.array_equal(np.stack([a["segmentation"]["mask"] for a in anns]), s_ref["masks"])


def test_rasterise_both_resize_paths_vs_oracle(dev):
    """Boxes chosen so that h+w <= 128 (ATen's small-output kernel) and > 128 (generic kernel),
    dyadic sizes with exact 0.5 ties, image-edge boxes, one-pixel boxes."""
    from unmore_b200.object_scoring import Object_Scoring
    args = O.make_args()
    img = synth.make_fields(31)
    boxes = [[100, 100, 164, 164], [200, 50, 232, 82], [10.2, 20.7, 50.1, 70.9], [300, 200, 556, 456],
             [0, 0, 640, 480], [600.5, 440.5, 640, 480], [320, 240, 321, 241], [50, 60, 114, 124],
             [400, 100, 463.5, 163.5], [120, 300, 376, 364], [5, 5, 21, 117]]
    s_ref = O.score_image(img, boxes, args)
    sc = Object_Scoring(device=dev)
    r = sc.score_batch(img.to(dev)[None].contiguous(), torch.tensor(boxes, dtype=torch.float64, device=dev)[None].contiguous())
    n = int(r["keep_counts"][0])
    keep = r["keep"][0, :n].cpu().numpy()
    assert np.array_equal(keep, s_ref["nms_index"])
    from unmore_b200.object_scoring import unpack_masks
    dense = unpack_masks(r["masks"][0][torch.tensor(keep, device=dev).long()], 640)
    assert np.array_equal(dense, s_ref["masks"])
    assert np.array_equal(r["bbox"][0, :n].cpu().numpy(), s_ref["bbox"])
    assert_rel(r["out"][0, :n, 0].cpu().numpy(), s_ref["score"], "score")


def test_config4_shape_1024_vs_oracle(dev):
    """configs[4] geometry: 1024x1024 fields (4093 anchors exist at that size; 96 of them + 32 random
    boxes here so the CPU oracle finishes in seconds)."""
    from unmore_b200.object_reasoning import Object_Discovery
    from unmore_b200.object_scoring import Object_Scoring
    H = W = 1024
    img = synth.make_fields(41, H, W)
    anchors = synth.anchor_proposals(H, W)
    assert anchors.shape == (4093, 4)
    props = np.concatenate([anchors[::43], synth.random_proposals(41, 32, H, W), anchors[-1:]])
    args = O.make_args()
    ref = O.discover_image(img, props, args)
    odl = Object_Discovery(device=dev)
    det = odl.discover_image(img.to(dev), props)
    assert_boxes_close(det, ref, "1024x1024 discovery")
    if len(ref):
        s_ref = O.score_image(img, ref.tolist(), args)
        anns = Object_Scoring(device=dev).score_image(img.to(dev), ref.astype(np.float64).tolist())
        assert np.array_equal(np.array([a["bbox"] for a in anns], np.float32), s_ref["bbox"])
        assert np.array_equal(np.stack([a["segmentation"]["mask"] for a in anns]), s_ref["masks"])
        assert_rel([a["score"] for a in anns], s_ref["score"], "score")


def test_batch_equals_single_and_ragged(dev, od):
    """Images are independent: a ragged batch must reproduce the per-image results exactly."""
    ids = [3, 4, 5, 6]
    n = [96, 0, 130, 61]
    cap = max(n)
    fields = torch.stack([synth.make_fields(i) for i in ids]).to(dev)
    boxes = torch.zeros((len(ids), cap, 4), dtype=torch.float64)
    for b, (i, k) in enumerate(zip(ids, n)):
        if k:
            boxes[b, :k] = torch.tensor(synth.make_proposals(i, 512)[100:100 + k])
    counts = torch.tensor(n, dtype=torch.int32, device=dev)
    kb, kc = od.discover_batch(fields, boxes.to(dev), counts)
    assert int(kc[1]) == 0
    for b, (i, k) in enumerate(zip(ids, n)):
        single = od.discover_image(fields[b], boxes[b, :k]) if k else np.zeros((0, 4), np.float32)
        assert np.array_equal(kb[b, : int(kc[b])].cpu().numpy(), single)


def test_empty_and_degenerate_inputs(dev, od, ops):
    fields = synth.make_fields(2).to(dev)
    assert od.existence_checking(fields, np.zeros((0, 4)))["existence_scores"].shape == (0,)
    cr = od.center_reasoning(fields, np.zeros((0, 4)))
    assert cr["proposals_pass_singularity"].shape == (0, 4) and cr["splited_new_proposals"].shape == (0, 4)
    assert od.boundary_reasoning(fields, np.zeros((0, 4))) == {"proposals": [], "labels": []}
    # boxes below the area threshold are filtered before round 0 (object_reasoning.py:293-299, strict >)
    small = torch.tensor([[10, 10, 20, 15], [10, 10, 20, 15.0001], [5, 5, 5, 80]], dtype=torch.float64)
    out, lab, rounds = ops.boundary_refine(fields[None].contiguous(), small.to(dev)[None].contiguous())
    assert lab[0, 0].item() == -2 and lab[0, 2].item() == -2
    assert rounds[0].tolist()[0] == 0 and rounds[0].tolist()[2] == 0 and rounds[0].tolist()[1] >= 1
    # NMS of nothing
    keep, kc, _ = ops.box_nms(torch.zeros((2, 0, 4), device=dev), None)
    assert kc.tolist() == [0, 0]


def test_error_reporting(dev, ops):
    from unmore_b200._lib import UnmoreError
    with pytest.raises(UnmoreError):
        ops.existence_scores(torch.zeros((1, 4, 8, 8)), torch.zeros((1, 1, 4), device=dev))  # CPU fields: no fallback
    with pytest.raises(UnmoreError):
        ops.box_nms(torch.zeros((1, 40000, 4), device=dev))  # capacity limit reported, not truncated


# ------------------------------------------------------------------------------------------
# north-star-only ops (parity unpinned by the reference: definitional oracles)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(3, 480, 640), (2, 37, 53), (1, 128, 1024), (5, 64, 130)])
def test_sat_and_box_sums(dev, ops, shape):
    g = torch.Generator().manual_seed(7)
    f = torch.rand(shape, generator=g)
    ref = O.sat_build(f)
    got = ops.sat_build(f.to(dev)).cpu()
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, rtol=1e-12, atol=1e-9)
    # in-place build from a [n_img, C, H, W] stack: planes (i, k) = channel (3, 0)[k] of image i
    if shape[0] >= 2:
        stack = torch.rand((shape[0], 4) + tuple(shape[1:]), generator=g)
        got2 = ops.sat_build_fields(stack.to(dev), [3, 0]).cpu()
        assert torch.allclose(got2[:, 0], O.sat_build(stack[:, 3]), rtol=1e-12, atol=1e-9)
        assert torch.allclose(got2[:, 1], O.sat_build(stack[:, 0]), rtol=1e-12, atol=1e-9)
    H, W = shape[1:]
    boxes = torch.rand((shape[0], 50, 4), generator=g, dtype=torch.float64)
    boxes[..., 0] *= W / 2; boxes[..., 1] *= H / 2
    boxes[..., 2] = boxes[..., 0] + boxes[..., 2] * W / 2; boxes[..., 3] = boxes[..., 1] + boxes[..., 3] * H / 2
    sums, means = ops.box_sums(got.to(dev)[:, None].contiguous(), 0, boxes.to(dev))
    for b in range(shape[0]):
        r = O.box_sums(ref[b], boxes[b])
        assert torch.allclose(sums[b].cpu(), r, rtol=1e-12, atol=1e-9)
        # against the definition, 1e-5 relative (what an fp32 table could not deliver)
        x1, y1, x2, y2 = O.snap_box(boxes[b, 0])
        assert abs(sums[b, 0].item() - f[b, y1:y2, x1:x2].double().sum().item()) <= 1e-9 * max(1.0, r[0].item())


def _random_masks(k, H, W, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    m = np.zeros((k, H, W), dtype=np.uint8)
    for i in range(k):
        cy, cx = rng.uniform(0, H), rng.uniform(0, W)
        ry, rx = rng.uniform(min(8, H / 3), max(8, H / 3)), rng.uniform(min(8, W / 3), max(8, W / 3))
        m[i] = (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2) < 1
    m[k // 2] = m[k // 3]          # exact duplicate
    m[k - 1] = 0                   # empty mask
    return m


@pytest.mark.parametrize("W", [640, 100])
def test_mask_pack_and_mask_nms(dev, ops, W):
    H, k = 120, 150
    dense = _random_masks(k, H, W, seed=5)
    packed = ops.mask_pack(torch.tensor(dense, device=dev))
    ref_packed = np.packbits(np.pad(dense, ((0, 0), (0, 0), (0, (-W) % 32))), axis=2, bitorder="little").view(np.uint32)
    assert np.array_equal(packed.cpu().numpy().view(np.uint32), ref_packed.reshape(k, H, -1))
    areas, tight = ops.mask_stats(packed, W)
    assert np.array_equal(areas.cpu().numpy(), dense.reshape(k, -1).sum(1))
    rng = np.random.default_rng(1)
    scores = rng.random(k).astype(np.float32)
    scores[10:40] = 0.5            # ties: index order decides
    keep = ops.mask_nms(packed, W, torch.tensor(scores, device=dev), 0.5).cpu().numpy()
    assert np.array_equal(keep, O.mask_nms_dense(dense, scores, 0.5))
    keep = ops.mask_nms(packed, W, torch.tensor(scores, device=dev), 0.1).cpu().numpy()
    assert np.array_equal(keep, O.mask_nms_dense(dense, scores, 0.1))


@pytest.mark.parametrize("H,W", [(96, 640), (96, 100), (75, 100), (480, 640), (33, 31), (1, 40), (40, 1), (64, 33), (1000, 1024), (1500, 1600)])
def test_rle_from_packed_masks(dev, ops, H, W):
    """a16 binary_mask_to_rle: GPU run lengths + host string against the restated codec; decode
    round trip back to the dense mask.  Heights that are / are not multiples of the 32-row transpose block,
    single-row and single-column masks, all-ones, checkerboards (a run per pixel) and columns that end / start set."""
    from unmore_b200 import rle
    k = 40 if H * W < 100000 else 12   # (1000, 1024): transposed words only in shared memory; (1500, 1600): bit-serial fallback
    dense = _random_masks(k, H, W, seed=9)
    dense[3] = 1                                   # all ones: zero-length leading run
    dense[4] = 0
    dense[4, ::2, ::2] = 1                         # many short runs
    dense[5] = (np.add.outer(np.arange(H), np.arange(W)) & 1).astype(dense.dtype)   # checkerboard
    dense[6] = 0
    dense[6, -1, :] = 1                            # last row set: every column ends with a one ...
    dense[7] = 0
    dense[7, 0, :] = 1                             # ... / starts with a one
    dense[8] = 0
    dense[8, -1, :] = 1; dense[8, 0, :] = 1        # runs that continue across the column boundary
    dense[9] = 0
    packed = ops.mask_pack(torch.tensor(dense, device=dev))
    enc = rle.encode_packed(packed, W, max_runs=256)   # forces the grow-and-retry path for the dense patterns
    for i in range(k):
        c = O.rle_counts_np(dense[i])
        assert enc[i] == {"size": [H, W], "counts": O.rle_to_string(c)}, i
        assert np.array_equal(rle.decode(enc[i]), dense[i])


def test_scoring_with_rle_segmentation(golden_dir, dev):
    from unmore_b200 import rle
    from unmore_b200.object_scoring import Object_Scoring
    g = _load(golden_dir, "scene_a.npz")
    fields = synth.make_fields(int(g["index"])).to(dev)
    anns = Object_Scoring(device=dev).score_image(fields, g["discovered"].astype(np.float64).tolist(), rle=True)
    masks = np.stack([rle.decode(a["segmentation"]) for a in anns])
    packed = np.packbits(masks.reshape(len(anns), -1), axis=1, bitorder="little")
    assert np.array_equal(packed, g["score_masks_packed"])
    assert rle.scored_annotations_json(anns).startswith("[")


@pytest.mark.parametrize("K", [1000, 16384])
def test_box_nms_sweep_sizes(dev, ops, K):
    """configs[2] sizes: 1k and 16k boxes per image, per-image kernel and matrix kernel, vs the oracle."""
    rng = np.random.default_rng(K)
    c = rng.uniform(0, 600, (K, 2)).astype(np.float32)
    wh = rng.uniform(8, 160, (K, 2)).astype(np.float32)
    boxes = np.concatenate([c, c + wh], axis=1)
    scores = rng.random(K).astype(np.float32)
    scores[::7] = 0.25                       # many exact ties
    ref = O.nms(boxes, scores, 0.5)
    b = torch.tensor(boxes, device=dev)
    s = torch.tensor(scores, device=dev)
    assert np.array_equal(ops.box_nms_matrix(b, s).cpu().numpy(), ref)
    keep, kc, _ = ops.box_nms(b[None].contiguous(), s[None].contiguous())
    assert np.array_equal(keep[0, : int(kc[0])].cpu().numpy(), ref)


# ------------------------------------------------------------------------------------------
# full-size properties (configs[1] shapes: 480x640 fields, 4096 proposals per image)
# ------------------------------------------------------------------------------------------
def test_full_size_properties(dev, od, ops):
    B, N = 6, 4096
    fields = torch.stack([synth.make_fields(100 + i) for i in range(B)]).to(dev)
    props = torch.tensor(np.stack([synth.make_proposals(100 + i, N) for i in range(B)])).to(dev)
    st = {}
    kb, kc = od.discover_batch(fields, props, stats=st)
    kb2, kc2 = od.discover_batch(fields, props)
    assert torch.equal(kc, kc2) and torch.equal(kb, kb2), "run-to-run determinism"
    assert int(kc.min()) > 0
    # 1. every label-1 box is a fixed point of one more round (idempotence of the converged set)
    lab = st["refine_labels"]
    rb = st["refine_boxes"]
    for b in range(B):
        n = int(st["refine_in"][b])
        ok = lab[b, :n] == 1
        conv = rb[b, :n][ok].contiguous()
        out, l2, _ = ops.boundary_refine(fields[b:b + 1], conv[None].contiguous(), n_round=1, apply_small_filter=False,
                                         early_exit=False)
        assert bool((l2[0] == 1).all()) and torch.equal(out[0], conv)
    # 2. NMS invariants on the label-1 set, in index order (all scores equal)
    for b in range(B):
        n = int(st["label1"][b])
        fin, _, _ = ops.compact_boxes(rb[b:b + 1], st["refine_in"][b:b + 1], ops.MODE_LABEL_EQ, lab[b:b + 1], thr=1.0,
                                      out_dtype=torch.float32)
        cand = fin[0, :n].cpu().numpy()
        ref_keep = O.nms(cand, np.ones(n, np.float32), 0.5)
        assert np.array_equal(kb[b, : int(kc[b])].cpu().numpy(), cand[ref_keep])
    # 3. existence scores are means of values in (0,1); lists only shrink where they must
    ex = st["existence_scores"]
    assert bool(((ex >= 0) & (ex <= 1)).all())
    assert bool((st["pass1"] <= N).all()) and bool((st["refine_in"] <= 5 * N).all())


# ------------------------------------------------------------------------------------------
# the BENCHMARK workload under the oracle (configs[1]: anchors + random proposals, 4096 per image,
# batched path) + trajectory-fork accounting
# ------------------------------------------------------------------------------------------
def fork_report(stats, b, dbg):
    """Per-proposal comparison of the refine stage of batch entry ``b`` with the oracle's debug record:
    returns (n_proposals, forks) where forks is a list of (index, kind, detail).  A fork is a proposal whose
    final label differs, or whose final box is off by more than 1e-5 relative."""
    n_in = int(stats["refine_in"][b])
    lab = stats["refine_labels"][b, :n_in].cpu().numpy()
    box = stats["refine_boxes"][b, :n_in].cpu().numpy()
    forks = []
    ref_lab = np.full(n_in, -2.0, np.float32)          # -2: left the list (area filter / zero box)
    ref_box = np.zeros((n_in, 4), np.float32)
    if "refine_index" in dbg and len(dbg["refine_out"]):
        idx = dbg["refine_index"].numpy()
        ref_lab[idx] = dbg["refine_labels"].numpy()
        ref_box[idx] = dbg["refine_out"].numpy()
    for k in range(n_in):
        if lab[k] != ref_lab[k]:
            forks.append((k, "label", (float(lab[k]), float(ref_lab[k]))))
        elif lab[k] >= 0:
            side = max(ref_box[k, 2] - ref_box[k, 0], ref_box[k, 3] - ref_box[k, 1])
            rel = np.abs(box[k] - ref_box[k]) / np.maximum(np.abs(ref_box[k]), max(side, 1e-30))
            if (rel > RTOL).any():
                forks.append((k, "box", float(rel.max())))
    return n_in, forks


def test_bench_workload_vs_oracle(dev):
    """Two full benchmark images (bench.py's generator: the 1225 anchors + 2871 random boxes = 4096 proposals,
    480x640 fields) inside a 4-image batch through ReasoningPipeline.run_chunk — the exact call bench.py times —
    against O.discover_image + O.score_image: lists / labels / masks exact, boxes / scores 1e-5.  Prints the
    fork count of the 50-round trajectories (proposals whose final label or box differs from the oracle's)."""
    from unmore_b200.pipeline import ReasoningPipeline
    from unmore_b200.object_scoring import unpack_masks
    H, W, N = 480, 640, 4096
    ids = [0, 1, 2, 3]
    check = [0, 3]
    fields = torch.stack([synth.make_fields(i, H, W) for i in ids]).to(dev)
    props_np = np.stack([synth.make_proposals(i, N, H, W) for i in ids])
    anchors = synth.anchor_proposals(H, W)
    assert np.array_equal(props_np[0, : len(anchors)], anchors)          # bench.py's layout: anchors first
    assert np.array_equal(props_np[0, len(anchors):], synth.random_proposals(0, N - len(anchors), H, W))
    pipe = ReasoningPipeline(dev)
    st = {}
    r = pipe.run_chunk(fields, torch.tensor(props_np, device=dev), stats=st)
    args = O.make_args()
    total_props = total_forks = 0
    for b in check:
        dbg = {}
        ref = O.discover_image(fields[b].cpu(), props_np[b], args, debug=dbg)
        n_in = int(st["refine_in"][b])
        assert np.array_equal(st["refine_in_boxes"][b, :n_in].cpu().numpy(), dbg["refine_in"].numpy()), "refine inputs"
        n, forks = fork_report(st, b, dbg)
        total_props += n
        total_forks += len(forks)
        print(f"image {ids[b]}: {n} proposals enter the refine loop, {len(forks)} forks {forks[:5]}")
        assert not forks, forks[:10]
        k = int(r["box_counts"][b])
        assert_boxes_close(r["boxes"][b, :k].cpu().numpy(), ref, f"discovered boxes, image {ids[b]}")
        # scoring: feed the ORACLE's boxes to the oracle and the GPU's to the GPU (they agree to 1e-5; the
        # masks depend on the snapped windows only, so they must still be bit-equal)
        s_ref = O.score_image(fields[b].cpu(), ref.tolist(), args)
        kk = int(r["keep_counts"][b])
        assert kk == len(s_ref["score"])
        keep = r["keep"][b, :kk].long()
        assert np.array_equal(keep.cpu().numpy(), s_ref["nms_index"])
        assert np.array_equal(r["bbox"][b, :kk].cpu().numpy(), s_ref["bbox"])
        assert_rel(r["out"][b, :kk, 0].cpu().numpy(), s_ref["score"], "score")
        assert np.array_equal(unpack_masks(r["masks"][b][keep], W), s_ref["masks"])
        sel_ref = O.post_process_filter(s_ref["existence_score"], s_ref["center_score"], s_ref["boundary_score"], args)
        assert np.array_equal(np.nonzero(r["selected"][b, :kk].cpu().numpy())[0], sel_ref)
    print(f"bench-workload parity: {total_props} proposals x <=50 rounds, {total_forks} forks")


@pytest.mark.parametrize("cc", [False, True])
def test_fuzz_parity_reduced(dev, cc):
    """tests/tools/fuzz_parity.py, reduced: 12 (+8 with --analyze_cc) seeded scenes x 160 proposals, full
    discovery + scoring against the oracle; forks are reported and must be zero."""
    from unmore_b200.object_reasoning import Object_Discovery, default_args
    from unmore_b200.object_scoring import Object_Scoring
    od = Object_Discovery(default_args(analyze_cc=cc), device=dev)
    sc = Object_Scoring(device=dev)
    args = O.make_args(analyze_cc=cc)
    n_det = n_forks = 0
    for s in range(3000, 3000 + (8 if cc else 12)):
        img = synth.make_fields(s)
        props = synth.make_proposals(s, 160)
        dbg = {}
        ref = O.discover_image(img, props, args, debug=dbg)
        st = {}
        kb, kc = od.discover_batch(img.to(dev)[None].contiguous(), torch.tensor(props, device=dev)[None].contiguous(), stats=st)
        n_in = int(st["refine_in"][0])
        if "refine_in" in dbg:
            assert np.array_equal(st["refine_in_boxes"][0, :n_in].cpu().numpy(), dbg["refine_in"].numpy()), s
        _, forks = fork_report(st, 0, dbg)
        n_forks += len(forks)
        assert not forks, (s, forks[:5])
        assert_boxes_close(kb[0, : int(kc[0])].cpu().numpy(), ref, f"seed {s}")
        n_det += len(ref)
        if len(ref):
            s_ref = O.score_image(img, ref.tolist(), args)
            anns = sc.score_image(img.to(dev), ref.astype(np.float64).tolist())
            assert len(anns) == len(s_ref["score"]), s
            assert np.array_equal(np.stack([a["segmentation"]["mask"] for a in anns]), s_ref["masks"]), s
            assert_rel([a["score"] for a in anns], s_ref["score"], f"score, seed {s}")
    print(f"fuzz cc={cc}: {n_det} detections, {n_forks} forks")


def test_zero_argument_mains_write_reference_json(golden_dir, dev, tmp_path):
    """main_object_discovery() / main_object_scoring() in the reference's zero-argument form
    (object_reasoning.py:615-665, object_scoring.py:172-272) over a duck-typed dataset: the JSON files they
    write reproduce the golden produced by running the reference's own mains (oracle/gen_golden.py gen_main_loop),
    and the RLE segmentations decode to the masks of the in-memory form."""
    import argparse, json
    from unmore_b200 import rle
    from unmore_b200.object_reasoning import FieldDataset, Object_Discovery
    from unmore_b200.object_scoring import Object_Scoring
    from unmore_b200.post_process import main as post_process_main
    g = _load(golden_dir, "main_loop.npz")
    H, W = int(g["H"]), int(g["W"])
    ids = [int(i) for i in g["ids"]]
    ds = FieldDataset([synth.make_fields(i, H, W) for i in ids], ids)
    od = Object_Discovery(None, dev, test_dataset=ds, result_folder=str(tmp_path))
    od.main_object_discovery()
    with open(tmp_path / "discovery_results.json") as f:
        disc = json.load(f)
    for i in ids:
        assert_boxes_close(np.asarray(disc.get(str(i), np.zeros((0, 4)))), g[f"disc_{i}"], f"image {i}")
    sc = Object_Scoring(argparse.Namespace(raw_annotations_path=str(tmp_path / "discovery_results.json")), dev,
                        test_dataset=ds, result_folder=str(tmp_path))
    assert set(sc.raw_annotations) == set(disc)
    sc.main_object_scoring()
    with open(tmp_path / "object_discovery_with_scores.json") as f:
        anns = json.load(f)
    assert [a["image_id"] for a in anns] == g["ann_image_id"].tolist()
    assert np.array_equal(np.array([a["bbox"] for a in anns], np.float32).reshape(-1, 4), g["ann_bbox"])
    assert_rel([a["score"] for a in anns], g["ann_score"], "score")
    dense = sc.main_object_scoring([f.to(dev) for f in ds.fields], ids)
    for a, d in zip(anns, dense):
        assert a["segmentation"]["size"] == [H, W]
        assert np.array_equal(rle.decode(a["segmentation"]), d["segmentation"]["mask"])
    out = post_process_main(["--pred_annotations_path", str(tmp_path / "object_discovery_with_scores.json")])
    assert [a["id"] for a in out["annotations"]] == g["pp_ids"].tolist()
    assert [a["image_id"] for a in out["annotations"]] == g["pp_image_id"].tolist()
    assert (tmp_path / "selected_training_annotations.json").exists()


def test_reference_named_helpers(golden_dir, dev, od):
    """center_field_to_anti_center_map / unravel_index / binary_mask_to_tight_bbox_coco_style /
    get_prediction_with_proposal_images under their reference names."""
    from unmore_b200.object_scoring import Object_Scoring
    g = _load(golden_dir, "units.npz")
    am = od.center_field_to_anti_center_map(torch.tensor(g["a6_in"], device=dev))
    assert np.allclose(am.cpu().numpy(), g["a6_out"], rtol=0, atol=1e-14)
    assert od.unravel_index(130, (128, 128)) == (1, 2)
    m = np.zeros((48, 70), np.uint8)
    assert Object_Scoring.binary_mask_to_tight_bbox_coco_style(m) == [0.0, 0.0, 0.0, 0.0]
    m[5:9, 33:65] = 1
    assert Object_Scoring.binary_mask_to_tight_bbox_coco_style(m) == [33.0, 5.0, 32.0, 4.0]
    tiles = torch.rand((3, 4, 128, 128), device=dev)
    sdf, cen = od.get_prediction_with_proposal_images(tiles)
    assert torch.equal(sdf, tiles[:, 0]) and torch.equal(cen, tiles[:, 1:3])


def test_analyze_cc_more_components_than_the_device_buffer(dev):
    """a passing proposal whose union mask has more 8-connected components than unmore_cc_cap()
    (speckled masks) must be handled like the reference (any number of components, object_reasoning.py:207-256),
    not abort the batch: center_reasoning, discover_batch and separate_connected_components vs the oracle."""
    from unmore_b200 import _lib
    from unmore_b200.object_reasoning import Object_Discovery, default_args
    cap = _lib.load().unmore_cc_cap()
    H, W = 480, 640
    f = torch.zeros((4, H, W), dtype=torch.float32)
    f[0] = -1.0                      # boundary-distance field: background
    f[3] = 1.0                       # existence: everything passes the check
    n_blobs = 0
    for by in range(6):
        for bx in range(7):
            y, x = 110 + 30 * by, 105 + 28 * bx
            f[0, y:y + 7, x:x + 8] = 1.0
            n_blobs += 1
    assert n_blobs > cap
    props = np.array([[100.0, 100.0, 310.0, 300.0],      # all 42 blobs: > cap components
                      [100.0, 100.0, 200.0, 180.0],      # a few blobs: < cap
                      [0.0, 0.0, 60.0, 60.0]])           # nothing
    args = O.make_args(analyze_cc=True)
    odc = Object_Discovery(default_args(analyze_cc=True), device=dev)
    ref = O.center_reasoning(f, torch.tensor(props), args, return_debug=True)
    from scipy.ndimage import label
    assert label(ref["union"][0].numpy(), np.ones((3, 3), int))[1] > cap
    cr = odc.center_reasoning(f.to(dev), props)
    assert np.array_equal(cr["proposals_pass_singularity"].cpu().numpy(), ref["proposals_pass_singularity"].numpy())
    assert np.array_equal(cr["splited_new_proposals"].cpu().numpy(), ref["splited_new_proposals"].numpy())
    comb, ind = Object_Discovery.separate_connected_components(ref["union"].to(dev))
    rc, ri = O.separate_connected_components(ref["union"])
    assert ind == ri and comb["multi"] == rc["multi"] and comb["single"] == rc["single"]
    det = odc.discover_image(f.to(dev), props)
    assert_boxes_close(det, O.discover_image(f, props, args), "discovery with > cap components")
    # batched: the overflowing image sits next to a normal one
    f2 = torch.stack([f, synth.make_fields(5)]).to(dev)
    p2 = torch.tensor(np.stack([np.concatenate([props, props[:1]]), synth.make_proposals(5, 4)]), device=dev)
    kb, kc = odc.discover_batch(f2, p2)
    assert_boxes_close(kb[0, : int(kc[0])].cpu().numpy(), O.discover_image(f, np.concatenate([props, props[:1]]), args), "batch[0]")
    assert_boxes_close(kb[1, : int(kc[1])].cpu().numpy(), O.discover_image(synth.make_fields(5), synth.make_proposals(5, 4), args), "batch[1]")


def test_antialias_mode_unit_ops(golden_dir, dev, ops):
    """Second resize mode (antialias=True, torchvision >= 0.17 default): the stand-alone a2 / N2 ops against
    goldens produced by running the reference with that default (UNMORE_REF_ANTIALIAS=1 python -m
    oracle.gen_golden units_aa): crops and get_prediction_with_proposals tiles bit-exact, mask resize bit-exact,
    and a10 on those tiles within 1e-5."""
    from unmore_b200.object_reasoning import Object_Discovery, default_args
    g = _load(golden_dir, "units_aa.npz")
    fields = synth.make_fields(3).to(dev)[None].contiguous()
    boxes = torch.tensor(g["crop_boxes"], device=dev)[None].contiguous()
    crops = ops.crop_resize(fields, boxes, [0, 1, 2, 3], antialias=True)[0]
    assert np.array_equal(crops.cpu().numpy(), g["crop_out"])
    plain = ops.crop_resize(fields, boxes, [0, 1, 2, 3])[0]
    assert not torch.equal(plain, crops)                       # the modes differ on crops larger than the tile
    oda = Object_Discovery(default_args(antialias=True), device=dev)
    sdf, cen = oda.get_prediction_with_proposals(g["crop_boxes"], fields[0])
    assert np.array_equal(sdf.cpu().numpy(), g["pred_sdf"]) and np.array_equal(cen.cpu().numpy(), g["pred_center"])
    d, _ = ops.update_bbox_from_tiles(sdf.contiguous())
    assert_rel(d.cpu().numpy(), g["a10_deltas"], "a10 on antialiased tiles")
    masks = torch.tensor(g["n2_masks"], device=dev)
    for k, (h, w) in enumerate(g["n2_sizes"]):
        out = ops.mask_resize(masks, int(h), int(w), antialias=True)
        assert np.array_equal(out.cpu().numpy(), g[f"n2_out_{k}"]), (h, w)
    # ragged counts + proposal slicing of the scratch-bounded wrapper
    cnt = torch.tensor([4], dtype=torch.int32, device=dev)
    part = ops.crop_resize(fields, boxes, [0], counts=cnt, antialias=True)[0]
    assert torch.equal(part[:4, 0], crops[:4, 0]) and float(part[4:].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------
# the whole parity suite of one scene in the SECOND resize mode (antialias=True, tile path) against goldens
# produced by running the reference with torchvision's current default
# (UNMORE_REF_ANTIALIAS=1 python -m oracle.gen_golden scene_aa)
# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def oda(dev):
    from unmore_b200.object_reasoning import Object_Discovery, default_args
    return Object_Discovery(default_args(antialias=True), device=dev)


def test_antialias_scene_existence_and_center(golden_dir, oda, dev):
    g = _load(golden_dir, "scene_aa.npz")
    idx, n_prop = int(g["index"]), int(g["n_prop"])
    fields = synth.make_fields(idx).to(dev)
    props = torch.tensor(synth.make_proposals(idx, n_prop))
    ex = oda.existence_checking(fields, props)["existence_scores"]
    assert_rel(ex.numpy(), g["existence_scores"], "existence (antialias)")
    assert np.array_equal(ex.numpy() >= 0.1, g["existence_scores"] >= 0.1)
    p1 = props[torch.tensor(g["existence_scores"]) >= 0.1]
    cr = oda.center_reasoning(fields, p1)
    assert np.array_equal(cr["proposals_pass_singularity"].cpu().numpy(), g["pass1"])
    assert np.array_equal(cr["splited_new_proposals"].cpu().numpy().reshape(-1, 4), g["split"].reshape(-1, 4))
    if len(g["split"]):
        sp = torch.tensor(g["split"])
        ex2 = oda.existence_checking(fields, sp)["existence_scores"].numpy()
        assert_rel(ex2, g["split_existence"], "split existence (antialias)")
        cr2 = oda.center_reasoning(fields, sp[torch.tensor(g["split_existence"]) >= 0.1])
        assert np.array_equal(cr2["proposals_pass_singularity"].cpu().numpy(), g["pass2"])


def test_antialias_scene_rounds_trajectory_discovery(golden_dir, oda, dev):
    g = _load(golden_dir, "scene_aa.npz")
    idx = int(g["index"])
    fields = synth.make_fields(idx).to(dev)
    for r in range(int(g["n_trace"])):                       # every recorded round, teacher-forced
        out = oda.optimize_one_image_single_round(fields, torch.tensor(g[f"r{r}_in"]))
        assert np.array_equal(out["labels"].cpu().numpy(), g[f"r{r}_labels"]), f"labels, round {r}"
        assert_boxes_close(out["updated_bboxes"].cpu().numpy(), g[f"r{r}_out"], f"round {r} (antialias)")
    br = oda.boundary_reasoning(fields, torch.tensor(g["refine_in"]))
    assert np.array_equal(br["labels"].cpu().numpy(), g["final_labels"])
    assert_boxes_close(br["proposals"].cpu().numpy(), g["final_proposals"], "final proposals (antialias)")
    det = oda.discover_image(fields, synth.make_proposals(idx, int(g["n_prop"])))
    assert_boxes_close(det, g["discovered"], "discovered (antialias)")
    plain = _load(golden_dir, "scene_b.npz")                 # same scene in the primary mode: the modes really differ
    assert plain["discovered"].shape != g["discovered"].shape or not np.allclose(plain["discovered"], g["discovered"], atol=1e-3)


def test_antialias_scene_scoring(golden_dir, dev):
    import argparse
    from unmore_b200.object_scoring import Object_Scoring
    g = _load(golden_dir, "scene_aa.npz")
    fields = synth.make_fields(int(g["index"])).to(dev)
    sc = Object_Scoring(argparse.Namespace(antialias=True), device=dev)
    anns = sc.score_image(fields, g["discovered"].astype(np.float64).tolist(), image_id=int(g["index"]))
    assert len(anns) == len(g["score_score"])
    assert np.array_equal(np.array([a["bbox"] for a in anns], np.float32), g["score_bbox"])
    for key in ("score", "existence_score", "center_score", "boundary_score", "area_score"):
        assert_rel([a[key] for a in anns], g["score_" + key], key + " (antialias)")
    masks = np.stack([a["segmentation"]["mask"] for a in anns])
    packed = np.packbits(masks.reshape(len(anns), -1), axis=1, bitorder="little")
    assert np.array_equal(packed, g["score_masks_packed"]), "binary masks must be bit-exact (antialias)"


def test_antialias_crops_extreme_windows_vs_oracle(dev, ops):
    """Antialiased a2 on windows the goldens do not hold: 1x1, 2x3, a 1-pixel-wide strip, and a 1000x1000 window of a
    1024x1024 field (16-17 taps per axis) — against the pinned numpy restatement (oracle.resize_bilinear_aa_np)."""
    g = torch.Generator().manual_seed(11)
    f = torch.randn((1, 2, 1024, 1024), generator=g)
    boxes = np.array([[10.0, 20.0, 11.0, 21.0], [100.2, 200.7, 101.9, 203.1], [5.0, 5.0, 6.0, 400.0],
                      [12.3, 7.9, 1011.5, 1007.2], [0.0, 0.0, 1024.0, 1024.0], [500.0, 500.0, 628.0, 628.0]])
    got = ops.crop_resize(f.to(dev), torch.tensor(boxes, device=dev)[None].contiguous(), [0, 1], antialias=True)[0].cpu().numpy()
    for k, b in enumerate(boxes):
        x1, y1, x2, y2 = O.snap_box(b)
        for c in range(2):
            ref = O.resize_bilinear_aa_np(f[0, c, y1:y2, x1:x2].numpy(), 128, 128)
            assert np.array_equal(got[k, c], ref), (k, c, float(np.abs(got[k, c] - ref).max()))


def test_pack_detections_rows_edge_cases(dev, ops):
    """unmore_pack_detections: append across batches, empty images, image ids beyond 2^24, capacity overflow flag."""
    from unmore_b200.sharding import merge_rows
    cap = 4
    ids = torch.tensor([2 ** 40 + 7, 3, 2 ** 33], dtype=torch.int64, device=dev)
    bbox = torch.arange(3 * cap * 4, dtype=torch.float32, device=dev).reshape(3, cap, 4).contiguous()
    out5 = torch.arange(3 * cap * 5, dtype=torch.float64, device=dev).reshape(3, cap, 5).contiguous() / 7
    kc = torch.tensor([2, 0, 4], dtype=torch.int32, device=dev)
    rows = ops.detection_rows(16, dev)
    ops.pack_detections(ids, bbox, out5, kc, rows)
    ops.pack_detections(ids[:1], bbox[:1], out5[:1], kc[:1], rows)          # second batch appends
    r = rows.cpu().numpy()
    assert r[0, 0] == 8 and r[0, 1] == 0
    exp_img = [2 ** 40 + 7] * 2 + [2 ** 33] * 4 + [2 ** 40 + 7] * 2
    assert r[1:9, 0].tolist() == [float(v) for v in exp_img]
    assert np.array_equal(r[1:3, 1:5], bbox[0, :2].cpu().numpy()) and np.array_equal(r[3:7, 1:5], bbox[2].cpu().numpy())
    assert np.array_equal(r[1:3, 5], out5[0, :2, 0].cpu().numpy())
    merged, total = merge_rows(rows[None])
    assert int(total) == 8 and merged[:8, 0].cpu().tolist() == sorted(float(v) for v in exp_img)
    small = ops.detection_rows(3, dev)
    ops.pack_detections(ids, bbox, out5, kc, small)
    s = small.cpu().numpy()
    assert s[0, 0] == 3 and s[0, 1] == 1                                     # clamped + overflow flag
    empty = ops.detection_rows(4, dev)
    ops.pack_detections(ids, bbox, out5, torch.zeros(3, dtype=torch.int32, device=dev), empty)
    assert float(empty[0, 0]) == 0


def test_antialias_fuzz_vs_oracle(dev, oda):
    """Seeded scenes in the second resize mode (tile path) against the oracle run with antialias_mode(True) — the
    oracle is pinned in that mode by the reference-run golden (tests/test_oracle_golden.py)."""
    import argparse
    from unmore_b200.object_scoring import Object_Scoring
    sc = Object_Scoring(argparse.Namespace(antialias=True), device=dev)
    args = O.make_args()
    n_det = 0
    for s in range(4000, 4006):
        img = synth.make_fields(s)
        props = synth.make_proposals(s, 120)
        with O.antialias_mode(True):
            ref = O.discover_image(img, props, args)
            s_ref = O.score_image(img, ref.tolist(), args) if len(ref) else None
        det = oda.discover_image(img.to(dev), props)
        assert_boxes_close(det, ref, f"seed {s} (antialias)")
        n_det += len(ref)
        if len(ref):
            anns = sc.score_image(img.to(dev), ref.astype(np.float64).tolist())
            assert len(anns) == len(s_ref["score"]), s
            assert np.array_equal(np.stack([a["segmentation"]["mask"] for a in anns]), s_ref["masks"]), s
            assert_rel([a["score"] for a in anns], s_ref["score"], f"score, seed {s} (antialias)")
    assert n_det > 0
