"""Pins the CPU oracle (oracle/oracle.py) against golden vectors produced by the
unmodified reference (oracle/gen_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from unmore_b200 import synth


def _load(golden_dir, name):
    path = os.path.join(golden_dir, name)
    if not os.path.isfile(path):
        pytest.skip(f"{name} not generated")
    return np.load(path)


def test_anchor_counts():
    # N=1225 at 480x640, 4093 at 1024x1024 [probed on the reference]
    assert synth.anchor_proposals(480, 640).shape == (1225, 4)
    assert synth.anchor_proposals(1024, 1024).shape == (4093, 4)
    a = synth.anchor_proposals(480, 640)
    assert a.dtype == np.float64 and (a[-1] == [0, 0, 640, 480]).all()


def test_crop_resize_bit_exact(golden_dir):
    g = _load(golden_dir, "units.npz")
    img = synth.make_fields(3)
    for box, ref in zip(g["crop_boxes"], g["crop_out"]):
        got = O.crop_resize(img, box).numpy()
        assert np.array_equal(got.view(np.int32), ref.view(np.int32))
        # explicit fp32 arithmetic (the formula the CUDA kernels implement) is bit-identical too
        x1, y1, x2, y2 = O.snap_box(box)
        for c in range(4):
            mine = O.resize_bilinear_np(img[c, y1:y2, x1:x2].numpy(), 128, 128)
            assert np.array_equal(mine.view(np.int32), ref[c].view(np.int32))


def test_update_bbox_with_boundary_fields(golden_dir):
    g = _load(golden_dir, "units.npz")
    d = torch.stack(O.update_bbox_with_boundary_fields(torch.tensor(g["a10_tiles"])), dim=1).numpy()
    assert np.array_equal(d.view(np.int32), g["a10_deltas"].view(np.int32))


def test_post_process_bbox_update(golden_dir):
    g = _load(golden_dir, "units.npz")
    b64, dl = torch.tensor(g["a12_boxes64"]), torch.tensor(g["a12_delta"])
    assert np.array_equal(O.post_process_bbox_update(b64, dl).numpy(), g["a12_out64"])
    assert np.array_equal(O.post_process_bbox_update(b64.float(), dl).numpy(), g["a12_out32"])


def test_batch_erode_and_anti_center(golden_dir):
    g = _load(golden_dir, "units.npz")
    e = O.batch_erode(torch.tensor(g["a5_masks"]).long()).numpy().astype(np.uint8)
    assert np.array_equal(e, g["a5_out"])
    # 3 x (9x9) erosion with zero border == one 25x25 erosion with zero border (used by the CUDA kernel)
    m = g["a5_masks"].astype(bool)
    pad = np.zeros((m.shape[0], 128 + 24, 128 + 24), dtype=bool)
    pad[:, 12:-12, 12:-12] = m
    win = np.lib.stride_tricks.sliding_window_view(pad, (25, 25), axis=(1, 2))
    assert np.array_equal(win.all(axis=(3, 4)).astype(np.uint8), g["a5_out"])
    a = O.center_field_to_anti_center_map(torch.tensor(g["a6_in"])).numpy()
    assert np.allclose(a, g["a6_out"], rtol=0, atol=1e-15)


def test_nms_matches_torchvision(golden_dir):
    g = _load(golden_dir, "units.npz")
    assert np.array_equal(O.nms(g["a14_boxes"], g["a14_scores"], 0.5), g["a14_keep"])
    assert np.array_equal(O.nms(g["a14_boxes"], np.ones(len(g["a14_boxes"]), np.float32), 0.5), g["a14_keep_allones"])


def test_filter_small(golden_dir):
    g = _load(golden_dir, "units.npz")
    p = torch.tensor(g["a9_in"])
    _, lab, _ = O.filter_small_proposal(p, torch.arange(len(p)).float(), O.make_args())
    assert np.array_equal(lab.numpy(), g["a9_keep_index"])


def test_mask_resize_round_half_even(golden_dir):
    g = _load(golden_dir, "units.npz")
    masks = torch.tensor(g["n2_masks"]).long()
    for k, (h, w) in enumerate(g["n2_sizes"]):
        ref = g[f"n2_out_{k}"]
        for i in range(masks.shape[0]):
            got = O.resize_mask_to_box(masks[i], int(h), int(w)).numpy().astype(np.uint8)
            assert np.array_equal(got, ref[i]), (h, w, i)
            # explicit arithmetic: bit = interp > 0.5 (round-half-even sends exactly 0.5 to 0)
            f = O.resize_bilinear_np(masks[i].numpy().astype(np.float32), int(h), int(w))
            assert np.array_equal((f > 0.5).astype(np.uint8), ref[i]), (h, w, i)


def test_sigmoid_threshold_constant():
    t = float(O.SIGMOID_HALF_THRESHOLD)
    x = np.float32(t)
    nxt = np.nextafter(x, np.float32(1))
    assert not bool(torch.sigmoid(torch.tensor(x)) > 0.5)
    assert bool(torch.sigmoid(torch.tensor(nxt)) > 0.5)
    assert nxt.view(np.int32) == 0x33C00001


@pytest.mark.parametrize("tag", ["a", "b"])
def test_scene_stages(golden_dir, tag):
    g = _load(golden_dir, f"scene_{tag}.npz")
    idx, n_prop = int(g["index"]), int(g["n_prop"])
    args = O.make_args(n_round=int(g["n_round"]))
    img = synth.make_fields(idx)
    props = torch.tensor(synth.make_proposals(idx, n_prop))
    ex = O.existence_checking(img, props)["existence_scores"].numpy()
    assert np.array_equal(ex.view(np.int32), g["existence_scores"].view(np.int32))
    p1 = props[torch.tensor(ex) >= args.class_score_thres]
    cr = O.center_reasoning(img, p1, args)
    assert np.array_equal(cr["proposals_pass_singularity"].numpy(), g["pass1"])
    assert np.array_equal(cr["splited_new_proposals"].numpy().reshape(-1, 4), g["split"].reshape(-1, 4))
    if len(g["split"]):
        sp = torch.tensor(g["split"])
        ex2 = O.existence_checking(img, sp)["existence_scores"].numpy()
        assert np.array_equal(ex2.view(np.int32), g["split_existence"].view(np.int32))
        cr2 = O.center_reasoning(img, sp[torch.tensor(ex2) >= args.class_score_thres], args)
        assert np.array_equal(cr2["proposals_pass_singularity"].numpy(), g["pass2"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_scene_single_rounds_teacher_forced(golden_dir, tag):
    g = _load(golden_dir, f"scene_{tag}.npz")
    args = O.make_args()
    img = synth.make_fields(int(g["index"]))
    n = int(g["n_trace"])
    rounds = sorted(set([0, 1, 2, 5, n // 2, n - 1]) & set(range(n)))
    for r in rounds:
        pin = torch.tensor(g[f"r{r}_in"])
        out = O.optimize_one_image_single_round(img, pin, args)
        assert np.array_equal(out["labels"].numpy(), g[f"r{r}_labels"]), r
        assert np.array_equal(out["updated_bboxes"].numpy().view(np.int32), g[f"r{r}_out"].view(np.int32)), r


@pytest.mark.slow
@pytest.mark.parametrize("tag", ["b"])
def test_scene_full_trajectory_and_scoring(golden_dir, tag):
    g = _load(golden_dir, f"scene_{tag}.npz")
    args = O.make_args(n_round=int(g["n_round"]))
    idx = int(g["index"])
    img = synth.make_fields(idx)
    dbg = {}
    det = O.discover_image(img, synth.make_proposals(idx, int(g["n_prop"])), args, debug=dbg)
    assert np.array_equal(dbg["refine_in"].numpy(), g["refine_in"])
    assert np.array_equal(det.view(np.int32), g["discovered"].view(np.int32))
    if len(det):
        sc = O.score_image(img, det.tolist(), args)
        assert np.array_equal(sc["bbox"], g["score_bbox"])
        for key in ("score", "existence_score", "center_score", "boundary_score", "area_score"):
            assert np.array_equal(np.asarray(sc[key], np.float64), g["score_" + key]), key
        packed = np.packbits(sc["masks"].reshape(len(sc["masks"]), -1), axis=1, bitorder="little")
        assert np.array_equal(packed, g["score_masks_packed"])


def test_main_loop_and_post_process(golden_dir):
    g = _load(golden_dir, "main_loop.npz")
    args = O.make_args()
    H, W = int(g["H"]), int(g["W"])
    anns = {k: [] for k in ("image_id", "bbox", "score", "existence_score", "center_score", "boundary_score", "area_score")}
    for i in g["ids"]:
        img = synth.make_fields(int(i), H, W)
        det = O.discover_image(img, synth.anchor_proposals(H, W), args)
        assert np.array_equal(det.view(np.int32), g[f"disc_{int(i)}"].view(np.int32)), i
        if len(det):
            sc = O.score_image(img, det.tolist(), args)
            for k in range(len(sc["score"])):
                anns["image_id"].append(int(i))
                anns["bbox"].append(sc["bbox"][k])
                for key in ("score", "existence_score", "center_score", "boundary_score", "area_score"):
                    anns[key].append(float(sc[key][k]))
    assert np.array_equal(np.array(anns["image_id"]), g["ann_image_id"])
    assert np.array_equal(np.array(anns["bbox"], np.float32).reshape(-1, 4), g["ann_bbox"])
    for key in ("score", "existence_score", "center_score", "boundary_score", "area_score"):
        assert np.array_equal(np.array(anns[key]), g["ann_" + key]), key
    keep = O.post_process_filter(anns["existence_score"], anns["center_score"], anns["boundary_score"], args)
    assert np.array_equal(np.arange(len(keep)), g["pp_ids"])
    assert np.array_equal(np.array(anns["area_score"])[keep], g["pp_score"])
    assert np.array_equal(np.array(anns["image_id"])[keep], g["pp_image_id"])


def test_analyze_cc_center_reasoning(golden_dir):
    """--analyze_cc (object_reasoning.py:561-572): component boxes of passing multi-component masks
    follow the split boxes; pinned on the scenes where the reference itself does not crash."""
    g = _load(golden_dir, "scene_cc.npz")
    args = O.make_args(analyze_cc=True)
    for index in g["indices"]:
        index = int(index)
        img = synth.make_fields(index)
        props = torch.tensor(synth.make_proposals(index, int(g[f"i{index}_n_prop"])))
        ex = O.existence_checking(img, props)["existence_scores"]
        cr = O.center_reasoning(img, props[ex >= args.class_score_thres], args)
        assert np.array_equal(cr["proposals_pass_singularity"].numpy(), g[f"i{index}_pass1"])
        assert np.array_equal(cr["splited_new_proposals"].numpy(), g[f"i{index}_split"])


@pytest.mark.slow
def test_analyze_cc_discovery(golden_dir):
    g = _load(golden_dir, "scene_cc.npz")
    idx, n_prop = int(g["disc_index"]), int(g["disc_n_prop"])
    det = O.discover_image(synth.make_fields(idx), synth.make_proposals(idx, n_prop), O.make_args(analyze_cc=True))
    assert np.array_equal(det.view(np.int32), g["disc"].view(np.int32))


def test_antialias_resize_bit_exact():
    """The second resize mode (torchvision >= 0.17 default antialias=True, SURVEY.md section 8c hazard 1): the
    numpy restatement of ATen's _upsample_bilinear2d_aa — weights with their float/double mix, horizontal pass
    first, and the compiled accumulation order (4-tap unrolled mul+add groups, fused remainder) — is bit-identical
    to torch on down-, up- and mixed sampling, on windows up to 1024 px, and through the int-mask path
    (Resize of an int64 mask = float resize + round half to even, object_scoring.py:206-207)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    for ih, iw, oh, ow in [(480, 640, 128, 128), (300, 200, 128, 128), (129, 257, 128, 128), (37, 500, 128, 128),
                           (64, 640, 128, 128), (1024, 1024, 128, 128), (100, 50, 128, 128), (13, 7, 128, 128),
                           (128, 128, 37, 61), (128, 128, 5, 9), (128, 128, 200, 100), (128, 128, 128, 64), (1, 1, 128, 128)]:
        x = torch.randn((1, 1, ih, iw), generator=g)
        ref = F.interpolate(x, size=(oh, ow), mode="bilinear", align_corners=False, antialias=True)[0, 0].numpy()
        assert np.array_equal(O.resize_bilinear_aa_np(x[0, 0].numpy(), oh, ow), ref), (ih, iw, oh, ow)
    m = (torch.rand((128, 128), generator=g) > 0.5).to(torch.int64)
    for oh, ow in [(37, 61), (200, 100), (64, 64), (5, 300), (128, 128)]:
        f = F.interpolate(m[None, None].float(), size=(oh, ow), mode="bilinear", align_corners=False, antialias=True)[0, 0]
        got = np.round(O.resize_bilinear_aa_np(m.numpy().astype(np.float32), oh, ow))
        assert np.array_equal(got, torch.round(f).numpy()), (oh, ow)


def test_oracle_antialias_mode_reproduces_the_reference_run(golden_dir):
    """The oracle with ``antialias_mode(True)`` (both Resize sites) against the goldens of the reference run with
    torchvision's current default (scene_aa.npz): stage lists, discovery output, scoring boxes and masks."""
    g = _load(golden_dir, "scene_aa.npz")
    idx = int(g["index"])
    img = synth.make_fields(idx)
    props = synth.make_proposals(idx, int(g["n_prop"]))
    args = O.make_args()
    with O.antialias_mode(True):
        dbg = {}
        det = O.discover_image(img, props, args, debug=dbg)
        sc = O.score_image(img, g["discovered"].astype(np.float64).tolist(), args)
    assert O.ANTIALIAS is False
    assert np.array_equal(dbg["existence_scores"].numpy(), g["existence_scores"])
    assert np.array_equal(dbg["pass1"].numpy(), g["pass1"]) and np.array_equal(dbg["refine_in"].numpy(), g["refine_in"])
    assert np.array_equal(det, g["discovered"])
    assert np.array_equal(sc["bbox"], g["score_bbox"])
    assert np.array_equal(np.packbits(sc["masks"].reshape(len(sc["score"]), -1), axis=1, bitorder="little"), g["score_masks_packed"])
    assert np.allclose(sc["score"], g["score_score"], rtol=1e-12, atol=0)
