"""TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of unMORE's second-stage
multi-object reasoning arithmetic, under the field-stub bridge.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file; the product (unmore_b200/) never does and has no CPU
fallback.

Pinning status: PINNED.  Every function below is checked (tests/test_oracle_golden.py)
against golden vectors produced by executing the *unmodified* reference functions
from /root/reference in the build container (oracle/gen_golden.py, via
oracle/ref_harness.py) and committed under tests/golden/.  Exceptions, which have
no reference counterpart and are therefore "parity unpinned" (SURVEY.md §8a rows A, C):
``sat_build`` / ``box_sums`` (north-star op (a)) and ``mask_nms_dense`` (north-star
op (c)); they are definitional restatements (cumsum; greedy NMS with mask IoU).

The restatement uses the same ATen CPU operators the reference calls (F.interpolate
is exactly what torchvision.transforms.Resize dispatches to for tensors; conv2d in
float64; sigmoid; amax) so that, within one torch build, its outputs are bit-identical
to the reference's.  ``torch.norm(x, dim=1)`` over two channels is restated as
a*a + b*b followed by an IEEE sqrt (``norm2``), verified bit-identical (and ~50x faster
on the strided slice the reference uses, object_reasoning.py:150).  ``resize_bilinear_np`` additionally spells
out the exact fp32 arithmetic (index, lambda, fma order) of ATen's CPU bilinear
kernels as found by experiment in torch 2.11 — that is the formula the CUDA kernels
implement.

Resize mode: antialias=False everywhere (reference pins torchvision 0.14.1,
README.md:25, where tensor resizing is never antialiased).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

CROP = 128  # hard-coded crop side, object_reasoning.py:319,407,505
ANTIALIAS = False  # resize mode of every transforms.Resize call of the path: False = torchvision 0.14.1 (pinned), True = >= 0.17 default


class antialias_mode:
    """``with oracle.antialias_mode(True): ...`` runs the oracle in the second resize mode (both Resize sites: the crops
    and the int-mask resize back to the box)."""

    def __init__(self, on: bool):
        self.on = bool(on)

    def __enter__(self):
        global ANTIALIAS
        self.prev, ANTIALIAS = ANTIALIAS, self.on
        return self

    def __exit__(self, *exc):
        global ANTIALIAS
        ANTIALIAS = self.prev
        return False

DEFAULTS = dict(class_score_thres=0.1, center_score_max_thres=0.009, analyze_cc=False,
                max_sdf_thres=0.5, max_shrink_threshold=16, delta_ratio=0.5, n_round=50,
                proposal_area_thres=50, existence_score_thres=0.5, center_score_thres=0.8,
                boundary_score_thres=0.75, nms_iou=0.5)


def make_args(**over):
    d = dict(DEFAULTS)
    d.update(over)
    return SimpleNamespace(**d)


# --------------------------------------------------------------------------------------
# a2: crop + resize (object_reasoning.py:402-410, 314-321, 500-508; object_scoring.py:126-134)
# --------------------------------------------------------------------------------------
def snap_box(box):
    """floor(x1), floor(y1), ceil(x2), ceil(y2) -> python ints (object_reasoning.py:404)."""
    x1, y1, x2, y2 = (float(v) for v in box)
    return int(math.floor(x1)), int(math.floor(y1)), int(math.ceil(x2)), int(math.ceil(y2))


def crop_resize(image: torch.Tensor, box, size=(CROP, CROP)) -> torch.Tensor:
    """image[:, y1:y2, x1:x2] -> bilinear, align_corners=False, antialias=False.
    Mirrors transforms.Resize on a float tensor (torchvision F_t.resize -> interpolate)."""
    x1, y1, x2, y2 = snap_box(box)
    crop = image[:, y1:y2, x1:x2]
    return F.interpolate(crop.unsqueeze(0), size=list(size), mode="bilinear", align_corners=False,
                         antialias=ANTIALIAS)[0]


def on_edge_flags(box, height, width) -> np.ndarray:
    x1, y1, x2, y2 = snap_box(box)
    return np.array([x1 == 0, y1 == 0, x2 == width, y2 == height])


def crops_for(image: torch.Tensor, proposals) -> torch.Tensor:
    """[N, C, 128, 128] stack of resized crops (the per-box Python loop of the reference)."""
    if len(proposals) == 0:
        return torch.zeros((0, image.shape[0], CROP, CROP), dtype=torch.float32)
    return torch.stack([crop_resize(image, b) for b in proposals], dim=0).to(torch.float32)


# ---- explicit fp32 arithmetic of ATen's CPU bilinear kernels (what the CUDA kernels implement)
def _fma32(a, b, c):
    # exact: fp32*fp32 fits fp64; one rounding to fp64 then to fp32 (double rounding is
    # possible in principle but was never observed against torch in the pinning tests)
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def bilinear_index_weights(in_size: int, out_size: int):
    """ATen area_pixel_compute_source_index + compute_source_index_and_lambda, fp32:
    scale = float(in)/out; src = fma(scale, i+0.5, -0.5) clamped at 0; i0 = int(src);
    i1 = i0 + (i0 < in-1); l1 = clamp(src - i0, 0, 1); l0 = 1 - l1."""
    scale = np.float32(in_size) / np.float32(out_size)
    i = np.arange(out_size, dtype=np.float32)
    src = _fma32(scale, i + np.float32(0.5), np.float32(-0.5))
    src = np.maximum(src, np.float32(0))
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = i0 + (i0 < in_size - 1)
    l1 = np.clip(src - i0.astype(np.float32), np.float32(0), np.float32(1)).astype(np.float32)
    l0 = (np.float32(1) - l1).astype(np.float32)
    return i0, i1, l0, l1


def resize_bilinear_np(v: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Bit-exact numpy restatement of F.interpolate(bilinear, align_corners=False,
    antialias=False) on one fp32 [H, W] plane for torch 2.11 CPU (AVX2/AVX512 builds):

      out_h + out_w > 128 (generic TensorIterator kernel, UpSampleKernel.cpp):
          t0 = fma(v00, w0, v01*w1); t1 = fma(v10, w0, v11*w1); out = fma(t0, h0, t1*h1)
      out_h + out_w <= 128 (channels-last vectorised kernel, scalar tail for C=1):
          out = fma(h1*w1, v11, fma(h1*w0, v10, fma(h0*w0, v00, (h0*w1)*v01)))
    """
    v = np.ascontiguousarray(v, dtype=np.float32)
    ih, iw = v.shape
    y0, y1, h0, h1 = bilinear_index_weights(ih, out_h)
    x0, x1, w0, w1 = bilinear_index_weights(iw, out_w)
    v00 = v[y0][:, x0]
    v01 = v[y0][:, x1]
    v10 = v[y1][:, x0]
    v11 = v[y1][:, x1]
    shape = v00.shape
    W0 = np.broadcast_to(w0[None, :], shape)
    W1 = np.broadcast_to(w1[None, :], shape)
    H0 = np.broadcast_to(h0[:, None], shape)
    H1 = np.broadcast_to(h1[:, None], shape)
    if out_h + out_w > 128:
        t0 = _fma32(v00, W0, v01 * W1)
        t1 = _fma32(v10, W0, v11 * W1)
        return _fma32(t0, H0, t1 * H1)
    p00, p01, p10, p11 = H0 * W0, H0 * W1, H1 * W0, H1 * W1
    return _fma32(p11, v11, _fma32(p10, v10, _fma32(p00, v00, p01 * v01)))


# ---- antialias=True: ATen's _upsample_bilinear2d_aa, the mode torchvision >= 0.17 gives transforms.Resize by
# default (SURVEY.md section 8c hazard 1; the reference pins torchvision 0.14.1 where tensors are never
# antialiased, so antialias=False is the primary mode of this repo and this is the second one).
# PINNED against torch 2.11 CPU in the build container: tests/test_oracle_golden.py::test_antialias_resize_bit_exact.
def aa_index_weights(in_size: int, out_size: int):
    """HelperInterpBase::_compute_indices_min_size_weights_aa (UpSampleKernel.cpp) with its float / double mix
    (opmath = float; literals 0.5 / 1.0 are double): per output index -> (xmin, fp32 weights[xsize]).

        scale = float(in) / float(out);  support = scale >= 1 ? scale : 1;  invscale = scale >= 1 ? float(1.0 / scale) : 1
        center = float(double(scale) * (i + 0.5))
        xmin = max(int64(double(float(center - support)) + 0.5), 0);  xmax = min(int64(double(float(center + support)) + 0.5), in)
        w_j = tri(float((double(float(float(j + xmin) - center)) + 0.5) * double(invscale))),  tri(x) = float(1.0 - |x|) if |x| < 1 else 0
        w_j /= sum_j w_j   (fp32 running sum, fp32 division)
    """
    f32 = np.float32
    scale = f32(in_size) / f32(out_size)
    support = scale if scale >= 1.0 else f32(1.0)
    invscale = f32(1.0 / np.float64(scale)) if scale >= 1.0 else f32(1.0)
    out = []
    for i in range(out_size):
        center = f32(np.float64(scale) * (i + 0.5))
        xmin = max(int(np.float64(f32(center - support)) + 0.5), 0)
        xmax = min(int(np.float64(f32(center + support)) + 0.5), in_size)
        n = xmax - xmin
        w = np.zeros(n, f32)
        total = f32(0)
        for j in range(n):
            x = f32((np.float64(f32(f32(j + xmin) - center)) + 0.5) * np.float64(invscale))
            x = -x if x < 0 else x
            w[j] = f32(1.0 - np.float64(x)) if x < 1.0 else f32(0)
            total = f32(total + w[j])
        if total != 0:
            w = (w / total).astype(f32)
        out.append((xmin, w))
    return out


def _aa_pass(src: np.ndarray, taps) -> np.ndarray:
    """One separable pass along the last axis: src [rows, in] -> [rows, out].  The accumulation order is the
    compiled loop of interpolate_aa_single_dim in this torch build, found by experiment and pinned by the test:
    t = s0 * w0 (rounded); the next 4 * floor((n - 1) / 4) taps are added with SEPARATE multiply and add (the
    unrolled main loop), the remaining (n - 1) mod 4 taps with a fused multiply-add (the scalar remainder loop)."""
    out = np.empty((src.shape[0], len(taps)), np.float32)
    for i, (xmin, w) in enumerate(taps):
        n = len(w)
        t = (src[:, xmin] * w[0]).astype(np.float32)
        main = ((n - 1) // 4) * 4
        for j in range(1, 1 + main):
            t = (t + (src[:, xmin + j] * w[j]).astype(np.float32)).astype(np.float32)
        for j in range(1 + main, n):
            t = _fma32(src[:, xmin + j], w[j], t)
        out[:, i] = t
    return out


def resize_bilinear_aa_np(v: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Bit-exact numpy restatement of F.interpolate(bilinear, align_corners=False, antialias=True) on one fp32
    [H, W] plane (torch 2.11 CPU): the horizontal pass over every source row first, then the vertical pass."""
    v = np.ascontiguousarray(v, dtype=np.float32)
    ih, iw = v.shape
    t = _aa_pass(v, aa_index_weights(iw, out_w))                       # [ih, out_w]
    return _aa_pass(np.ascontiguousarray(t.T), aa_index_weights(ih, out_h)).T.copy()


def crop_resize_aa(image: torch.Tensor, box, size=(CROP, CROP)) -> torch.Tensor:
    """crop_resize with torchvision's current default antialias=True (the real call, for pinning)."""
    x1, y1, x2, y2 = snap_box(box)
    crop = image[:, y1:y2, x1:x2]
    return F.interpolate(crop.unsqueeze(0), size=list(size), mode="bilinear", align_corners=False, antialias=True)[0]


def norm2(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """torch.norm(stack(a, b), dim=1) restated: round(a*a) + round(b*b), then a CORRECTLY
    ROUNDED sqrt (the norm kernel calls std::sqrt).  torch.sqrt must not be used here: its
    vectorised CPU kernel is off by one ulp for ~0.65% of inputs (found while pinning)."""
    s = (a * a + b * b).contiguous()
    return torch.from_numpy(np.sqrt(s.numpy()))


# smallest fp32 x with torch.sigmoid(x) > 0.5 on CPU: 0x33c00001 (found exhaustively);
# i.e. sigmoid(x) > 0.5  <=>  x > 1.5 * 2**-24.  The CUDA kernels use this threshold.
SIGMOID_HALF_THRESHOLD = np.float32(1.5 * 2.0 ** -24)


# --------------------------------------------------------------------------------------
# a3: existence_checking (object_reasoning.py:491-523) with the ExistNet stub
# --------------------------------------------------------------------------------------
def existence_checking(image: torch.Tensor, proposals) -> Dict[str, torch.Tensor]:
    scores = []
    n = len(proposals)
    for b0 in range(0, n, 128):  # num_img_per_batch = 128
        crops = crops_for(image, proposals[b0:b0 + 128])
        scores.append(crops[:, 3].mean((1, 2))[:, None])
    if not scores:
        return {"existence_scores": torch.zeros((0,), dtype=torch.float32)}
    return {"existence_scores": torch.cat(scores, dim=0).squeeze(1)}


# --------------------------------------------------------------------------------------
# a4: get_prediction_with_proposals (object_reasoning.py:301-337) with the FieldNet stub
# --------------------------------------------------------------------------------------
def get_prediction_with_proposals(proposals, image: torch.Tensor):
    sdf, cen = [], []
    for b0 in range(0, len(proposals), 50):  # num_img_per_batch = 50
        crops = crops_for(image, proposals[b0:b0 + 50])
        sdf.append(crops[:, 0])
        cen.append(crops[:, 1:3])
    if not sdf:
        return torch.zeros((0, CROP, CROP)), torch.zeros((0, 2, CROP, CROP))
    return torch.cat(sdf, 0), torch.cat(cen, 0)


# --------------------------------------------------------------------------------------
# a5: batch_erode (utils/misc.py:10-20)
# --------------------------------------------------------------------------------------
def batch_erode(binary_masks: torch.Tensor, kernel_size: int = 9, num_round: int = 3) -> torch.Tensor:
    m = binary_masks.unsqueeze(1)
    kernel = torch.ones(1, 1, kernel_size, kernel_size, dtype=torch.float64)
    for _ in range(num_round):
        conv = F.conv2d(m.double(), kernel, padding=int((kernel_size - 1) / 2))[:, 0]
        m = torch.where(conv >= kernel_size * kernel_size, 1, 0).unsqueeze(1)
    return m.squeeze(1)


# --------------------------------------------------------------------------------------
# a6: center_field_to_anti_center_map (object_reasoning.py:360-377)
# --------------------------------------------------------------------------------------
def anti_center_filter(kernel_size: int = 5) -> torch.Tensor:
    """[1, 2, k, k] float64 filter: f[0,i,j] = (c-i)/n, f[1,i,j] = (c-j)/n, n = L2 norm over
    the channel axis computed in float32 (F.normalize, eps=1e-12) then cast to double."""
    xv, yv = torch.meshgrid([torch.arange(kernel_size), torch.arange(kernel_size)], indexing="ij")
    grid = torch.stack((xv, yv), 2).view((1, kernel_size, kernel_size, 2)).float()
    c = int(kernel_size / 2)
    filt = -grid.permute(0, 3, 1, 2) + torch.tensor([c, c]).unsqueeze(0).unsqueeze(-1).unsqueeze(-1)
    return F.normalize(filt, dim=1).double()


def center_field_to_anti_center_map(vote_maps: torch.Tensor, kernel_size: int = 5) -> torch.Tensor:
    filt = anti_center_filter(kernel_size)
    score = F.conv2d(vote_maps.double(), filt, padding=int((kernel_size - 1) / 2))[:, 0]
    return score / (kernel_size ** 2 - 1)


# --------------------------------------------------------------------------------------
# a8: separate_connected_components (object_reasoning.py:207-256), enlarge_proposals (:259-291)
# --------------------------------------------------------------------------------------
def separate_connected_components(binary_masks: torch.Tensor):
    """8-connected labelling (scipy.ndimage.label, structure = ones(3,3)); a mask with exactly one
    component is "single"; otherwise every component's bbox [x_start, y_start, x_stop, y_stop]
    (slice bounds, stop exclusive) goes to "multi", in label (= raster first-pixel) order."""
    from scipy.ndimage import find_objects, label
    single, multi, indicators = [], [], []
    structure = np.ones((3, 3), dtype=int)
    for b in range(binary_masks.shape[0]):
        labeled, num = label(binary_masks[b].cpu().numpy(), structure)
        boxes = []
        for sl in find_objects(labeled.astype(np.int32)):
            ys, xs = sl
            boxes.append([xs.start, ys.start, xs.stop, ys.stop])
        if num == 1:
            single.append(boxes[0])
            indicators.append(1)
        else:
            indicators.append(0)
            multi.extend(boxes)
    return {"single": single, "multi": multi}, indicators


def enlarge_proposals(proposals, image_shape, ratio):
    """Scale about the centre, int() truncation, clip to (height, width).  NB the reference feeds
    128-crop coordinates here yet clips against the IMAGE size (:567) — reproduced as is."""
    height, width = image_shape
    out = []
    for x1, y1, x2, y2 in proposals:
        cx, cy = (x1 + x2) / 2, (y1 + y2) / 2
        nw, nh = (x2 - x1) * ratio, (y2 - y1) * ratio
        out.append([int(max(cx - nw / 2, 0)), int(max(cy - nh / 2, 0)), int(min(cx + nw / 2, width)),
                    int(min(cy + nh / 2, height))])
    return out


# --------------------------------------------------------------------------------------
# a7: center_reasoning (object_reasoning.py:525-580); analyze_cc adds a8 (:561-572)
# --------------------------------------------------------------------------------------
def center_reasoning(image: torch.Tensor, proposals: torch.Tensor, args, return_debug: bool = False):
    sdf_maps, center_fields = get_prediction_with_proposals(proposals, image)
    sdf_binary = torch.where(torch.sigmoid(sdf_maps) > 0.5, 1, 0)
    cnorm = norm2(center_fields[:, 0], center_fields[:, 1])
    cen_binary = torch.where(cnorm > 0.5, 1, 0)
    union = torch.where((cen_binary + sdf_binary) > 0, 1, 0)
    eroded = batch_erode(union, kernel_size=9, num_round=3)
    score = center_field_to_anti_center_map(center_fields, kernel_size=5)
    fg = score * eroded
    fg[:, 0:10, :] = 0
    fg[:, -10:, :] = 0
    fg[:, :, 0:10] = 0
    fg[:, :, -10:] = 0
    if fg.shape[0]:
        maxv = torch.amax(fg, dim=(1, 2))
    else:
        maxv = torch.zeros((0,), dtype=torch.float64)
    passed = maxv <= args.center_score_max_thres
    out_pass = proposals[passed]
    fail = proposals[~passed]
    splits = []
    argmaxes = []
    for bi, box in zip(torch.nonzero(~passed).flatten().tolist(), fail):
        x1, y1, x2, y2 = (float(v) for v in box)
        flat = int(fg[bi].argmax())
        yc, xc = flat // CROP, flat % CROP
        argmaxes.append((yc, xc))
        # x_center / 128 is an int64 tensor true-divided -> float32 (exact for k/128)
        yr = float(np.float32(yc) / np.float32(CROP))
        xr = float(np.float32(xc) / np.float32(CROP))
        splits.append([x1, y1, x1 + (x2 - x1) * xr, y2])
        splits.append([x1 + (x2 - x1) * xr, y1, x2, y2])
        splits.append([x1, y1, x2, y1 + (y2 - y1) * yr])
        splits.append([x1, y1 + (y2 - y1) * yr, x2, y2])
    out_split = torch.tensor(splits, dtype=torch.float64).reshape(-1, 4)
    if getattr(args, "analyze_cc", False):
        # (:561-572) components of the un-eroded union masks of the PASSING proposals; the bboxes of
        # multi-component masks, enlarged x1.5, are appended (float32 values) to the split list.
        # The reference crashes here when nothing failed singularity; defined as "append to empty".
        H, W = image.shape[-2], image.shape[-1]
        cc, _ = separate_connected_components(union[passed])
        multi = enlarge_proposals(cc["multi"], (H, W), ratio=1.5)
        if len(multi):
            extra = torch.tensor(multi, dtype=torch.float32).to(torch.float64).reshape(-1, 4)
            out_split = torch.cat((out_split, extra), dim=0)
    out = {"proposals_pass_singularity": out_pass, "splited_new_proposals": out_split}
    if return_debug:
        out.update(max_values=maxv, argmax=argmaxes, union=union, eroded=eroded, score=score)
    return out


# --------------------------------------------------------------------------------------
# a9: filter_small_proposal (object_reasoning.py:293-299)
# --------------------------------------------------------------------------------------
def filter_small_proposal(proposals: torch.Tensor, labels: torch.Tensor, args):
    area = (proposals[:, 2] - proposals[:, 0]) * (proposals[:, 3] - proposals[:, 1])
    keep = area > args.proposal_area_thres
    return proposals[keep], labels[keep], keep


# --------------------------------------------------------------------------------------
# a10: update_bbox_with_boundary_fields (object_reasoning.py:140-174)
# --------------------------------------------------------------------------------------
def update_bbox_with_boundary_fields(sdf_maps: torch.Tensor):
    dy = sdf_maps[:, 1:, :] - sdf_maps[:, :-1, :]        # image_gradients, rows 0..H-2
    dx = sdf_maps[:, :, 1:] - sdf_maps[:, :, :-1]        # cols 0..W-2
    dy = dy[:, :, :-1]                                   # [B, H-1, W-1]
    dx = dx[:, :-1, :]
    s = sdf_maps[:, :-1, :-1]
    gnorm = norm2(dy, dx)                                # == torch.norm(cat(dy,dx), dim=1), bitwise
    fg = torch.sigmoid(s)
    bg = 1 - fg
    avg_fg = (fg * gnorm).sum(-1).sum(-1) / (fg.sum(-1).sum(-1) + 1e-8)
    avg_bg = (bg * gnorm).sum(-1).sum(-1) / (bg.sum(-1).sum(-1) + 1e-8)
    step_fg = 1 / (avg_fg + 1e-10)
    step_bg = 1 / (avg_bg + 1e-10)
    step = step_fg[:, None, None] * fg + step_bg[:, None, None] * bg
    movement = step * s
    d_x1 = torch.amax(movement[:, :, 0], dim=1) * (-1)
    d_y1 = torch.amax(movement[:, 0, :], dim=1) * (-1)
    d_x2 = torch.amax(movement[:, :, -1], dim=1)
    d_y2 = torch.amax(movement[:, -1, :], dim=1)
    return d_x1, d_y1, d_x2, d_y2


# --------------------------------------------------------------------------------------
# a12: post_process_bbox_update (object_reasoning.py:177-196)
# --------------------------------------------------------------------------------------
def post_process_bbox_update(original: torch.Tensor, delta: torch.Tensor, sx=CROP, sy=CROP):
    xr = (original[:, 2] - original[:, 0]) / sx
    yr = (original[:, 3] - original[:, 1]) / sy
    out = original.clone()
    out[:, 0] = original[:, 0] + delta[:, 0] * xr
    out[:, 1] = original[:, 1] + delta[:, 1] * yr
    out[:, 2] = original[:, 2] + delta[:, 2] * xr
    out[:, 3] = original[:, 3] + delta[:, 3] * yr
    return out


# --------------------------------------------------------------------------------------
# a11: optimize_one_image_single_round (object_reasoning.py:379-487)
# --------------------------------------------------------------------------------------
def optimize_one_image_single_round(image: torch.Tensor, proposals: torch.Tensor, args):
    n = len(proposals)
    _, H, W = image.shape
    out_bboxes = torch.zeros((n, 4), dtype=torch.float32)
    labels = torch.zeros((n,), dtype=torch.float32)
    if n == 0:
        return {"updated_bboxes": out_bboxes, "labels": labels}
    sdf = torch.cat([crops_for(image, proposals[b0:b0 + 50])[:, 0] for b0 in range(0, n, 50)], dim=0)
    edge = torch.tensor(np.stack([on_edge_flags(b, H, W) for b in proposals]), dtype=torch.float32)
    max_sdf = torch.amax(sdf, dim=(1, 2)).to(torch.float32)
    alive = max_sdf > args.max_sdf_thres
    sdf, edge, props = sdf[alive], edge[alive], proposals[alive]
    labels[:] = torch.where(alive, 0, -1).to(torch.float32)
    if int(alive.sum()) == 0:
        return {"updated_bboxes": out_bboxes, "labels": labels}
    d_x1, d_y1, d_x2, d_y2 = update_bbox_with_boundary_fields(sdf)
    signed = torch.stack([-d_x1, -d_y1, d_x2, d_y2], dim=1)
    signed = torch.where((signed > 0) & (edge == 1), 0, 1).to(torch.float32) * signed
    max_exp = torch.amax(signed, dim=1)
    max_shr = torch.amin(signed, dim=1)
    sub = torch.zeros((len(props),), dtype=torch.float32)
    sub[(max_exp <= 0) & (max_shr >= -args.max_shrink_threshold)] = 1
    labels[labels == 0] = sub
    d_x1 = d_x1 - torch.abs(d_x1) * args.delta_ratio
    d_y1 = d_y1 - torch.abs(d_y1) * args.delta_ratio
    d_x2 = d_x2 + torch.abs(d_x2) * args.delta_ratio
    d_y2 = d_y2 + torch.abs(d_y2) * args.delta_ratio
    delta = torch.stack([d_x1, d_y1, d_x2, d_y2], dim=1)
    delta[sub == 1] = 0
    upd = post_process_bbox_update(props, delta)
    upd[:, 0][upd[:, 0] < 0] = 0
    upd[:, 1][upd[:, 1] < 0] = 0
    upd[:, 2][upd[:, 2] > W] = W
    upd[:, 3][upd[:, 3] > H] = H
    out_bboxes[labels >= 0] = upd.to(torch.float32)
    return {"updated_bboxes": out_bboxes, "labels": labels}


# --------------------------------------------------------------------------------------
# a13: boundary_reasoning (object_reasoning.py:582-612)
# --------------------------------------------------------------------------------------
def boundary_reasoning(image: torch.Tensor, proposals: torch.Tensor, args, trace: Optional[list] = None):
    """``trace``: if a list, receives per round (orig_index, in_boxes, out_boxes, labels) so
    tests can teacher-force single rounds."""
    labels = torch.zeros((len(proposals),), dtype=torch.float32)
    cur = proposals
    idx = torch.arange(len(proposals))
    for _ in range(args.n_round):
        cur, labels, keep = filter_small_proposal(cur, labels, args)
        idx = idx[keep]
        if len(cur) == 0:
            return {"proposals": [], "labels": [], "index": idx}
        out = optimize_one_image_single_round(image, cur, args)
        if trace is not None:
            trace.append((idx.clone(), cur.clone(), out["updated_bboxes"].clone(), out["labels"].clone()))
        cur, labels = out["updated_bboxes"], out["labels"]
    return {"proposals": cur, "labels": labels, "index": idx}


# --------------------------------------------------------------------------------------
# a14: torchvision.ops.nms CPU semantics (object_reasoning.py:661, object_scoring.py:238)
# --------------------------------------------------------------------------------------
def nms(boxes, scores, iou_threshold: float = 0.5) -> np.ndarray:
    """Greedy box NMS as torchvision's CPU kernel computes it, all in fp32:
    order = stable descending sort of scores; area = (x2-x1)*(y2-y1);
    inter = max(0, xx2-xx1) * max(0, yy2-yy1); suppress j if inter/(a_i+a_j-inter) > thr.
    Returns kept indices (int64) in descending-score order."""
    b = np.asarray(boxes, dtype=np.float32).reshape(-1, 4)
    s = np.asarray(scores, dtype=np.float32).reshape(-1)
    n = b.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = np.argsort(-s, kind="stable")
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    area = (x2 - x1) * (y2 - y1)
    thr = np.float32(iou_threshold)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (area[i] + area[rest] - inter)
        suppressed[rest[ovr > thr]] = True
    return np.asarray(keep, dtype=np.int64)


# --------------------------------------------------------------------------------------
# main_object_discovery body for one image (object_reasoning.py:623-662)
# --------------------------------------------------------------------------------------
def discover_image(image: torch.Tensor, proposals, args, debug: Optional[dict] = None) -> np.ndarray:
    """Returns the final ``[K, 4]`` float32 boxes of one image (``results_dict[image_id]``);
    an empty array where the reference ``continue``s."""
    empty = np.zeros((0, 4), dtype=np.float32)
    props = torch.as_tensor(np.asarray(proposals), dtype=torch.float64)
    ex = existence_checking(image, props)["existence_scores"]
    props = props[ex >= args.class_score_thres]
    if debug is not None:
        debug["existence_scores"] = ex
    if len(props) == 0:
        return empty
    cr = center_reasoning(image, props, args)
    p_pass = cr["proposals_pass_singularity"]
    split = cr["splited_new_proposals"]
    if debug is not None:
        debug["pass1"], debug["split"] = p_pass, split
    # the reference crashes on an empty split list (SURVEY.md §7 hard part 9); defined as "no splits"
    if len(split) > 0:
        ex2 = existence_checking(image, split)["existence_scores"]
        split = split[ex2 >= args.class_score_thres]
    if len(split) > 0:
        cr2 = center_reasoning(image, split, args)
        props = torch.cat((p_pass, cr2["proposals_pass_singularity"]), dim=0)
    else:
        props = p_pass
    if debug is not None:
        debug["refine_in"] = props
    if len(props) == 0:
        return empty
    br = boundary_reasoning(image, props, args)
    if len(br["proposals"]) == 0:
        return empty
    final = br["proposals"][br["labels"] == 1]
    if debug is not None:
        debug["refine_out"], debug["refine_labels"], debug["refine_index"] = br["proposals"], br["labels"], br["index"]
    if len(final) == 0:
        return empty
    keep = nms(final.numpy(), np.ones(len(final), dtype=np.float32), args.nms_iou)
    return final.numpy()[keep].astype(np.float32)


# --------------------------------------------------------------------------------------
# a15/a16: main_object_scoring body for one image (object_scoring.py:180-268)
# --------------------------------------------------------------------------------------
def resize_mask_to_box(mask128: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """int64 {0,1} [128,128] -> Resize((h,w), BILINEAR) as torchvision does for integer
    tensors: cast to float32, interpolate, torch.round (half to even), cast back."""
    f = F.interpolate(mask128.to(torch.float32)[None, None], size=[out_h, out_w], mode="bilinear",
                      align_corners=False, antialias=ANTIALIAS)[0, 0]
    return torch.round(f).to(torch.int64)


def tight_bbox_xywh(mask: np.ndarray) -> List[float]:
    """pycocotools rleToBbox semantics: [xmin, ymin, xmax-xmin+1, ymax-ymin+1]; zeros if empty."""
    ys, xs = np.nonzero(mask)
    if ys.size == 0:
        return [0.0, 0.0, 0.0, 0.0]
    return [float(xs.min()), float(ys.min()), float(xs.max() - xs.min() + 1), float(ys.max() - ys.min() + 1)]


def score_image(image: torch.Tensor, raw_proposals, args) -> Dict[str, np.ndarray]:
    """Returns arrays over the kept detections (NMS order): ``bbox`` xywh fp32, ``score``,
    ``existence_score``, ``center_score``, ``boundary_score``, ``area_score``, ``masks``
    (uint8 [K,H,W]), plus ``nms_index`` into the raw proposals."""
    _, H, W = image.shape
    raw = [list(map(float, b)) for b in raw_proposals]
    n = len(raw)
    if n == 0:
        z = np.zeros((0,), dtype=np.float32)
        return dict(bbox=np.zeros((0, 4), np.float32), score=z, existence_score=z, center_score=z,
                    boundary_score=z, area_score=z, masks=np.zeros((0, H, W), np.uint8),
                    nms_index=np.zeros((0,), np.int64))
    crops = torch.cat([crops_for(image, raw[b0:b0 + 50]) for b0 in range(0, n, 50)], dim=0)
    sdf, cen, ex = crops[:, 0], crops[:, 1:3], crops[:, 3].mean((1, 2))
    cnorm = norm2(cen[:, 0], cen[:, 1])
    center_score = torch.amax(cnorm, dim=(1, 2))
    boundary_score = torch.amax(sdf, dim=(1, 2)).to(torch.float32)
    cmask = torch.where(cnorm > 0.5, 1, 0)
    bmask = torch.where(torch.sigmoid(sdf) > 0.5, 1, 0)
    union = np.zeros((n, H, W), dtype=np.uint8)
    tight = []
    for i, box in enumerate(raw):
        x1, y1, x2, y2 = snap_box(box)
        rc = resize_mask_to_box(cmask[i], y2 - y1, x2 - x1)
        rb = resize_mask_to_box(bmask[i], y2 - y1, x2 - x1)
        union[i, y1:y2, x1:x2] = ((rc + rb) > 0).numpy().astype(np.uint8)
        t = tight_bbox_xywh(union[i])
        tight.append([t[0], t[1], t[0] + t[2], t[1] + t[3]])
    tight = np.asarray(tight, dtype=np.float32)
    keep = nms(tight, boundary_score.numpy(), args.nms_iou)
    fb = tight[keep]
    fm = union[keep]
    areas = fm.reshape(len(keep), -1).sum(1).astype(np.int64)
    max_area = areas.max()
    mask_scores = areas / max_area
    area_score = np.power(mask_scores, 0.25)
    ex_k = ex.numpy()[keep]
    cs_k = center_score.numpy()[keep]
    bs_k = boundary_score.numpy()[keep]
    score = ex_k * cs_k * bs_k * area_score
    bbox = np.stack([fb[:, 0], fb[:, 1], fb[:, 2] - fb[:, 0], fb[:, 3] - fb[:, 1]], axis=1)
    return dict(bbox=bbox.astype(np.float32), score=score, existence_score=ex_k, center_score=cs_k,
                boundary_score=bs_k, area_score=area_score, masks=fm, nms_index=keep)


# --------------------------------------------------------------------------------------
# a16: binary_mask_to_rle (object_scoring.py:167-170) — pycocotools 2.0.7 is absent: restated from
# the published maskApi.c (rleEncode / rleToString).  PARITY UNPINNED: no pycocotools, no RLE
# fixture in the reference; anchored on hand-derived known answers and the decode round trip.
# --------------------------------------------------------------------------------------
def rle_counts(mask: np.ndarray) -> List[int]:
    """rleEncode: column-major run lengths, counts[0] = leading zeros (possibly 0)."""
    m = np.asarray(mask).astype(np.uint8).T.reshape(-1)   # column-major scan
    cnts, p, c = [], 0, 0
    for v in m:
        if v != p:
            cnts.append(c)
            c, p = 0, v
        c += 1
    cnts.append(c)
    return cnts


def rle_counts_np(mask: np.ndarray) -> List[int]:
    """Vectorised form of ``rle_counts`` for large masks."""
    m = np.asarray(mask).astype(np.uint8).T.reshape(-1)
    change = np.nonzero(np.diff(np.concatenate([[0], m])))[0]
    edges = np.concatenate([[0], change, [m.size]])
    return np.diff(edges).astype(np.int64).tolist()


def rle_to_string(cnts) -> str:
    """rleToString: 6 bits per ASCII char (48..111), 5 data bits + continuation, sign-extended;
    counts from index 3 on are stored as differences to the count two places earlier."""
    out = []
    for i, c in enumerate(cnts):
        x = int(c) - int(cnts[i - 2]) if i > 2 else int(c)
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(chr(ch + 48))
    return "".join(out)


# --------------------------------------------------------------------------------------
# a17: post_process filter (post_process.py:61-74)
# --------------------------------------------------------------------------------------
def post_process_filter(existence, center, boundary, args) -> np.ndarray:
    """Indices kept by the three-threshold predicate, in input order (their position in the
    result is the new ``id``; ``score`` becomes ``area_score``)."""
    e, c, b = (np.asarray(v) for v in (existence, center, boundary))
    keep = ~((e < args.existence_score_thres) | (c < args.center_score_thres) | (b < args.boundary_score_thres))
    return np.nonzero(keep)[0]


# --------------------------------------------------------------------------------------
# A: summed-area table + box sums (north-star op (a); no reference counterpart — UNPINNED)
# --------------------------------------------------------------------------------------
def sat_build(field: torch.Tensor) -> torch.Tensor:
    """[..., H, W] fp32 -> [..., H+1, W+1] fp64 exclusive 2-D prefix sums."""
    s = field.double().cumsum(-1).cumsum(-2)
    return F.pad(s, (1, 0, 1, 0))


def box_sums(sat: torch.Tensor, boxes) -> torch.Tensor:
    """Sum of the field over the snapped window [floor y1:ceil y2, floor x1:ceil x2]
    (snapping rule of object_reasoning.py:502)."""
    out = []
    for b in boxes:
        x1, y1, x2, y2 = snap_box(b)
        out.append(sat[y2, x2] - sat[y1, x2] - sat[y2, x1] + sat[y1, x1])
    return torch.stack(out) if out else torch.zeros((0,), dtype=torch.float64)


# --------------------------------------------------------------------------------------
# C: mask-IoU NMS (north-star op (c); no reference counterpart — UNPINNED)
# --------------------------------------------------------------------------------------
def mask_nms_dense(masks: np.ndarray, scores, iou_threshold: float = 0.5) -> np.ndarray:
    """Greedy NMS in the order/tie-break of ``nms`` but with exact integer mask IoU:
    suppress j if inter/(area_i+area_j-inter) > thr evaluated as
    float32(inter)/float32(union) > float32(thr)."""
    m = np.asarray(masks).reshape(len(masks), -1).astype(bool)
    s = np.asarray(scores, dtype=np.float32)
    n = m.shape[0]
    order = np.argsort(-s, kind="stable")
    area = m.sum(1).astype(np.int64)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = np.float32(iou_threshold)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        if rest.size == 0:
            break
        inter = (m[rest] & m[i]).sum(1).astype(np.int64)
        union = area[i] + area[rest] - inter
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = inter.astype(np.float32) / union.astype(np.float32)
        suppressed[rest[iou > thr]] = True
    return np.asarray(keep, dtype=np.int64)
