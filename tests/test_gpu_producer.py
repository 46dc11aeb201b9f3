"""configs[0] / configs[4] of BASELINE.json on the GPU: fields produced by the (random-init,
calibrated) DPT-L objectness net feed the CUDA reasoning path; the result must equal the CPU
oracle run on the very same field stack (512 proposals, 480x640)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from unmore_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def produced(dev):
    from unmore_b200.producer import FieldProducer
    torch.manual_seed(0)
    prod = FieldProducer().to(dev).eval()          # full ViT-L/16 trunk, fp32
    g = torch.Generator(device=dev).manual_seed(1)
    img = torch.rand((1, 3, 480, 640), generator=g, device=dev)
    prod.calibrate_random_init(img)
    fields = prod(img)
    return prod, img, fields


def test_producer_fields_are_usable(produced):
    _, _, fields = produced
    assert fields.shape == (1, 4, 480, 640) and fields.dtype == torch.float32 and fields.is_contiguous()
    assert torch.isfinite(fields).all()
    assert float(fields[0, 0].abs().max()) <= 1.0
    assert float(fields[0, 0].max()) > 0.5          # some proposals survive the max-sdf test
    assert 0.0 < float(fields[0, 3].min()) and float(fields[0, 3].max()) < 1.0


def test_config0_reasoning_on_net_fields_matches_oracle(produced, dev):
    """one 480x640 image, random-init objectness-net fields, 512 proposals: CUDA discovery + scoring
    against the CPU path (oracle port of the reference) on identical fields."""
    from unmore_b200.object_reasoning import Object_Discovery
    from unmore_b200.object_scoring import Object_Scoring
    _, _, fields = produced
    img_cpu = fields[0].cpu()
    props = synth.make_proposals(3, 512, 480, 640)
    args = O.make_args()
    ref = O.discover_image(img_cpu, props, args)
    det = Object_Discovery(device=dev).discover_image(fields[0], props)
    assert det.shape == ref.shape, (det.shape, ref.shape)
    if len(ref):
        side = np.maximum(ref[:, 2] - ref[:, 0], ref[:, 3] - ref[:, 1])[:, None]
        assert (np.abs(det - ref) <= 1e-5 * np.maximum(np.abs(ref), side)).all()
        s_ref = O.score_image(img_cpu, ref.tolist(), args)
        anns = Object_Scoring(device=dev).score_image(fields[0], ref.astype(np.float64).tolist())
        assert len(anns) == len(s_ref["score"])
        assert np.allclose(np.array([a["score"] for a in anns]), s_ref["score"], rtol=1e-5, atol=0)
        masks = np.stack([a["segmentation"]["mask"] for a in anns])
        assert np.array_equal(masks, s_ref["masks"])


def test_bf16_producer_runs_and_pipeline_accepts_it(produced, dev):
    """The throughput configuration of scripts/e2e_producer.py: bf16 autocast producer -> fp32 field stack."""
    from unmore_b200.pipeline import ReasoningPipeline
    from unmore_b200.producer import FieldProducer
    prod, img, f32 = produced
    fast = FieldProducer(prod.objectness_model, prod.binary_classifier_model, autocast_dtype=torch.bfloat16)
    f16 = fast(img)
    assert f16.dtype == torch.float32 and torch.isfinite(f16).all()
    assert float((f16 - f32).abs().mean()) < 0.05
    props = torch.from_numpy(synth.make_proposals(3, 512, 480, 640)).to(dev)[None].contiguous()
    r = ReasoningPipeline(dev, with_sat=False).run_chunk(f16, props)
    assert r["keep_counts"].shape == (1,)


@pytest.mark.gpu
def test_per_crop_nets_mode_vs_oracle_on_the_same_tiles():
    """The reference's ORIGINAL mode (nets on every crop, object_reasoning.py:311-333 / 398-417 / 496-512): RGB image ->
    batched crops (bit-exact vs the reference's Resize) -> ObjectnessNet / Binary_Classifier per crop -> the tile-path
    kernels.  The nets are the producer's (pinned against the reference's model code in tests/test_producer.py); the
    reasoning arithmetic is checked against the oracle fed with the SAME tiles."""
    import numpy as np
    from oracle import oracle as O
    from unmore_b200.object_reasoning import Object_Discovery
    from unmore_b200.object_scoring import Object_Scoring, unpack_masks
    from unmore_b200.producer import FieldProducer, PerCropNets
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    H, W = 240, 320
    img = torch.rand((3, H, W))
    prod = FieldProducer().to(dev).eval()
    nets = PerCropNets(prod.objectness_model, prod.binary_classifier_model)
    rng = np.random.default_rng(0)
    c = rng.uniform(20, 200, (20, 2)); wh = rng.uniform(24, 110, (20, 2))
    boxes = np.clip(np.concatenate([c, c + wh], axis=1), 0, [W, H, W, H])
    bt = torch.tensor(boxes, device=dev)
    crops = nets.crops(img.to(dev), bt)
    ref_crops = O.crops_for(img, boxes)                                       # the reference's crop + Resize, CPU
    assert torch.equal(crops.cpu(), ref_crops), "batched RGB crops must equal the per-box Resize loop bit for bit"
    prod.calibrate_random_init(crops[:8])                                     # non-degenerate heads on crops
    od = Object_Discovery(device=dev, tile_provider=nets)
    tiles = nets.fields(img.to(dev), bt)                                      # [N,3,128,128]
    ex = nets.existence(img.to(dev), bt)
    assert tiles.shape == (20, 3, 128, 128) and ex.shape == (20,)
    got_ex = od.existence_checking(img.to(dev), boxes)["existence_scores"]
    assert torch.allclose(got_ex, ex.cpu(), rtol=1e-5, atol=1e-7)
    # the oracle on the same tiles: its crop function is replaced by a lookup of the GPU-made tiles
    table = {tuple(np.round(b, 6)): k for k, b in enumerate(boxes)}
    t_cpu, ex_cpu = tiles.cpu(), ex.cpu()

    def lookup(image, proposals):
        idx = [table[tuple(np.round(np.asarray(p, dtype=np.float64), 6))] for p in proposals]
        return torch.cat([t_cpu[idx], ex_cpu[idx][:, None, None, None].expand(-1, 1, 128, 128)], dim=1)

    orig = O.crops_for
    O.crops_for = lookup
    try:
        args = O.make_args()
        cr_ref = O.center_reasoning(img, torch.tensor(boxes), args)
        cr = od.center_reasoning(img.to(dev), boxes)
        assert np.array_equal(cr["proposals_pass_singularity"].cpu().numpy(), cr_ref["proposals_pass_singularity"].numpy())
        assert np.array_equal(cr["splited_new_proposals"].cpu().numpy().reshape(-1, 4), cr_ref["splited_new_proposals"].numpy().reshape(-1, 4))
        r_ref = O.optimize_one_image_single_round(img, torch.tensor(boxes), args)
        r = od.optimize_one_image_single_round(img.to(dev), boxes)
        assert np.array_equal(r["labels"].cpu().numpy(), r_ref["labels"].numpy())
        ref_b, got_b = r_ref["updated_bboxes"].numpy(), r["updated_bboxes"].cpu().numpy()
        side = np.maximum(ref_b[:, 2] - ref_b[:, 0], ref_b[:, 3] - ref_b[:, 1])[:, None]
        assert (np.abs(got_b - ref_b) <= 1e-5 * np.maximum(np.abs(ref_b), np.maximum(side, 1e-30))).all()
        s_ref = O.score_image(img, boxes.tolist(), args)
        anns = Object_Scoring(device=dev, tile_provider=nets).score_image(img.to(dev), boxes.tolist())
        assert len(anns) == len(s_ref["score"])
        assert np.array_equal(np.array([a["bbox"] for a in anns], np.float32), s_ref["bbox"])
        assert np.allclose([a["score"] for a in anns], s_ref["score"], rtol=1e-5, atol=0)
        assert np.array_equal(np.stack([a["segmentation"]["mask"] for a in anns]), s_ref["masks"])
    finally:
        O.crops_for = orig
    det = od.discover_image(img.to(dev), boxes)      # the whole loop body runs in this mode (per-crop nets every round)
    assert det.ndim == 2 and det.shape[1] == 4
