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
