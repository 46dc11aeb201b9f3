"""The field producer (unmore_b200/producer.py) against the reference's own model code.

The fixture tests/golden/producer.npz holds the outputs of the REFERENCE ObjectnessNet
(models/objectness_net.py, run from /root/reference by oracle/gen_producer_golden.py) for this
repo's seeded weights; here the seeded model is rebuilt and must reproduce them."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import gen_producer_golden as G  # noqa: E402  (test infrastructure)
from unmore_b200 import producer as P  # noqa: E402


@pytest.fixture(scope="module")
def seeded():
    return G.seeded_model()


def test_objectness_net_matches_reference_outputs(golden_dir, seeded):
    g = np.load(os.path.join(golden_dir, "producer.npz"))
    if abs(G.weight_checksum(seeded) - float(g["checksum"])) > 1e-6 * float(g["checksum"]):
        pytest.skip("seeded weights differ on this torch build; regenerate with oracle/gen_producer_golden.py")
    with torch.no_grad():
        out = seeded(G.seeded_input())
    assert set(out) == {"center_fields", "sdf_maps"}
    for k in ("center_fields", "sdf_maps"):
        ref = g[k]
        assert out[k].shape == ref.shape
        assert np.abs(out[k].numpy() - ref).max() <= 1e-5 * np.abs(ref).max(), k
    assert out["sdf_maps"].abs().max() <= 1.0   # tanh head


def test_state_dict_keys_are_the_reference_checkpoint_keys(golden_dir, seeded):
    """object_reasoning.py:71-72 loads ckpt['model_state_dict'] strictly: same key set here."""
    g = np.load(os.path.join(golden_dir, "producer.npz"))
    assert sorted(seeded.state_dict().keys()) == list(g["keys"])
    sd = seeded.state_dict()
    assert sd["backbone.pretrained.model.pos_embed"].shape == (1, 577, 1024)
    assert sd["backbone.pretrained.act_postprocess1.4.weight"].shape == (256, 256, 4, 4)
    assert sd["backbone.pretrained.act_postprocess4.4.weight"].shape == (1024, 1024, 3, 3)
    assert sd["backbone.scratch.layer1_rn.weight"].shape == (256, 256, 3, 3)
    assert sd["sdf_prediction_head.3.weight"].shape == (1, 1024, 1, 1)
    assert sd["center_field_prediction_head.6.weight"].shape == (2, 1024, 1, 1)


def _small_vit():
    return P.ViTLarge16(dim=32, depth=24, num_heads=2, num_classes=0)   # narrow trunk: the decoder/heads are what is tested


def test_head_variants_follow_the_reference_args():
    """objectness_net.py:119-165: relu stack without use_bg_sdf or with 'relu'; linear stack + tanh / sin / nothing otherwise."""
    import argparse
    for act, bg, n_relu, last in (("tanh", True, 0, torch.nn.Tanh), ("sine", True, 0, P._Sin), (None, True, 0, torch.nn.Conv2d),
                                  ("relu", True, 3, torch.nn.Conv2d), ("tanh", False, 3, torch.nn.Conv2d)):
        net = P.ObjectnessNet(args=argparse.Namespace(sdf_activation=act, use_bg_sdf=bg), vit=_small_vit())
        h = net.sdf_prediction_head
        assert sum(isinstance(m, torch.nn.ReLU) for m in h) == n_relu, (act, bg)
        assert isinstance(h[-1], last), (act, bg)
        assert sum(isinstance(m, torch.nn.ReLU) for m in net.center_field_prediction_head) == 3
    with pytest.raises(NotImplementedError):
        P.ObjectnessNet(backbone_type="resnet50")
    with pytest.raises(NotImplementedError):
        P.ObjectnessNet(sdf_activation="gelu", vit=_small_vit())


def test_dpt_rejects_sizes_that_do_not_realign():
    net = P.DPTLarge(vit=_small_vit())
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 48, 64))


def test_dense_classifier_is_the_hoisted_per_crop_classifier():
    """fc and the binary head are linear, so the mean of the dense LOGIT map over a crop equals the
    per-crop logit (objectness_net.py:220-223) when the crop is the whole input."""
    torch.manual_seed(3)
    clf = P.Binary_Classifier().eval()
    x = torch.rand(2, 3, 96, 128)
    with torch.no_grad():
        per_crop = clf(x)                                   # sigmoid(head(fc(avgpool(trunk))))
        f = clf._trunk(x)
        r, h = clf.classifier_backbone, clf.binary_classification_head
        logit = torch.nn.functional.conv2d(torch.nn.functional.conv2d(f, r.fc.weight[:, :, None, None], r.fc.bias),
                                           h.weight[:, :, None, None], h.bias)
        dense = clf.dense(x)
    assert torch.allclose(torch.sigmoid(logit.mean(dim=(2, 3))), per_crop, atol=1e-6)
    assert dense.shape == (2, 1, 96, 128) and float(dense.min()) > 0 and float(dense.max()) < 1


def test_field_producer_stack_layout():
    torch.manual_seed(5)
    small = P.ObjectnessNet(vit=_small_vit())
    fp = P.FieldProducer(small.eval(), P.Binary_Classifier().eval())
    x = torch.rand(1, 3, 64, 96)
    out = fp(x)
    assert out.shape == (1, 4, 64, 96) and out.dtype == torch.float32
    with torch.no_grad():
        pred = small(x)
    assert torch.equal(out[:, 0:1], pred["sdf_maps"]) and torch.equal(out[:, 1:3], pred["center_fields"])
    assert torch.equal(out[:, 3:4], fp.binary_classifier_model.dense(x))


def test_calibrated_random_init_gives_non_degenerate_fields():
    torch.manual_seed(7)
    fp = P.FieldProducer(P.ObjectnessNet(vit=_small_vit()).eval(), P.Binary_Classifier().eval())
    x = torch.rand(1, 3, 96, 128)
    fp.calibrate_random_init(x)
    out = fp(x)
    sdf, cen, ex = out[0, 0], out[0, 1:3], out[0, 3]
    assert float(sdf.abs().max()) <= 1.0 and 0.3 < float(torch.atanh(sdf.clamp(-0.999, 0.999)).std()) < 1.5
    assert 0.3 < float(cen.std()) < 1.0
    assert 0.05 < float(ex.min()) and float(ex.max()) < 0.99 and float(ex.std()) > 0.02
