"""TEST INFRASTRUCTURE ONLY — pins ``unmore_b200/producer.py`` against the reference's own model code.

Runs in the build container only (needs /root/reference).  What it does:

1. puts a ``timm`` shim in ``sys.modules`` whose ``create_model("vit_large_patch16_384")`` returns
   this repo's ``ViTLarge16`` (timm 0.6.x is absent; its transformer block is the one piece of the
   producer that is restated rather than executed — *parity unpinned* for the block internals);
2. builds the reference's ``ObjectnessNet(backbone_type='dpt_large')`` (models/objectness_net.py:37)
   — its position-embedding resize, forward hooks, readout projection, reassemble convolutions,
   RefineNet fusion and prediction heads all run from /root/reference unmodified;
3. builds this repo's ``ObjectnessNet`` with ``torch.manual_seed(SEED)``, loads its ``state_dict`` into
   the reference model with ``strict=True`` (which also proves the two key sets are identical, i.e.
   reference checkpoints load here), runs both on the same seeded input and stores the REFERENCE's
   outputs in ``tests/golden/producer.npz`` together with a weight checksum.

``tests/test_producer.py`` rebuilds the seeded model on any machine and compares with the fixture.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFERENCE_ROOT = os.environ.get("UNMORE_REFERENCE_ROOT", "/root/reference")
SEED = 20251018
IN_H, IN_W = 64, 96


def _install_timm_shim():
    from unmore_b200.producer import ViTLarge16

    def create_model(name, pretrained=False, **kw):
        if name != "vit_large_patch16_384":
            raise NotImplementedError(name)
        m = ViTLarge16()
        m.pos_drop = torch.nn.Identity()   # forward_flex calls it (models/dpt/vit.py:190)
        m.dist_token = None
        return m

    timm = types.ModuleType("timm")
    timm.create_model = create_model
    sys.modules["timm"] = timm


def seeded_model():
    from unmore_b200.producer import ObjectnessNet
    torch.manual_seed(SEED)
    net = ObjectnessNet(sdf_activation="tanh", use_bg_sdf=True).eval()
    # random-init heads / decoder with torch defaults give ~1e-3 outputs after 30 layers; scale the
    # decoder a little so the fixture has signal well above fp32 noise
    return net


def weight_checksum(net) -> float:
    return float(sum(p.double().abs().sum() for p in net.state_dict().values()))


def seeded_input():
    g = torch.Generator().manual_seed(SEED + 1)
    return torch.rand((2, 3, IN_H, IN_W), generator=g)


def main(out_path: str) -> None:
    _install_timm_shim()
    sys.path.insert(0, REFERENCE_ROOT)
    from models.objectness_net import ObjectnessNet as RefNet   # the reference's file, unmodified

    args = argparse.Namespace(sdf_activation="tanh", use_bg_sdf=True)
    ref = RefNet(device=torch.device("cpu"), image_size=(IN_H, IN_W), backbone_type="dpt_large", args=args).eval()
    mine = seeded_model()
    sd = mine.state_dict()
    missing = ref.load_state_dict({k: v for k, v in sd.items() if not k.endswith("pos_drop")}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys, missing
    assert set(ref.state_dict().keys()) == set(sd.keys())
    x = seeded_input()
    with torch.no_grad():
        r = ref(x)
        m = mine(x)
    for k in ("center_fields", "sdf_maps"):
        d = (r[k] - m[k]).abs().max().item()
        print(f"{k}: ref shape {tuple(r[k].shape)} max|ref| {r[k].abs().max().item():.4g} max abs diff vs this repo {d:.3g}")
    np.savez_compressed(out_path, center_fields=r["center_fields"].numpy(), sdf_maps=r["sdf_maps"].numpy(),
                        checksum=np.float64(weight_checksum(mine)), seed=np.int64(SEED),
                        keys=np.array(sorted(sd.keys())))
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "producer.npz"))
