"""The producer side of the hot path: per-image prior fields from the objectness nets.

The reasoning kernels consume one ``[4, H, W]`` fp32 stack per image,
``[sdf, center_row, center_col, existence]``.  In the reference the nets run on
every 128x128 crop in every round (object_reasoning.py:398-417); the north-star hoists them out of
the loop and runs them ONCE per image.  This module is that producer, in plain PyTorch (it is
plumbing, not the product; BASELINE.json: "the DPT-based models/objectness_net.py forward remains
the unchanged PyTorch producer of the fields"):

* ``ObjectnessNet`` — DPT-Large (ViT-L/16 backbone, hooks after blocks 5/11/17/23, "project"
  readout, reassemble to 1/4, 1/8, 1/16, 1/32 resolution, four RefineNet fusion blocks, x2 bilinear)
  + the center-field head (2 ch) and the boundary-distance head (1 ch).  Architecture after
  models/objectness_net.py:37-183, models/dpt/models.py:26-94, models/dpt/vit.py:104-336,
  models/dpt/blocks.py:68-115, 248-383.  The reference builds the ViT with ``timm`` (absent here and
  on the GPU box); ``ViTLarge16`` restates timm's ``vit_large_patch16_384`` forward
  (pre-norm blocks, LayerNorm eps 1e-6, exact GELU, qkv bias) — *timm internals are parity-unpinned*;
  everything around the blocks is pinned against the reference's own code by
  ``oracle/gen_producer_golden.py`` (fixture ``tests/golden/producer.npz``).
* module / parameter names reproduce the reference's ``state_dict`` keys, so a checkpoint trained
  with train_objectness_net.py (``ckpt['model_state_dict']``, object_reasoning.py:71-72) loads with
  ``load_state_dict`` unchanged.
* ``Binary_Classifier`` — ResNet-50 + Linear(1000, 1) + sigmoid (objectness_net.py:203-223).  The
  reference evaluates it per crop; hoisted out of the loop it becomes a *dense* predictor: the
  global average pool is dropped, ``fc`` and the head act as 1x1 convolutions on the stride-32
  feature map, and the sigmoid map is bilinearly upsampled to H x W.  Its box mean is what
  ``existence_checking`` reduces.  (The per-image existence map has no reference counterpart —
  SURVEY.md section 0 — so this definition is this repo's, stated here and in DESIGN.md.)
* ``FieldProducer`` — both nets -> ``[B, 4, H, W]`` stacks, written straight into the HBM-resident
  field batch the kernels read.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# ViT-L/16 (timm "vit_large_patch16_384" layout: same parameter names, same forward)
# ---------------------------------------------------------------------------------------------
class _Attention(nn.Module):
    def __init__(self, dim: int, num_heads: int):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b, n, c = x.shape
        qkv = self.qkv(x).view(b, n, 3, self.num_heads, c // self.num_heads).permute(2, 0, 3, 1, 4)
        out = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])   # softmax(q k^T * scale) v
        return self.proj(out.transpose(1, 2).reshape(b, n, c))


class _Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim: int, num_heads: int, mlp_ratio: float):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class _PatchEmbed(nn.Module):
    def __init__(self, dim: int, patch: int):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)


class ViTLarge16(nn.Module):
    """Token trunk.  ``forward_taps`` returns the token sequences after the hooked blocks; the position
    embedding (trained on a 24x24 grid + class token) is bilinearly resized to the input's patch grid
    (models/dpt/vit.py:148-201), so any H, W that are multiples of 16 work."""

    def __init__(self, dim: int = 1024, depth: int = 24, num_heads: int = 16, mlp_ratio: float = 4.0,
                 patch: int = 16, train_grid: int = 24, num_classes: int = 1000):
        super().__init__()
        self.patch_size = patch
        self.patch_embed = _PatchEmbed(dim, patch)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, train_grid * train_grid + 1, dim))
        self.blocks = nn.ModuleList([_Block(dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.head = nn.Linear(dim, num_classes) if num_classes else nn.Identity()  # unused; keeps timm's keys
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)

    def resized_pos_embed(self, gh: int, gw: int) -> torch.Tensor:
        tok, grid = self.pos_embed[:, :1], self.pos_embed[0, 1:]
        g0 = int(math.sqrt(grid.shape[0]))
        grid = grid.reshape(1, g0, g0, -1).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, size=(gh, gw), mode="bilinear")
        return torch.cat([tok, grid.permute(0, 2, 3, 1).reshape(1, gh * gw, -1)], dim=1)

    def forward_taps(self, x: torch.Tensor, taps):
        b, _, h, w = x.shape
        gh, gw = h // self.patch_size, w // self.patch_size
        t = self.patch_embed.proj(x).flatten(2).transpose(1, 2)
        t = torch.cat([self.cls_token.expand(b, -1, -1), t], dim=1) + self.resized_pos_embed(gh, gw)
        out = []
        last = max(taps)
        for i, blk in enumerate(self.blocks):
            t = blk(t)
            if i in taps:
                out.append(t)
            if i == last:
                break   # the final norm / head never reach the decoder
        return out, (gh, gw)


# ---------------------------------------------------------------------------------------------
# DPT decoder
# ---------------------------------------------------------------------------------------------
class _ProjectReadout(nn.Module):
    """Concatenate the class token to every patch token, project back to `dim`, GELU (vit.py:75-86)."""

    def __init__(self, dim: int):
        super().__init__()
        self.project = nn.Sequential(nn.Linear(2 * dim, dim), nn.GELU())

    def forward(self, t):
        patches = t[:, 1:]
        return self.project(torch.cat([patches, t[:, :1].expand_as(patches)], dim=-1))


def _reassemble(dim: int, ch: int, mode: str) -> nn.Sequential:
    """Indices match the reference's nn.Sequential (readout, transpose, unflatten, conv, resample) so the
    state_dict keys are `.0.project.0.*`, `.3.*`, `.4.*` (vit.py:259-336)."""
    layers = [_ProjectReadout(dim), nn.Identity(), nn.Identity(), nn.Conv2d(dim, ch, 1)]
    if mode == "up4":
        layers.append(nn.ConvTranspose2d(ch, ch, kernel_size=4, stride=4))
    elif mode == "up2":
        layers.append(nn.ConvTranspose2d(ch, ch, kernel_size=2, stride=2))
    elif mode == "down2":
        layers.append(nn.Conv2d(ch, ch, kernel_size=3, stride=2, padding=1))
    return nn.Sequential(*layers)


class _ResidualConvUnit(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv1 = nn.Conv2d(ch, ch, 3, padding=1)
        self.conv2 = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv2(F.relu(self.conv1(F.relu(x)))) + x


class _FusionBlock(nn.Module):
    """RefineNet-style fusion (blocks.py:318-383): [skip through unit 1 added,] unit 2, x2 bilinear
    (align_corners=True), 1x1 conv."""

    def __init__(self, ch: int):
        super().__init__()
        self.out_conv = nn.Conv2d(ch, ch, 1)
        self.resConfUnit1 = _ResidualConvUnit(ch)
        self.resConfUnit2 = _ResidualConvUnit(ch)

    def forward(self, x, skip=None):
        if skip is not None:
            x = x + self.resConfUnit1(skip)
        x = self.resConfUnit2(x)
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        return self.out_conv(x)


class _Holder(nn.Module):
    """Plain attribute container (the reference hangs sub-modules on bare nn.Module()s)."""


class DPTLarge(nn.Module):
    """DPT with the ViT-L/16 trunk, features=256, no head: [B,3,H,W] -> [B,256,H,W]
    (models/dpt/models.py:26-94 with backbone 'vitl16_384', readout 'project', use_bn False).
    H and W must be multiples of 32."""

    HOOKS = (5, 11, 17, 23)
    CHANNELS = (256, 512, 1024, 1024)

    def __init__(self, features: int = 256, vit: Optional[ViTLarge16] = None):
        super().__init__()
        vit = vit or ViTLarge16()
        dim = vit.pos_embed.shape[-1]
        self.pretrained = _Holder()
        self.pretrained.model = vit
        for k, (ch, mode) in enumerate(zip(self.CHANNELS, ("up4", "up2", "same", "down2")), start=1):
            setattr(self.pretrained, f"act_postprocess{k}", _reassemble(dim, ch, mode))
        self.scratch = _Holder()
        for k, ch in enumerate(self.CHANNELS, start=1):
            setattr(self.scratch, f"layer{k}_rn", nn.Conv2d(ch, features, 3, padding=1, bias=False))
            setattr(self.scratch, f"refinenet{k}", _FusionBlock(features))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape[-2] % 32 or x.shape[-1] % 32:
            raise ValueError(f"DPT input must be a multiple of 32 in both dimensions, got {tuple(x.shape[-2:])}")
        taps, (gh, gw) = self.pretrained.model.forward_taps(x, self.HOOKS)
        feats = []
        for k, t in enumerate(taps, start=1):
            post = getattr(self.pretrained, f"act_postprocess{k}")
            t = post[0](t).transpose(1, 2).unflatten(2, (gh, gw))
            for layer in list(post)[3:]:
                t = layer(t)
            feats.append(getattr(self.scratch, f"layer{k}_rn")(t))
        path = self.scratch.refinenet4(feats[3])
        path = self.scratch.refinenet3(path, feats[2])
        path = self.scratch.refinenet2(path, feats[1])
        path = self.scratch.refinenet1(path, feats[0])
        return F.interpolate(path, scale_factor=2, mode="bilinear", align_corners=True)


class _Sin(nn.Module):
    def forward(self, x):
        return torch.sin(x)


def _head(in_ch: int, out_ch: int, relu: bool, last: Optional[nn.Module]) -> nn.Sequential:
    """1x1(512) - 3x3(512) - 1x1(1024) - 1x1(out) with the reference's Sequential indices: ReLUs occupy
    slots 1/3/5 when present, so conv keys are .0/.2/.4/.6 (relu) or .0/.1/.2/.3 (linear stack)."""
    convs = [nn.Conv2d(in_ch, 512, 1), nn.Conv2d(512, 512, 3, padding=1), nn.Conv2d(512, 1024, 1), nn.Conv2d(1024, out_ch, 1)]
    layers = []
    for i, c in enumerate(convs):
        layers.append(c)
        if relu and i < 3:
            layers.append(nn.ReLU())
    if last is not None:
        layers.append(last)
    return nn.Sequential(*layers)


class ObjectnessNet(nn.Module):
    """models/objectness_net.py:37-183 for backbone_type 'dpt_large'.  forward(images [B,3,H,W]) ->
    {'center_fields': [B,2,H,W], 'sdf_maps': [B,1,H,W]}.  ``sdf_activation`` / ``use_bg_sdf`` select the
    boundary head exactly as the reference's args do (script.sh trains with tanh + use_bg_sdf)."""

    def __init__(self, device=None, image_size=None, backbone_type: str = "dpt_large", args=None,
                 sdf_activation: Optional[str] = "tanh", use_bg_sdf: bool = True, vit: Optional[ViTLarge16] = None):
        super().__init__()
        if backbone_type != "dpt_large":
            raise NotImplementedError(f"backbone_type {backbone_type!r}: only 'dpt_large' (the released recipe) is built")
        if args is not None:
            sdf_activation = getattr(args, "sdf_activation", sdf_activation)
            use_bg_sdf = getattr(args, "use_bg_sdf", use_bg_sdf)
        self.image_size, self.backbone_type = image_size, backbone_type
        self.backbone = DPTLarge(features=256, vit=vit)   # vit: inject a narrower trunk (tests); default ViT-L/16
        self.center_field_prediction_head = _head(256, 2, relu=True, last=None)
        if not use_bg_sdf or sdf_activation == "relu":
            self.sdf_prediction_head = _head(256, 1, relu=True, last=None)
        elif sdf_activation in ("tanh", "sine", None):
            last = {"tanh": nn.Tanh(), "sine": _Sin(), None: None}[sdf_activation]
            self.sdf_prediction_head = _head(256, 1, relu=False, last=last)
        else:
            raise NotImplementedError(f"sdf_activation {sdf_activation!r}")
        if device is not None:
            self.to(device)

    def forward(self, images: torch.Tensor):
        feat = self.backbone(images)
        return {"center_fields": self.center_field_prediction_head(feat), "sdf_maps": self.sdf_prediction_head(feat)}

    get_prediction = forward


class Binary_Classifier(nn.Module):
    """objectness_net.py:203-223.  ``forward`` is the reference's per-crop form ([B,3,h,w] -> [B,1]);
    ``dense`` is the hoisted per-image form: [B,3,H,W] -> existence map [B,1,H,W] in (0,1)."""

    def __init__(self, device=None, image_size=None, args=None):
        super().__init__()
        import torchvision
        self.image_size = image_size
        self.classifier_backbone = torchvision.models.resnet50(weights=None)
        self.binary_classification_head = nn.Linear(1000, 1)
        self.sigmoid = nn.Sigmoid()
        if device is not None:
            self.to(device)

    def forward(self, images):
        return self.sigmoid(self.binary_classification_head(self.classifier_backbone(images)))

    def _trunk(self, x):
        r = self.classifier_backbone
        x = r.maxpool(r.relu(r.bn1(r.conv1(x))))
        return r.layer4(r.layer3(r.layer2(r.layer1(x))))

    def dense(self, images: torch.Tensor) -> torch.Tensor:
        r = self.classifier_backbone
        f = self._trunk(images)                                             # [B,2048,H/32,W/32]
        logits = F.conv2d(f, r.fc.weight[:, :, None, None], r.fc.bias)      # fc as a 1x1 conv
        h = self.binary_classification_head
        logit = F.conv2d(logits, h.weight[:, :, None, None], h.bias)        # [B,1,H/32,W/32]
        return F.interpolate(torch.sigmoid(logit), size=images.shape[-2:], mode="bilinear", align_corners=False)


class FieldProducer(nn.Module):
    """images [B,3,H,W] in [0,1] -> field stacks [B,4,H,W] fp32 = [sdf, center_row, center_col, existence]
    (the channel order of ``unmore_b200.ops.Channels``; the center field's channels are (row, col) as the
    reference's training target builds them, datasets.py:200-206)."""

    def __init__(self, objectness: Optional[ObjectnessNet] = None, classifier: Optional[Binary_Classifier] = None,
                 autocast_dtype: Optional[torch.dtype] = None):
        super().__init__()
        self.objectness_model = objectness or ObjectnessNet()
        self.binary_classifier_model = classifier or Binary_Classifier()
        self.autocast_dtype = autocast_dtype

    @torch.no_grad()
    def forward(self, images: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        b, _, h, w = images.shape
        if out is None:
            out = torch.empty((b, 4, h, w), dtype=torch.float32, device=images.device)
        ctx = (torch.autocast(images.device.type, dtype=self.autocast_dtype) if self.autocast_dtype is not None
               else torch.autocast(images.device.type, enabled=False))
        with ctx:
            pred = self.objectness_model(images)
            exist = self.binary_classifier_model.dense(images)
        out[:, 0:1].copy_(pred["sdf_maps"])
        out[:, 1:3].copy_(pred["center_fields"])
        out[:, 3:4].copy_(exist)
        return out

    @torch.no_grad()
    def calibrate_random_init(self, images: torch.Tensor, sdf_std: float = 1.0, center_std: float = 0.6,
                              exist_logit_std: float = 2.0) -> None:
        """Random-init nets put out ~1e-2 everywhere (sdf ~ 0 -> every proposal dies in round 0), which
        would make configs[0]/[4] a degenerate workload.  This rescales the LAST convolution of each head
        (and the classifier head) so that, on ``images``, the boundary head's pre-activation has standard
        deviation ``sdf_std``, the center field ``center_std`` and the existence logit ``exist_logit_std``:
        still random weights of the reference architecture, but fields with objects-sized structure in
        them.  Benchmarks that use it say so in their ``config``."""
        feat = self.objectness_model.backbone(images)

        def rescale(head: nn.Sequential, target: float):
            convs = [m for m in head if isinstance(m, nn.Conv2d)]
            pre = feat
            for m in head:
                if m is convs[-1]:
                    break
                pre = m(pre)
            y = F.conv2d(pre, convs[-1].weight)       # without bias
            mean = y.mean(dim=(0, 2, 3))
            s = float((y - mean[None, :, None, None]).std())
            if s > 0:
                convs[-1].weight.mul_(target / s)
                convs[-1].bias.copy_(-mean * (target / s))   # centred: random features carry a large constant part

        rescale(self.objectness_model.sdf_prediction_head, sdf_std)
        rescale(self.objectness_model.center_field_prediction_head, center_std)
        clf = self.binary_classifier_model
        f = clf._trunk(images)
        logits = F.conv2d(f, clf.classifier_backbone.fc.weight[:, :, None, None], clf.classifier_backbone.fc.bias)
        h = clf.binary_classification_head
        y = F.conv2d(logits, h.weight[:, :, None, None])
        s = float(y.std())
        if s > 0:
            h.weight.mul_(exist_logit_std / s)
            h.bias.fill_(float(-(y * (exist_logit_std / s)).mean()) + 0.5)   # centred, slightly positive


class PerCropNets:
    """The reference's ORIGINAL mode: the nets run on every 128x128 crop of the RGB image, every time a stage needs
    them (object_reasoning.py:311-333, 398-417, 496-512; object_scoring.py:112-157) — instead of once per image.

    A *tile provider* for ``Object_Discovery`` / ``Object_Scoring`` (``tile_provider=PerCropNets(...)``): the Python
    loop of per-box ``image[:, y1:y2, x1:x2]`` + ``Resize`` + four device->host syncs per box becomes ONE batched
    ``unmore_crop_resize`` (either resize mode) of all boxes, the crops go through the nets in batches (50 / 128 like
    the reference), and the reasoning stages consume the resulting tiles on the device (the ``*_from_tiles`` C-ABI
    entry points).  ``image`` is then the RGB image [3, H, W], exactly as in the reference."""

    def __init__(self, objectness_model: nn.Module, binary_classifier_model: nn.Module, antialias: bool = False,
                 objectness_batch: int = 50, classifier_batch: int = 128):
        self.objectness_model = objectness_model.eval()
        self.binary_classifier_model = binary_classifier_model.eval()
        self.antialias = antialias
        self.objectness_batch = objectness_batch
        self.classifier_batch = classifier_batch

    def crops(self, image: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
        """[3,H,W] image + [N,4] boxes -> [N,3,128,128] crops (a2, batched on the device)."""
        from . import ops
        img = image.to(torch.float32)
        return ops.crop_resize(img[None].contiguous(), boxes.reshape(1, -1, 4).contiguous(), [0, 1, 2], antialias=self.antialias)[0]

    @torch.no_grad()
    def fields(self, image: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
        """-> [N,3,128,128] = (sdf, center_row, center_col) tiles of every box (get_prediction_with_proposals)."""
        crops = self.crops(image, boxes)
        out = torch.empty((crops.shape[0], 3, crops.shape[2], crops.shape[3]), dtype=torch.float32, device=crops.device)
        for b0 in range(0, crops.shape[0], self.objectness_batch):
            pred = self.objectness_model(crops[b0:b0 + self.objectness_batch])
            out[b0:b0 + self.objectness_batch, 0:1] = pred["sdf_maps"]
            out[b0:b0 + self.objectness_batch, 1:3] = pred["center_fields"]
        return out

    @torch.no_grad()
    def existence(self, image: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
        """-> [N] existence scores, one classifier output per crop (existence_checking)."""
        crops = self.crops(image, boxes)
        out = [self.binary_classifier_model(crops[b0:b0 + self.classifier_batch]).reshape(-1)
               for b0 in range(0, crops.shape[0], self.classifier_batch)]
        return torch.cat(out) if out else torch.zeros((0,), dtype=torch.float32, device=crops.device)
