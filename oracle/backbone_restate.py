"""TEST INFRASTRUCTURE — not product code.

CPU restatement (plain PyTorch functional ops) of the backbone the reference builds with detectron2's
`build_resnet_backbone` (configs/Base.yaml:2-12: DEPTH 50, STEM_OUT_CHANNELS 64, STRIDE_IN_1X1 False, OUT_FEATURES res2..res5,
default NORM FrozenBN, weights `detectron2://ImageNetPretrained/torchvision/R-50.pkl`), SURVEY.md §8 row f2, together with the
input normalisation of `preprocess_image` ((x - PIXEL_MEAN) / PIXEL_STD, Base.yaml:6-7).

detectron2 (0.4, README.md:28) is a pip dependency that is absent from /root/reference and from this image, so the algorithm is
restated from its published `modeling/backbone/resnet.py` (BasicStem: 7x7/2 conv + FrozenBN + ReLU + MaxPool(3,2,1);
BottleneckBlock: 1x1 -> 3x3 (carries the stride) -> 1x1, FrozenBN after each, projection shortcut when shape changes,
relu(out + shortcut)).  PINNED against an independent public implementation of the same network: with STRIDE_IN_1X1 False this
IS torchvision's `resnet50` (the checkpoint the reference loads is torchvision's, renamed) — tests/test_oracle_backbone.py maps
a seeded torchvision resnet50 (eval mode) into detectron2's parameter names and compares every stage output bit for bit.

Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

STAGES = (("res2", 3, 64, 256, 1), ("res3", 4, 128, 512, 2), ("res4", 6, 256, 1024, 2), ("res5", 3, 512, 2048, 2))
BN_EPS = 1e-5


def state_shapes() -> Dict[str, tuple]:
    """detectron2 parameter / buffer names of the R-50 backbone -> shapes."""
    shapes = {}

    def conv(prefix, cout, cin, k):
        shapes[prefix + ".weight"] = (cout, cin, k, k)
        for n in ("weight", "bias", "running_mean", "running_var"):
            shapes[f"{prefix}.norm.{n}"] = (cout,)

    conv("stem.conv1", 64, 3, 7)
    cin = 64
    for name, blocks, mid, cout, _ in STAGES:
        for i in range(blocks):
            p = f"{name}.{i}"
            if i == 0:
                conv(p + ".shortcut", cout, cin, 1)
            conv(p + ".conv1", mid, cin, 1)
            conv(p + ".conv2", mid, mid, 3)
            conv(p + ".conv3", cout, mid, 1)
            cin = cout
    return shapes


def _conv_bn(sd, p, x, stride=1, padding=0, relu=True):
    x = F.conv2d(x, sd[p + ".weight"], None, stride, padding)
    scale = sd[p + ".norm.weight"] * (sd[p + ".norm.running_var"] + BN_EPS).rsqrt()          # FrozenBatchNorm2d.forward
    bias = sd[p + ".norm.bias"] - sd[p + ".norm.running_mean"] * scale
    x = x * scale.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1)
    return F.relu(x) if relu else x


def normalize(images: torch.Tensor, pixel_mean, pixel_std) -> torch.Tensor:
    m = torch.tensor(pixel_mean, dtype=images.dtype).view(1, 3, 1, 1)
    s = torch.tensor(pixel_std, dtype=images.dtype).view(1, 3, 1, 1)
    return (images - m) / s


def resnet50(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> Dict[str, torch.Tensor]:
    """x: normalised images [N,3,H,W] -> {'res2': [N,256,H/4,W/4], ..., 'res5': [N,2048,H/32,W/32]}."""
    x = _conv_bn(sd, "stem.conv1", x, 2, 3)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    out = {}
    for name, blocks, _, _, stride in STAGES:
        for i in range(blocks):
            p = f"{name}.{i}"
            s = stride if i == 0 else 1
            y = _conv_bn(sd, p + ".conv1", x)
            y = _conv_bn(sd, p + ".conv2", y, s, 1)
            y = _conv_bn(sd, p + ".conv3", y, relu=False)
            sc = _conv_bn(sd, p + ".shortcut", x, s, relu=False) if (p + ".shortcut.weight") in sd else x
            x = F.relu(y + sc)
        out[name] = x
    return out


def from_torchvision(tv_state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """torchvision resnet50 state dict -> detectron2 names (what `R-50.pkl` is)."""
    sd = {}

    def put(dst, conv, bn):
        sd[dst + ".weight"] = tv_state[conv + ".weight"]
        for n in ("weight", "bias", "running_mean", "running_var"):
            sd[f"{dst}.norm.{n}"] = tv_state[f"{bn}.{n}"]

    put("stem.conv1", "conv1", "bn1")
    for li, (name, blocks, _, _, _) in enumerate(STAGES, start=1):
        for i in range(blocks):
            for c in (1, 2, 3):
                put(f"{name}.{i}.conv{c}", f"layer{li}.{i}.conv{c}", f"layer{li}.{i}.bn{c}")
            if i == 0:
                put(f"{name}.{i}.shortcut", f"layer{li}.{i}.downsample.0", f"layer{li}.{i}.downsample.1")
    return sd
