"""ResNet-50 backbone on the tensor-core engine — SURVEY.md §8 row f2 (the producer of the camera head's `res3..res5` inputs).

Mirrors what the reference gets from detectron2 (`cfg.MODEL.BACKBONE.NAME = "build_resnet_backbone"`, configs/Base.yaml:2-12:
DEPTH 50, STEM_OUT_CHANNELS 64, STRIDE_IN_1X1 False, OUT_FEATURES res2..res5, FrozenBN, weights = torchvision's R-50 renamed):
same registry name, same parameter / buffer names (`stem.conv1.weight`, `stem.conv1.norm.running_mean`,
`res2.0.shortcut.weight`, `res4.5.conv3.norm.bias`, ...), same `output_shape()`, and `forward` fuses the
`(x - PIXEL_MEAN) / PIXEL_STD` of `preprocess_image` (siamese_planeTR.py `preprocess_image`, Base.yaml:6-7).

Every convolution is a GEMM on NHWC 16-bit hi/lo planes (3 passes ~ fp32): 1x1 -> `nsac_gemm_split`, 3x3 stride 1 ->
`nsac_conv3x3_split` (implicit GEMM, 4-D TMA gather), 3x3 stride 2 -> `nsac_conv3x3_split_strided` (same kernel, the tensor
map's traversal stride picks every second pixel), stem 7x7/2 -> im2col + GEMM; FrozenBN is folded into weights and bias, ReLU runs in the GEMM epilogue, and `relu(out + shortcut)` of every bottleneck
block runs in the epilogue of its last GEMM (`nsac_gemm_split_residual`): activations exist only as hi/lo planes between layers
(r1 went through fp32 NHWC + a separate add kernel: 11 of 62 ms).  uint8 images (what the reference's loader delivers) take the
fast stem: one exact fp16 plane of raw pixels, normalisation folded into the weights, borders recomputed exactly
(`nsac_stem_im2col_u8` / `nsac_stem_border_fix`).  `forward(..., planes=True)` hands the NHWC planes straight to the camera
head's pixel network (`PlaneFeatures`), skipping the NCHW fp32 round trip.  No cuDNN, no CPU path.
"""
from __future__ import annotations

import os

from typing import Dict

import torch
from torch import nn

from . import ops
from .compat import Registry, ShapeSpec

__all__ = ["BACKBONE_REGISTRY", "ResNet50Backbone", "PlaneFeatures", "build_resnet_backbone", "build_backbone"]

BACKBONE_REGISTRY = Registry("BACKBONE")
STAGES = (("res2", 3, 64, 256, 1), ("res3", 4, 128, 512, 2), ("res4", 6, 256, 1024, 2), ("res5", 3, 512, 2048, 2))
BN_EPS = 1e-5


class _FrozenBN(nn.Module):
    """detectron2.layers.FrozenBatchNorm2d: four buffers, y = x * (w * rsqrt(var + eps)) + (b - mean * w * rsqrt(var + eps))."""

    def __init__(self, c: int):
        super().__init__()
        self.register_buffer("weight", torch.ones(c))
        self.register_buffer("bias", torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c) - BN_EPS)


class _ConvBN(nn.Module):
    """detectron2.layers.Conv2d(bias=False, norm=FrozenBN): `.weight` + `.norm.*`."""

    def __init__(self, cin: int, cout: int, k: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")      # c2_msra_fill
        self.norm = _FrozenBN(cout)

    def folded(self):
        scale = self.norm.weight * (self.norm.running_var + BN_EPS).rsqrt()
        bias = self.norm.bias - self.norm.running_mean * scale
        w = self.weight.detach() * scale.view(-1, 1, 1, 1)
        return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous(), bias.contiguous()      # [Cout, (ky,kx,cin)]


class _Bottleneck(nn.Module):
    def __init__(self, cin: int, mid: int, cout: int, stride: int, project: bool):
        super().__init__()
        self.stride = stride
        if project:
            self.shortcut = _ConvBN(cin, cout, 1)
        self.conv1 = _ConvBN(cin, mid, 1)
        self.conv2 = _ConvBN(mid, mid, 3)
        self.conv3 = _ConvBN(mid, cout, 1)


class _Stem(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = _ConvBN(3, 64, 7)


class ResNet50Backbone(nn.Module):
    def __init__(self, cfg=None, input_shape=None):
        super().__init__()
        if cfg is not None:
            assert cfg.MODEL.RESNETS.DEPTH == 50 and not cfg.MODEL.RESNETS.STRIDE_IN_1X1, "only the reference's R-50 (stride in the 3x3)"
            self.pixel_mean, self.pixel_std = list(cfg.MODEL.PIXEL_MEAN), list(cfg.MODEL.PIXEL_STD)
            self._out_features = list(cfg.MODEL.RESNETS.OUT_FEATURES)
        else:
            self.pixel_mean, self.pixel_std = [123.675, 116.280, 103.530], [58.395, 57.120, 57.375]
            self._out_features = ["res2", "res3", "res4", "res5"]
        self.stem = _Stem()
        cin = 64
        for name, blocks, mid, cout, stride in STAGES:
            seq = nn.Sequential(*[_Bottleneck(cin if i == 0 else cout, mid, cout, stride if i == 0 else 1, i == 0) for i in range(blocks)])
            setattr(self, name, seq)
            cin = cout
        self._packed = None
        self.tc_passes = 3
        self.use_stage_entry = not os.environ.get("NSAC_PY_STAGES")      # nsac_backbone_forward; NSAC_PY_STAGES=1: same launches from Python
        self.eval()

    def output_shape(self) -> Dict[str, ShapeSpec]:
        full = {"res2": ShapeSpec(channels=256, stride=4), "res3": ShapeSpec(channels=512, stride=8),
                "res4": ShapeSpec(channels=1024, stride=16), "res5": ShapeSpec(channels=2048, stride=32)}
        return {k: full[k] for k in self._out_features}

    # ------------------------------------------------------------------ weight packing (invalidated like the heads')
    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def prepare(self):
        """FrozenBN folded into the weights, (ky,kx,cin)-ordered fp16 hi/lo planes + fp32 bias per convolution."""
        ver = ops.weights_version(self)
        if self._packed is None or self._packed_version != ver:
            self._packed_version = ver
            with torch.no_grad():
                def pack(cb: _ConvBN):
                    w, b = cb.folded()
                    return ops.split_weight(w), b          # planes are zero-padded to a multiple of 64 columns (stem: 147 -> 192)
                pk = {"stem": pack(self.stem.conv1)}
                # 8-bit image path: conv(w, (p - mean) / std) = conv(w / std, p) - sum w mean / std (interior pixels)
                wf, bf = self.stem.conv1.folded()                                   # [64, 147] in (ky, kx, c) order
                istd = (1.0 / torch.tensor(self.pixel_std, dtype=torch.float64, device=wf.device)).repeat(49)
                mean = torch.tensor(self.pixel_mean, dtype=torch.float64, device=wf.device).repeat(49)
                w8 = wf.double() * istd
                pk["stem.u8"] = (ops.split_weight(w8.float().contiguous()), (bf.double() - (w8 * mean).sum(1)).float().contiguous())
                pk["stem.f32"] = (wf.float().contiguous(), bf.float().contiguous())
                for name, *_ in STAGES:
                    for i, blk in enumerate(getattr(self, name)):
                        for c in ("conv1", "conv2", "conv3") + (("shortcut",) if hasattr(blk, "shortcut") else ()):
                            pk[f"{name}.{i}.{c}"] = pack(getattr(blk, c))
                self._packed = pk
        return self._packed

    # ------------------------------------------------------------------ forward
    def _block(self, pk, key, blk, xp, N, H, W):
        """One bottleneck block on planes: conv1 (1x1) -> conv2 (3x3, stride here) -> conv3 (1x1) + shortcut + ReLU in its epilogue."""
        P = self.tc_passes
        w1, b1 = pk[key + ".conv1"]
        w2, b2 = pk[key + ".conv2"]
        w3, b3 = pk[key + ".conv3"]
        Ho, Wo = H, W
        _, y1p = ops.gemm_tc(xp, w1, b1, ops.ACT_RELU, P, want_f32=False, want_split=True)
        # 3x3, stride 1 or 2 (STRIDE_IN_1X1 = False): implicit GEMM, the stride is the TMA tensor map's traversal stride
        _, y2p = ops.conv3x3_tc(y1p, N, H, W, w2, b2, ops.ACT_RELU, P, want_f32=False, want_split=True, stride=blk.stride)
        Ho, Wo = (H - 1) // blk.stride + 1, (W - 1) // blk.stride + 1
        if hasattr(blk, "shortcut"):
            ws, bs = pk[key + ".shortcut"]
            if blk.stride == 1:
                _, sc = ops.gemm_tc(xp, ws, bs, ops.ACT_NONE, P, want_f32=False, want_split=True)
            else:       # strided projection: the TMA gather skips the pixels a stride-2 1x1 convolution never reads
                sc = ops.conv1x1_tc_strided(xp, N, H, W, ws, bs, ops.ACT_NONE, P, stride=blk.stride)
        else:
            sc = xp
        _, out = ops.gemm_tc(y2p, w3, b3, ops.ACT_RELU, P, want_f32=False, want_split=True, residual=sc)
        return out, Ho, Wo

    def _stem_u8_weights(self, pk, H, W):
        """Stem weights for raw uint8 pixels of an H x W image: [64, 171] = w / std (147 taps) followed by the 24 border-class
        corrections sum_{taps outside the image} w * mean / std (fp64 on the host, then hi/lo planes), + the folded bias."""
        key = f"stem.u8.{H}x{W}"
        if key not in pk:
            with torch.no_grad():
                wf, bf = self.stem.conv1.folded()                                   # [64, 147] in (ky, kx, c) order
                dev = wf.device
                istd = (1.0 / torch.tensor(self.pixel_std, dtype=torch.float64, device=dev)).repeat(49)
                mean = torch.tensor(self.pixel_mean, dtype=torch.float64, device=dev).repeat(49)
                w8 = wf.double() * istd
                wm = (w8 * mean).view(64, 7, 7, 3)
                ext = torch.zeros(64, 171, dtype=torch.float64, device=dev)
                ext[:, :147] = w8
                for idx, oob in ops.stem_border_classes(H, W):
                    ext[:, 147 + idx] = (wm * oob.to(dev)[None, :, :, None]).sum(dim=(1, 2, 3))
                pk[key] = (ops.split_weight(ext.float().contiguous()), (bf.double() - (w8 * mean).sum(1)).float().contiguous())
        return pk[key]

    def backbone_weights(self, H: int, W: int):
        """`nsac_backbone_weights` for `ops.backbone_forward` on H x W uint8 images (the stem's border corrections depend on the
        image size; borrowed pointers into this weight version's packed planes)."""
        pk = self.prepare()
        key = f"struct.{H}x{W}"
        if key not in pk:
            from . import _lib
            blocks = []
            for name, *_ in STAGES:
                for i, blk in enumerate(getattr(self, name)):
                    b = _lib.Bottleneck()
                    for c in ("conv1", "conv2", "conv3"):
                        setattr(b, c, ops.tc_layer(*pk[f"{name}.{i}.{c}"]))
                    b.has_shortcut = 1 if hasattr(blk, "shortcut") else 0
                    if b.has_shortcut:
                        b.shortcut = ops.tc_layer(*pk[f"{name}.{i}.shortcut"])
                    b.stride = int(blk.stride)
                    blocks.append(b)
            arr = (_lib.Bottleneck * len(blocks))(*blocks)
            Wt = _lib.BackboneWeights()
            Wt.stem = ops.tc_layer(*self._stem_u8_weights(pk, H, W))
            Wt.blocks, Wt.num_blocks = arr, len(blocks)
            for j, (name, *_r) in enumerate(STAGES):
                Wt.stage_blocks[j] = len(getattr(self, name))
            Wt.fmt, Wt.passes = ops.SPLIT_F16, self.tc_passes
            pk[key], pk[key + ".keep"] = Wt, arr
        return pk[key]

    def _stem(self, pk, images):
        N = images.shape[0]
        if images.dtype == torch.uint8 and min(images.shape[2:]) >= 9:
            # raw pixels are exact in fp16: one plane, normalisation folded into the weights.  The reference zero-pads the
            # NORMALISED image, i.e. an out-of-image tap contributes 0 instead of w * (0 - mean) / std: 24 one-hot border-class
            # columns of the im2col matrix select the matching correction row of the weight matrix (_stem_u8_weights)
            cols, H, W = ops.stem_im2col_u8(images, border_classes=True)
            ws, bs = self._stem_u8_weights(pk, images.shape[2], images.shape[3])
            x, _ = ops.gemm_tc(cols, ws, bs, ops.ACT_RELU, self.tc_passes)
        elif images.dtype == torch.uint8:
            # tiny images (border classes overlap): plain im2col + exact fp32 recomputation of the border pixels
            cols, H, W = ops.stem_im2col_u8(images)
            ws, bs = pk["stem.u8"]
            x, _ = ops.gemm_tc(cols, ws, bs, ops.ACT_RELU, self.tc_passes)
            wf, bf = pk["stem.f32"]
            ops.stem_border_fix(images, wf, bf, self.pixel_mean, self.pixel_std, x)
        else:
            cols, H, W = ops.stem_im2col_planes(images.float(), self.pixel_mean, self.pixel_std)
            ws, bs = pk["stem"]
            x, _ = ops.gemm_tc(cols, ws, bs, ops.ACT_RELU, self.tc_passes)
        _, xp, H, W = ops.maxpool3x3s2_nhwc(x, N, H, W, want_f32=False)
        return xp, H, W

    @torch.no_grad()
    def forward(self, images: torch.Tensor, nhwc: bool = False, planes: bool = False):
        """images: [N,3,H,W], NOT normalised (0..255, cfg.INPUT.FORMAT order), uint8 (fast stem) or float ->
        {'res2'..'res5'}: [N,C,h,w] fp32 like the reference's backbone; `nhwc=True`: the NHWC rows [N*h*w, C] fp32;
        `planes=True`: a `PlaneFeatures` (NHWC hi/lo planes, the engine's own format) for `PlaneCameraHead`."""
        pk = self.prepare()
        N = images.shape[0]
        if self.use_stage_entry and images.dtype == torch.uint8 and min(images.shape[2:]) >= 9:
            # the whole backbone behind ONE C call (nsac_backbone_forward, csrc/forward.cu); the loop below issues the same launches
            Wt = self.backbone_weights(images.shape[2], images.shape[3])
            Wt.passes = self.tc_passes
            lv = ops.backbone_forward(Wt, images, keep=self._out_features)
            out = PlaneFeatures(N) if planes else {}
            for name, (sp, h, w) in lv.items():
                if planes:
                    out[name] = (sp, h, w)
                else:
                    x_f32 = sp.float()
                    out[name] = x_f32 if nhwc else x_f32.view(N, h, w, -1).permute(0, 3, 1, 2).contiguous()
            return out
        xp, H, W = self._stem(pk, images)
        out = PlaneFeatures(N) if planes else {}
        for name, *_ in STAGES:
            for i, blk in enumerate(getattr(self, name)):
                xp, H, W = self._block(pk, f"{name}.{i}", blk, xp, N, H, W)
            if name in self._out_features:
                if planes:
                    out[name] = (xp, H, W)
                else:
                    x_f32 = xp.float()
                    out[name] = x_f32 if nhwc else x_f32.view(N, H, W, -1).permute(0, 3, 1, 2).contiguous()
        return out


class PlaneFeatures(dict):
    """Backbone output in the tensor-core engine's own format: name -> (ops.Split NHWC planes [N*h*w, C], h, w) for a batch of
    `N` images.  `PlaneCameraHead` takes it as `features1` (with `features2=None`): images [0, N/2) are the first views,
    [N/2, N) the second views — exactly the stacking its pixel network builds from two NCHW dicts."""

    def __init__(self, num_images: int):
        super().__init__()
        self.num_images = num_images


class RawImages:
    """The uint8 images of both views of B pairs (stacked [2B,3,H,W]: first views, then second views) together with the backbone
    that turns them into features.  `PlaneCameraHead` takes it as `features1` (with `features2=None`): when everything else allows,
    backbone AND head then run behind ONE C call (`nsac_model_forward`); otherwise `.features()` runs the backbone first."""

    def __init__(self, backbone: "ResNet50Backbone", images: torch.Tensor):
        self.backbone, self.images, self.num_images = backbone, images, images.shape[0]

    @property
    def single_call_ok(self) -> bool:
        im = self.images
        return bool(self.backbone.use_stage_entry and im.dtype == torch.uint8 and im.shape[2] % 32 == 0 and im.shape[3] % 32 == 0
                    and im.shape[2] >= 32 and im.shape[3] >= 32)

    def features(self) -> "PlaneFeatures":
        return self.backbone(self.images, planes=True)


@BACKBONE_REGISTRY.register()
def build_resnet_backbone(cfg, input_shape=None):
    return ResNet50Backbone(cfg, input_shape)


def build_backbone(cfg, input_shape=None):
    """detectron2.modeling.build_backbone: BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, input_shape)."""
    return BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, input_shape)
