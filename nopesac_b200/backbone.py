"""ResNet-50 backbone on the tensor-core engine — SURVEY.md §8 row f2 (the producer of the camera head's `res3..res5` inputs).

Mirrors what the reference gets from detectron2 (`cfg.MODEL.BACKBONE.NAME = "build_resnet_backbone"`, configs/Base.yaml:2-12:
DEPTH 50, STEM_OUT_CHANNELS 64, STRIDE_IN_1X1 False, OUT_FEATURES res2..res5, FrozenBN, weights = torchvision's R-50 renamed):
same registry name, same parameter / buffer names (`stem.conv1.weight`, `stem.conv1.norm.running_mean`,
`res2.0.shortcut.weight`, `res4.5.conv3.norm.bias`, ...), same `output_shape()`, and `forward` fuses the
`(x - PIXEL_MEAN) / PIXEL_STD` of `preprocess_image` (siamese_planeTR.py `preprocess_image`, Base.yaml:6-7).

Every convolution is a GEMM on NHWC 16-bit hi/lo planes (3 passes ~ fp32): 1x1 -> `nsac_gemm_split`, 3x3 stride 1 ->
`nsac_conv3x3_split` (implicit GEMM, 4-D TMA gather), 3x3 stride 2 -> `nsac_im2col3x3_planes` + GEMM, stem 7x7/2 ->
`nsac_stem_im2col_planes` + GEMM; FrozenBN is folded into weights and bias, ReLU runs in the GEMM epilogue; max-pool,
stride-2 subsampling of the shortcut input and relu(out + shortcut) are the byte movers of csrc/backbone.cu.  No cuDNN, no
CPU path.  First version: the residual add is a separate kernel (fusing it into the GEMM epilogue is the next step).
"""
from __future__ import annotations

from typing import Dict

import torch
from torch import nn

from . import ops
from .compat import Registry, ShapeSpec

__all__ = ["BACKBONE_REGISTRY", "ResNet50Backbone", "build_resnet_backbone", "build_backbone"]

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
                for name, *_ in STAGES:
                    for i, blk in enumerate(getattr(self, name)):
                        for c in ("conv1", "conv2", "conv3") + (("shortcut",) if hasattr(blk, "shortcut") else ()):
                            pk[f"{name}.{i}.{c}"] = pack(getattr(blk, c))
                self._packed = pk
        return self._packed

    # ------------------------------------------------------------------ forward
    def _block(self, pk, key, blk, x_f32, xp, N, H, W):
        P = self.tc_passes
        w1, b1 = pk[key + ".conv1"]
        w2, b2 = pk[key + ".conv2"]
        w3, b3 = pk[key + ".conv3"]
        Ho, Wo = H, W
        if blk.stride == 1:
            _, y1p = ops.gemm_tc(xp, w1, b1, ops.ACT_RELU, P, want_f32=False, want_split=True)
            _, y2p = ops.conv3x3_tc(y1p, N, H, W, w2, b2, ops.ACT_RELU, P, want_f32=False, want_split=True)
        else:
            y1, _ = ops.gemm_tc(xp, w1, b1, ops.ACT_RELU, P)
            cols, Ho, Wo = ops.im2col3x3_planes(y1, N, H, W, blk.stride)
            _, y2p = ops.gemm_tc(cols, w2, b2, ops.ACT_RELU, P, want_f32=False, want_split=True)
        y3, _ = ops.gemm_tc(y2p, w3, b3, ops.ACT_NONE, P)
        if hasattr(blk, "shortcut"):
            ws, bs = pk[key + ".shortcut"]
            src = xp if blk.stride == 1 else ops.subsample2_planes(xp, N, H, W)[0]
            sc, _ = ops.gemm_tc(src, ws, bs, ops.ACT_NONE, P)
        else:
            sc = x_f32
        x_f32, xp = ops.add_relu_nhwc(y3, sc)
        return x_f32, xp, Ho, Wo

    @torch.no_grad()
    def forward(self, images: torch.Tensor, nhwc: bool = False) -> Dict[str, torch.Tensor]:
        """images: [N,3,H,W] fp32, NOT normalised (0..255, cfg.INPUT.FORMAT order) -> {'res2'..'res5'}: [N,C,h,w] fp32 (or the
        NHWC rows [N*h*w, C] the kernels produce, with `nhwc=True`)."""
        pk = self.prepare()
        N = images.shape[0]
        cols, H, W = ops.stem_im2col_planes(images.float(), self.pixel_mean, self.pixel_std)
        ws, bs = pk["stem"]
        x, _ = ops.gemm_tc(cols, ws, bs, ops.ACT_RELU, self.tc_passes)
        x_f32, xp, H, W = ops.maxpool3x3s2_nhwc(x, N, H, W)
        out = {}
        for name, *_ in STAGES:
            for i, blk in enumerate(getattr(self, name)):
                x_f32, xp, H, W = self._block(pk, f"{name}.{i}", blk, x_f32, xp, N, H, W)
            if name in self._out_features:
                out[name] = x_f32 if nhwc else x_f32.view(N, H, W, -1).permute(0, 3, 1, 2).contiguous()
        return out


@BACKBONE_REGISTRY.register()
def build_resnet_backbone(cfg, input_shape=None):
    return ResNet50Backbone(cfg, input_shape)


def build_backbone(cfg, input_shape=None):
    """detectron2.modeling.build_backbone: BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, input_shape)."""
    return BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, input_shape)
