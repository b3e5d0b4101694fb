"""CAMERA_HEAD of the path — same registry surface, constructor, call signature, 6-tuple return value and
state-dict names as the reference `PlaneCameraHead` (camera_net/camera_head.py:21-138, 140-149, 400-640),
computed by the CUDA kernels of libnopesac_b200 and *batched*: every `[0]`-indexed shortcut of the
reference (quaternion sign flips :436,:600; early exits on matched_nums[0] :964,:1052,:1068) is applied
per pair inside the kernels, with no host synchronisation anywhere on the path.

    CAMERA_HEAD_REGISTRY.get(cfg.MODEL.CAMERA_HEAD.NAME)(cfg, input_shape)
    output_cameras, trans_list, rot_list, [log_scores_padded], output_planeAss, pose_ref_outputs = head(
        features1, features2, planeParam1, planeParam2, planeApp1, planeApp2, matching_net=matching_head)

Inference only.  Extra keyword arguments (not in the reference):
    hyp_pairs     int32 [H,2]  explicit hypothesis list (the "P planes x H hypotheses" stress mapping,
                               SURVEY.md §8(d)); the matcher still runs and its assignment is reported.
    assignment_override [B,n1,n2]  0/1 matrix whose row-major nonzeros replace the matcher's assignment as
                               the hypothesis list of each pair (what POSE_REFINEMENT_WITH_GT_MATCHERS does in
                               the reference, camera_head.py:520-547, minus the dataset lookup).
    initial_pose  (t [B,3], q [B,4])  skips the pixel pose network (stage set S3).
    result_exchange  nopesac_b200.dist.FusedResultExchange: the selection kernel also stores every [16]-float result
                               row into all ranks' result buffers over NVLink (multi-GPU; replaces the all-gather).
"""
from __future__ import annotations

import os

from typing import Dict, Optional

import torch
from torch import nn

from . import ops
from .compat import Registry, ShapeSpec

__all__ = ["build_camera_head", "CAMERA_HEAD_REGISTRY", "PlaneCameraHead"]

CAMERA_HEAD_REGISTRY = Registry("CAMERA_HEAD")
CAMERA_HEAD_REGISTRY.__doc__ = "Registry for camera head. The call is expected to return an nn.Module."


def build_camera_head(cfg, input_shape):
    """camera_head.py:27-32."""
    return CAMERA_HEAD_REGISTRY.get(cfg.MODEL.CAMERA_HEAD.NAME)(cfg, input_shape)


# ---------------------------------------------------------------------------------------------------
# parameter containers (names = reference state-dict keys; initialisers as in the reference)
# ---------------------------------------------------------------------------------------------------
def _c2_xavier_fill(m):
    nn.init.kaiming_uniform_(m.weight, a=1)
    if m.bias is not None:
        nn.init.constant_(m.bias, 0)


def _c2_msra_fill(m):
    nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
    if m.bias is not None:
        nn.init.constant_(m.bias, 0)


class MLP(nn.Module):
    """camera_modules.py:226-244 (parameters only; evaluated through ops.linear)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))
        for layer in self.layers:
            _c2_xavier_fill(layer)

    def run(self, x: torch.Tensor, final_act: int = ops.ACT_NONE, out: Optional[torch.Tensor] = None):
        """Exact-fp32 CUDA-core path (small batches / K not a multiple of 64)."""
        for i, layer in enumerate(self.layers):
            last = i == self.num_layers - 1
            x = ops.linear(x, layer.weight, layer.bias, ops.ACT_RELU if not last else final_act,
                           out=out if last else None)
        return x

    def split_weights(self):
        """hi/lo planes of every layer whose K is a multiple of 64 (None otherwise)."""
        return [ops.split_weight(l.weight) if l.weight.shape[1] % 64 == 0 else None for l in self.layers]

    def run_tc(self, x, wsplit, final_act: int = ops.ACT_NONE, out_split: Optional["ops.Split"] = None,
               want_f32: bool = False, first_bias=None, first_group_rows: int = 0, first_weight=None, passes: int = 3):
        """tcgen05 split-bf16 path.  `x` is an ops.Split, or an fp32 matrix when layer 0 has a small K (that
        layer then runs on the CUDA cores and is re-split).  Returns (fp32 or None, Split)."""
        f32 = None
        for i, layer in enumerate(self.layers):
            last = i == self.num_layers - 1
            act = ops.ACT_RELU if not last else final_act
            w = wsplit[i] if not (i == 0 and first_weight is not None) else first_weight
            bias = layer.bias if not (i == 0 and first_bias is not None) else first_bias
            grp = first_group_rows if i == 0 else 0
            if w is None:
                assert i == 0 and isinstance(x, torch.Tensor)
                x = ops.split(ops.linear(x, layer.weight, bias, act))
                continue
            f32, x = ops.gemm_tc(x, w, bias, act, passes=passes, want_f32=want_f32 and last, want_split=True,
                                 out_split=out_split if last else None, bias_group_rows=grp)
        return f32, x


class _NormConv(nn.Conv2d):
    """detectron2.layers.Conv2d parameter layout: conv weight (+bias) and an optional `.norm` child."""

    def __init__(self, cin, cout, k, padding=0, norm: Optional[nn.Module] = None, bias=True):
        super().__init__(cin, cout, kernel_size=k, stride=1, padding=padding, bias=bias)
        if norm is not None:
            self.norm = norm


def conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=None):
    """camera_modules.py:36-48."""
    return nn.Sequential(
        nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding, bias=False),
        nn.BatchNorm2d(out_channels, eps=0.001, momentum=0.01),
        nn.LeakyReLU(inplace=True),
    )


class BasePixelDecoder(nn.Module):
    """camera_modules.py:246-333 (res2 dropped; GroupNorm(32) when NORM == "GN")."""

    def __init__(self, cfg, input_shape: Dict[str, ShapeSpec]):
        super().__init__()
        shapes = {k: v for k, v in input_shape.items() if k in cfg.MODEL.SEM_SEG_HEAD.IN_FEATURES}
        items = sorted(shapes.items(), key=lambda x: x[1].stride)[1:]
        self.in_features = [k for k, _ in items]
        conv_dim, mask_dim = cfg.MODEL.SEM_SEG_HEAD.CONVS_DIM, cfg.MODEL.SEM_SEG_HEAD.MASK_DIM
        norm = cfg.MODEL.SEM_SEG_HEAD.NORM
        assert norm in ("GN", ""), f"SEM_SEG_HEAD.NORM={norm!r} not supported"
        mk = (lambda: nn.GroupNorm(32, conv_dim)) if norm == "GN" else (lambda: None)
        use_bias = norm == ""
        for idx, (_, spec) in enumerate(items):
            if idx == len(items) - 1:
                oc = _NormConv(spec.channels, conv_dim, 3, 1, mk(), use_bias)
                _c2_xavier_fill(oc)
                self.add_module(f"layer_{idx + 1}", oc)
            else:
                lc = _NormConv(spec.channels, conv_dim, 1, 0, mk(), use_bias)
                oc = _NormConv(conv_dim, conv_dim, 3, 1, mk(), use_bias)
                _c2_xavier_fill(lc)
                _c2_xavier_fill(oc)
                self.add_module(f"adapter_{idx + 1}", lc)
                self.add_module(f"layer_{idx + 1}", oc)
        self.mask_dim = mask_dim
        self.mask_features = _NormConv(conv_dim, mask_dim, 3, 1)
        _c2_xavier_fill(self.mask_features)


# ---------------------------------------------------------------------------------------------------
@CAMERA_HEAD_REGISTRY.register()
class PlaneCameraHead(nn.Module):
    def __init__(self, cfg, input_shape):
        super().__init__()
        self.cfg = cfg
        self.plane_matcher_on = cfg.MODEL.EMBEDDING_ON and cfg.MODEL.MASK_ON
        self.rand_cam_on = cfg.MODEL.CAMERA_HEAD.RAND_ON
        self.cam_rec_on = cfg.MODEL.CAMERA_HEAD.CAM_REC_ON
        self.cam_ref_on = cfg.MODEL.CAMERA_HEAD.REFINE_ON
        self.use_sparsePlane_Top1Cam_testSet = cfg.MODEL.CAMERA_HEAD.INFERENCE_SP_TOPCAM_ON
        self.num_queries = cfg.MODEL.SEM_SEG_HEAD.NUM_OBJECT_QUERIES
        self.inference_out_cam_type = cfg.MODEL.CAMERA_HEAD.INFERENCE_OUT_CAM_TYPE
        self.matching_score_threshold = cfg.TEST.MATCHING_SCORE_THRESHOLD
        self.warp_plane_in_cam_ref_on = cfg.MODEL.CAMERA_HEAD.WARP_PLANE_IN_CAM_REF_ON
        if self.use_sparsePlane_Top1Cam_testSet:
            raise NotImplementedError("INFERENCE_SP_TOPCAM_ON (cached SparsePlanes top-1 camera) is out of scope")
        if not self.warp_plane_in_cam_ref_on:
            raise NotImplementedError("WARP_PLANE_IN_CAM_REF_ON=False is not used by any reference config")
        if cfg.TEST.POSE_REFINEMENT_WITH_GT_MATCHERS:
            raise NotImplementedError("GT-matcher ablation (camera_head.py:520-547) is out of scope")

        # pixel camera head (camera_head.py:76-114)
        self.pixel_decoder = BasePixelDecoder(cfg, input_shape)
        self.convs_backbone = nn.Sequential(
            conv2d(256, 256, 3, padding=1), conv2d(256, 256, 3, padding=1), nn.MaxPool2d(2, 2),
            conv2d(256, 256, 3, padding=1), conv2d(256, 256, 3, padding=1), nn.MaxPool2d(2, 2),
            conv2d(256, 256, 3, padding=1), conv2d(256, 256, 3, padding=1))
        for block in self.convs_backbone:
            if isinstance(block, nn.Sequential):
                _c2_msra_fill(block[0])
        strides = (1, 2, 1, 2, 1, 2)
        self.convs_trans = nn.Sequential(*[conv2d(300 if i == 0 else 128, 128, 3, stride=s, padding=1)
                                           for i, s in enumerate(strides)])
        self.convs_rots = nn.Sequential(*[conv2d(300 if i == 0 else 128, 128, 3, stride=s, padding=1)
                                          for i, s in enumerate(strides)])
        self.fc_trans = nn.Linear(768, 256)
        self.fc_rots = nn.Linear(768, 256)
        # shared pose regressors (:64-65)
        self.trans = nn.Linear(256, 3)
        self.rots = nn.Linear(256, 4)
        if self.cam_rec_on:       # AIM (:116-120)
            self.rot_emb_proj = MLP(4, 256, 256, 6)
            self.trans_emb_proj = MLP(3, 256, 256, 6)
        if self.cam_ref_on:       # NOPE-SAC refinement (:122-138)
            self.geo_encoder = MLP(8, 1024, 1024, 6)
            self.geo_proj_s1 = MLP(1024, 1024, 1024, 3)
            self.decoder_rot = MLP(1024, 512, 256, 6)
            self.geo_proj_s2 = MLP(1024 + 256, 1024, 1024, 3)
            self.decoder_tran = MLP(1024, 512, 256, 6)
            self.decoder_rot2 = MLP(512, 512, 256, 3)
            self.decoder_tran2 = MLP(512, 512, 256, 3)
            self.normal_score_proj = MLP(self.num_queries, 128, 64, 3)
            self.rot_score_reg = nn.Linear(64, 1)
            self.param_score_proj = MLP(self.num_queries, 128, 64, 3)
            self.trans_score_reg = nn.Linear(64, 1)
        self._packed = None
        # K6 .. K10 through the single C entry nsac_refine_forward (False / NSAC_PY_STAGES=1: the same launches from Python)
        self.use_stage_entry = not os.environ.get("NSAC_PY_STAGES")
        self.tc_passes = 3     # MMA passes of the tensor-core layers (3 = hi.hi + lo.hi + hi.lo on fp16 planes, ~fp32)

    # ------------------------------------------------------------------ weight packing
    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def prepare(self):
        ver = ops.weights_version(self)
        if self._packed is None or self._packed_version != ver:
            self._packed_version = ver
            with torch.no_grad():
                pk = {}
                if self.cam_ref_on:
                    for name in ("decoder_rot2", "decoder_tran2"):
                        w = getattr(self, name).layers[0].weight
                        pk[name + ".w_init"] = w[:, :256].contiguous()   # half applied to the initial-pose feature
                        pk[name + ".w_geo"] = w[:, 256:].contiguous()    # half applied to the one-plane feature
                    for name, reg in (("normal_score_proj", "rot_score_reg"), ("param_score_proj", "trans_score_reg")):
                        m = getattr(self, name)
                        r = getattr(self, reg)
                        pk[name] = tuple(t.detach().contiguous() for t in (
                            m.layers[0].weight, m.layers[0].bias, m.layers[1].weight, m.layers[1].bias,
                            m.layers[2].weight, m.layers[2].bias, r.weight, r.bias))
                self._packed = pk
        return self._packed

    def prepare_tc(self):
        """hi/lo weight planes for the tcgen05 engine (device-side; built once per weight version)."""
        pk = self.prepare()
        if "geo_encoder.split" not in pk:
            with torch.no_grad():
                for name in ("geo_encoder", "geo_proj_s1", "decoder_rot", "geo_proj_s2", "decoder_tran",
                             "decoder_rot2", "decoder_tran2", "rot_emb_proj", "trans_emb_proj"):
                    pk[name + ".split"] = getattr(self, name).split_weights()
                for name in ("decoder_rot2", "decoder_tran2"):
                    pk[name + ".w_geo_split"] = ops.split_weight(pk[name + ".w_geo"])
                pk["score_pack"] = ops.score_pack(pk["normal_score_proj"], pk["param_score_proj"], self.num_queries)
                pk["score_vecs_host"] = ops.score_pack_host_vectors(pk["score_pack"], self.num_queries)
                self._prepare_pixel_tc(pk)
        return pk

    # ------------------------------------------------------------------ K1 weights for the tensor-core engine
    @staticmethod
    def _conv_planes(w, row_scale=None, cin_pad=None):
        """conv weight [Cout,Cin,3,3] -> planes [Cout, 9*Cin'] in (ky, kx, cin) order (Cin zero-padded to cin_pad)."""
        w = w.detach().permute(0, 2, 3, 1)
        if cin_pad is not None and cin_pad > w.shape[-1]:
            w = torch.nn.functional.pad(w, (0, cin_pad - w.shape[-1]))
        if row_scale is not None:
            w = w * row_scale[:, None, None, None]
        return ops.split_weight(w.reshape(w.shape[0], -1).contiguous())

    @staticmethod
    def _bn_fold(block):
        """Conv-BN(eval)-LeakyReLU block (camera_modules.py:36-48): y = conv(x) * s + (beta - mean * s)."""
        bn = block[1]
        s = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
        return s, (bn.bias.detach() - bn.running_mean * s).contiguous()

    def _prepare_pixel_tc(self, pk):
        pd = self.pixel_decoder
        for name in ("layer_1", "layer_2", "layer_3", "mask_features"):
            pk[f"pd.{name}.w"] = self._conv_planes(getattr(pd, name).weight)
        for name in ("adapter_1", "adapter_2"):
            w = getattr(pd, name).weight.detach()
            pk[f"pd.{name}.w"] = ops.split_weight(w.reshape(w.shape[0], w.shape[1]).contiguous())
        for i in (0, 1, 3, 4, 6, 7):
            s, b = self._bn_fold(self.convs_backbone[i])
            pk[f"cb.{i}.w"], pk[f"cb.{i}.b"] = self._conv_planes(self.convs_backbone[i][0].weight, s), b
        # first conv of both correlation branches shares its input: one GEMM with N = 256, Cin 300 -> 320
        ws, bs = [], []
        for convs in (self.convs_trans, self.convs_rots):
            s, b = self._bn_fold(convs[0])
            w = convs[0][0].weight.detach().permute(0, 2, 3, 1)
            w = torch.nn.functional.pad(w, (0, 320 - w.shape[-1])) * s[:, None, None, None]
            ws.append(w.reshape(w.shape[0], -1))
            bs.append(b)
        pk["ct0.w"], pk["ct0.b"] = ops.split_weight(torch.cat(ws, 0).contiguous()), torch.cat(bs).contiguous()
        for name, convs, fc in (("trans", self.convs_trans, self.fc_trans), ("rots", self.convs_rots, self.fc_rots)):
            for i in range(1, 6):
                s, b = self._bn_fold(convs[i])
                pk[f"convs_{name}.{i}.w"], pk[f"convs_{name}.{i}.b"] = self._conv_planes(convs[i][0].weight, s), b
            # reference flattens [B,128,2,3] channel-major (c*6 + h*3 + w); NHWC rows flatten as (h*3 + w)*128 + c
            w = fc.weight.detach()
            pk[f"fc_{name}.w_nhwc"] = w.view(w.shape[0], 128, -1).permute(0, 2, 1).reshape(w.shape[0], -1).contiguous()

    # ------------------------------------------------------------------ K1: pixel pose network
    def _forward_pixel_camera_head(self, features1, features2):
        """camera_head.py:642-670 on the tensor-core engine: NHWC activations, every convolution an (implicit) GEMM
        (nsac_conv3x3_split / nsac_gemm_split, fp16 hi/lo planes, 3 passes), BatchNorm folded into the weights,
        GroupNorm / nearest-upsample-add / max-pool / correlation-softmax as small fused kernels.  Both views are
        stacked along the batch dimension (N = 2B images)."""
        pk = self.prepare_tc()
        pd = self.pixel_decoder
        if not hasattr(pd.layer_3, "norm"):
            raise NotImplementedError("SEM_SEG_HEAD.NORM must be 'GN' (as in every reference config)")
        P = self.tc_passes
        from .backbone import PlaneFeatures
        stacked = isinstance(features1, PlaneFeatures)      # both views already stacked as NHWC planes (backbone.forward(planes=True))
        B = features1.num_images // 2 if stacked else features1["res5"].shape[0]
        N = 2 * B
        LEAKY = ops.ACT_LEAKY

        def planes(name):
            if stacked:
                return features1[name]
            f1, f2 = features1[name], features2[name]
            _, C, H, W = f1.shape
            out = ops.Split(torch.empty(N * H * W, C, device=f1.device, dtype=torch.float16),
                            torch.empty(N * H * W, C, device=f1.device, dtype=torch.float16), C)
            ops.nchw_to_planes(f1.float(), out=out, row_offset=0)
            ops.nchw_to_planes(f2.float(), out=out, row_offset=B * H * W)
            return out, H, W

        def gn(x, H, W, m, relu, skip=None, f32=True, split=False):
            return ops.groupnorm_nhwc(x, N, H, W, m.norm.weight, m.norm.bias, m.norm.num_groups, m.norm.eps, relu, skip,
                                      want_f32=f32, want_split=split)

        # --- BasePixelDecoder.forward_features (camera_modules.py:335-348): res5 -> res4 -> res3, top-down
        p5, H5, W5 = planes("res5")
        t, _ = ops.conv3x3_tc(p5, N, H5, W5, pk["pd.layer_3.w"], passes=P)
        y5, _ = gn(t, H5, W5, pd.layer_3, True)
        p4, H4, W4 = planes("res4")
        t, _ = ops.gemm_tc(p4, pk["pd.adapter_2.w"], passes=P)
        _, u4 = gn(t, H4, W4, pd.adapter_2, False, skip=y5, f32=False, split=True)
        t, _ = ops.conv3x3_tc(u4, N, H4, W4, pk["pd.layer_2.w"], passes=P)
        y4, _ = gn(t, H4, W4, pd.layer_2, True)
        p3, H3, W3 = planes("res3")
        t, _ = ops.gemm_tc(p3, pk["pd.adapter_1.w"], passes=P)
        _, u3 = gn(t, H3, W3, pd.adapter_1, False, skip=y4, f32=False, split=True)
        t, _ = ops.conv3x3_tc(u3, N, H3, W3, pk["pd.layer_1.w"], passes=P)
        _, y3 = gn(t, H3, W3, pd.layer_1, True, f32=False, split=True)
        _, x = ops.conv3x3_tc(y3, N, H3, W3, pk["pd.mask_features.w"], pd.mask_features.bias, passes=P,
                              want_f32=False, want_split=True)
        # --- convs_backbone (camera_head.py:78-91): conv-BN-LeakyReLU x2, pool, x2, pool, x2
        H, W = H3, W3
        for i in (0, 3, 6):
            _, x = ops.conv3x3_tc(x, N, H, W, pk[f"cb.{i}.w"], pk[f"cb.{i}.b"], LEAKY, P, want_f32=False, want_split=True)
            f, _ = ops.conv3x3_tc(x, N, H, W, pk[f"cb.{i + 1}.w"], pk[f"cb.{i + 1}.b"], LEAKY, P)
            if i < 6:
                x = ops.maxpool2_planes(f, N, H, W)
                H, W = H // 2, W // 2
        # --- correlation volume + softmax (:652, :1117-1133), then the two regression branches (:655-662)
        HW = H * W
        aff = ops.corr_softmax(f[:B * HW], f[B * HW:], B, H, W)
        t, _ = ops.conv3x3_tc(aff, B, H, W, pk["ct0.w"], pk["ct0.b"], LEAKY, P)          # [B*HW, 256] = trans | rots
        feats = []
        for name, off, fc in (("trans", 0, self.fc_trans), ("rots", 128, self.fc_rots)):
            y, h, w = t[:, off:off + 128].contiguous(), H, W
            for i in range(1, 6):
                stride = 2 if i % 2 == 1 else 1
                cols, h, w = ops.im2col3x3_planes(y, B, h, w, stride)
                y, _ = ops.gemm_tc(cols, pk[f"convs_{name}.{i}.w"], pk[f"convs_{name}.{i}.b"], LEAKY, passes=P)
            feats.append(ops.linear(y.view(B, -1), pk[f"fc_{name}.w_nhwc"], fc.bias, ops.ACT_RELU))
        trans_feat, rots_feat = feats
        rot, tran = ops.pose_heads(rots_feat, trans_feat, self.rots.weight, self.rots.bias,
                                   self.trans.weight, self.trans.bias)
        return tran, rot, trans_feat, rots_feat

    # ------------------------------------------------------------------ K2: AIM (:685-735)
    def _forward_rec_heads(self, initial_rot, initial_trans):
        # AIM (:685-735): layer 0 (K = 4 / 3) on the CUDA cores, the five 256-wide layers on the tcgen05 engine (M = batch
        # size is one tile: 7 us per layer instead of 35 us on the 128x128-tile fp32 kernel, which runs them on two CTAs)
        pk = self.prepare_tc()
        P = self.tc_passes
        rot_feat, _ = self.rot_emb_proj.run_tc(initial_rot, pk["rot_emb_proj.split"], final_act=ops.ACT_RELU, want_f32=True, passes=P)
        trans_feat, _ = self.trans_emb_proj.run_tc(initial_trans + 1e-10, pk["trans_emb_proj.split"], final_act=ops.ACT_RELU,
                                                   want_f32=True, passes=P)
        rec_rot, rec_trans = ops.pose_heads(rot_feat, trans_feat, self.rots.weight, self.rots.bias,
                                            self.trans.weight, self.trans.bias)
        return rec_rot, rot_feat, rec_trans, trans_feat

    # ------------------------------------------------------------------ K7: hypothesis features (:957-986)
    def _hypothesis_features(self, geo8, rot_feat0, trans_feat0, B, NQ):
        """The ~28-layer one-plane pose MLP chain on the tcgen05 split-precision engine (fp16 hi/lo planes, 3 passes: ~fp32).
        Activations stay as 16-bit hi/lo planes between layers; cat[s1, rot] (:961) is a column slice of one
        1280-wide plane pair, cat[init_feat, geo_feat] (:983-986) becomes a per-pair bias."""
        pk = self.prepare_tc()
        rows = B * NQ
        dev = geo8.device
        P = self.tc_passes
        _, fea = self.geo_encoder.run_tc(geo8.view(rows, 8), pk["geo_encoder.split"], passes=P)
        cat = ops.Split.empty(rows, 1280, dev)
        self.geo_proj_s1.run_tc(fea, pk["geo_proj_s1.split"], out_split=cat.cols(0, 1024), passes=P)
        self.decoder_rot.run_tc(cat.cols(0, 1024), pk["decoder_rot.split"], out_split=cat.cols(1024, 1280), passes=P)
        _, s2 = self.geo_proj_s2.run_tc(cat, pk["geo_proj_s2.split"], passes=P)
        _, ftran = self.decoder_tran.run_tc(s2, pk["decoder_tran.split"], passes=P)
        fused = []
        for name, feat0, geo_feat in (("decoder_rot2", rot_feat0, cat.cols(1024, 1280)), ("decoder_tran2", trans_feat0, ftran)):
            m = getattr(self, name)
            # cat[init_feat (broadcast over the pair's rows), geo_feat] @ W^T  ==  geo_feat @ W_geo^T + per-pair bias
            gb = ops.linear(feat0, pk[name + ".w_init"], m.layers[0].bias)
            f32, _ = m.run_tc(geo_feat, pk[name + ".split"], final_act=ops.ACT_RELU, want_f32=True, first_bias=gb,
                              first_group_rows=NQ, first_weight=pk[name + ".w_geo_split"], passes=P)
            fused.append(f32)                                                     # F.relu(decoder_*2(.))
        return fused[0], fused[1]

    def pixel_weights(self):
        """`nsac_pixel_weights` for `ops.pixel_forward` (K1 + K2; borrowed pointers into this weight version's packed planes)."""
        pk = self.prepare_tc()
        if "pixel_struct" not in pk:
            from . import _lib
            W = _lib.PixelWeights()
            pd = self.pixel_decoder
            if not hasattr(pd.layer_3, "norm"):
                raise NotImplementedError("SEM_SEG_HEAD.NORM must be 'GN' (as in every reference config)")
            tl = ops.tc_layer
            W.pd_layer_3, W.pd_layer_2, W.pd_layer_1 = (tl(pk[f"pd.{k}.w"], None) for k in ("layer_3", "layer_2", "layer_1"))
            W.pd_mask_features = tl(pk["pd.mask_features.w"], pd.mask_features.bias)
            W.pd_adapter_2, W.pd_adapter_1 = tl(pk["pd.adapter_2.w"], None), tl(pk["pd.adapter_1.w"], None)
            norms = [getattr(pd, k).norm for k in ("layer_3", "adapter_2", "layer_2", "adapter_1", "layer_1")]
            assert len({(m.num_groups, m.eps) for m in norms}) == 1
            for i, m in enumerate(norms):
                W.gn_w[i], W.gn_b[i] = m.weight.data_ptr(), m.bias.data_ptr()
            W.gn_groups, W.gn_eps = norms[0].num_groups, float(norms[0].eps)
            for j, i in enumerate((0, 1, 3, 4, 6, 7)):
                W.cb[j] = tl(pk[f"cb.{i}.w"], pk[f"cb.{i}.b"])
            W.ct0 = tl(pk["ct0.w"], pk["ct0.b"])
            for i in range(1, 6):
                W.convs_trans[i - 1] = tl(pk[f"convs_trans.{i}.w"], pk[f"convs_trans.{i}.b"])
                W.convs_rots[i - 1] = tl(pk[f"convs_rots.{i}.w"], pk[f"convs_rots.{i}.b"])
            W.fc_trans_w, W.fc_trans_b = pk["fc_trans.w_nhwc"].data_ptr(), self.fc_trans.bias.data_ptr()
            W.fc_rots_w, W.fc_rots_b = pk["fc_rots.w_nhwc"].data_ptr(), self.fc_rots.bias.data_ptr()
            r0, t0 = self.rot_emb_proj.layers[0], self.trans_emb_proj.layers[0]
            W.rot_emb0_w, W.rot_emb0_b, W.trans_emb0_w, W.trans_emb0_b = (x.data_ptr() for x in (r0.weight, r0.bias, t0.weight, t0.bias))
            for i in range(1, 6):
                W.rot_emb[i - 1] = tl(pk["rot_emb_proj.split"][i], self.rot_emb_proj.layers[i].bias)
                W.trans_emb[i - 1] = tl(pk["trans_emb_proj.split"][i], self.trans_emb_proj.layers[i].bias)
            W.rots_w, W.rots_b = self.rots.weight.data_ptr(), self.rots.bias.data_ptr()
            W.trans_w, W.trans_b = self.trans.weight.data_ptr(), self.trans.bias.data_ptr()
            W.fmt, W.passes = ops.SPLIT_F16, self.tc_passes
            pk["pixel_struct"] = W
        return pk["pixel_struct"]

    def _feature_planes(self, features1, features2):
        """res3 / res4 / res5 of both views as stacked NHWC planes: a backbone.PlaneFeatures as is, NCHW fp32 dicts converted."""
        from .backbone import PlaneFeatures
        if isinstance(features1, PlaneFeatures):
            B = features1.num_images // 2
            return B, {k: features1[k] for k in ("res3", "res4", "res5")}
        B = features1["res5"].shape[0]
        out = {}
        for name in ("res3", "res4", "res5"):
            f1, f2 = features1[name], features2[name]
            _, C, H, W = f1.shape
            sp = ops.Split(torch.empty(2 * B * H * W, C, device=f1.device, dtype=torch.float16),
                           torch.empty(2 * B * H * W, C, device=f1.device, dtype=torch.float16), C)
            ops.nchw_to_planes(f1.float(), out=sp, row_offset=0)
            ops.nchw_to_planes(f2.float(), out=sp, row_offset=B * H * W)
            out[name] = (sp, H, W)
        return B, out

    def refine_weights(self):
        """`nsac_refine_weights` for `ops.refine_forward` (borrowed pointers into the packed weights of this weight version; the
        struct and everything it points to live in the prepare() cache)."""
        pk = self.prepare_tc()
        if "refine_struct" not in pk:
            from . import _lib
            W = _lib.RefineWeights()
            keep = []

            def fill(arr, name, first=None):
                m = getattr(self, name)
                layers = list(m.layers)
                splits = list(pk[name + ".split"])
                if splits[0] is None:                       # geo_encoder: K = 8 layer runs on the CUDA cores
                    layers, splits = layers[1:], splits[1:]
                for i, (l, sp) in enumerate(zip(layers, splits)):
                    if i == 0 and first is not None:
                        arr[i] = ops.tc_layer(first, None)
                    else:
                        arr[i] = ops.tc_layer(sp, l.bias)
            g0 = self.geo_encoder.layers[0]
            W.geo0_w, W.geo0_b = g0.weight.data_ptr(), g0.bias.data_ptr()
            fill(W.geo_encoder, "geo_encoder")
            fill(W.geo_proj_s1, "geo_proj_s1")
            fill(W.decoder_rot, "decoder_rot")
            fill(W.geo_proj_s2, "geo_proj_s2")
            fill(W.decoder_tran, "decoder_tran")
            fill(W.decoder_rot2, "decoder_rot2", pk["decoder_rot2.w_geo_split"])
            fill(W.decoder_tran2, "decoder_tran2", pk["decoder_tran2.w_geo_split"])
            W.rot2_w_init, W.rot2_b0 = pk["decoder_rot2.w_init"].data_ptr(), self.decoder_rot2.layers[0].bias.data_ptr()
            W.tran2_w_init, W.tran2_b0 = pk["decoder_tran2.w_init"].data_ptr(), self.decoder_tran2.layers[0].bias.data_ptr()
            W.rots_w, W.rots_b = self.rots.weight.data_ptr(), self.rots.bias.data_ptr()
            W.trans_w, W.trans_b = self.trans.weight.data_ptr(), self.trans.bias.data_ptr()
            W.score_pack = pk["score_pack"].data_ptr()
            vh = pk["score_vecs_host"]
            W.score_vecs_host = None if vh is None else vh.data_ptr()
            rs, ts = ops._score_mlp_struct(pk["normal_score_proj"]), ops._score_mlp_struct(pk["param_score_proj"])
            keep += [rs, ts]
            import ctypes
            W.rot_mlp, W.tran_mlp = ctypes.pointer(rs), ctypes.pointer(ts)
            W.fmt, W.passes = ops.SPLIT_F16, self.tc_passes
            pk["refine_struct"], pk["refine_struct.keep"] = W, keep
        return pk["refine_struct"]

    @staticmethod
    def check_finite(outputs) -> None:
        """Debug guard (one host synchronisation; never called on the hot path): raises if an fp16 plane overflowed since the last
        check (`nsac_plane_overflow`, a sticky device flag set by the plane writers) or if the result rows are not finite.  The
        tensor-core layers carry activations as fp16 hi/lo planes (|x| <= 65504)."""
        pose = outputs[5]["pose"] if outputs[5] is not None else outputs[0]["camera"]["rot"]
        bad = ~torch.isfinite(pose).all(dim=-1)
        if ops.plane_overflow(clear=True):
            raise RuntimeError("fp16 plane overflow: a finite activation or input with |x| > 65504 was written into an fp16 hi/lo plane "
                               "since the last check (it became inf; normalisation layers can turn that back into finite but WRONG "
                               "poses) - feature maps / weights are far outside a trained network's scale; use bf16 planes "
                               "(nopesac_b200.ops.SPLIT_BF16) or rescale the inputs")
        if bool(bad.any()):
            raise RuntimeError(f"non-finite camera poses for pairs {bad.nonzero().flatten().tolist()}: an activation left the fp16 plane "
                               "range (|x| > 65504) - feature maps / weights are far outside a trained network's scale; run these "
                               "pairs with bf16 planes (nopesac_b200.ops.SPLIT_BF16) or rescale the inputs")

    # ------------------------------------------------------------------ forward
    def forward(self, features1, features2, planeParam1, planeParam2, planeApp1=None, planeApp2=None,
                gt_pose=None, gt_corr_matrix=None, batched_inputs=None, ite=0, matching_net=None,
                hyp_pairs=None, initial_pose=None, want_diag=False, assignment_override=None, result_exchange=None,
                plane_count1=None, plane_count2=None):
        if self.training:
            raise NotImplementedError("nopesac_b200.PlaneCameraHead is inference-only")
        return self.inference_Joint(features1, features2, planeParam1, planeParam2, planeApp1, planeApp2,
                                    matching_net=matching_net, hyp_pairs=hyp_pairs, initial_pose=initial_pose,
                                    want_diag=want_diag, assignment_override=assignment_override, result_exchange=result_exchange,
                                    plane_count1=plane_count1, plane_count2=plane_count2)

    @torch.no_grad()
    def inference_Joint(self, cam_feats1, cam_feats2, planeParam1, planeParam2, planeApp1, planeApp2,
                        gt_corr_matrix=None, batched_inputs=None, gt_pose=None, matching_net=None,
                        hyp_pairs=None, initial_pose=None, want_diag=False, assignment_override=None, result_exchange=None,
                        plane_count1=None, plane_count2=None):
        """`plane_count1` / `plane_count2` (int32 [B], device): ragged batch — pair b has that many planes in the first rows
        of the padded planeParam / planeApp tensors (what `nopesac_b200.plane_postprocess.PlaneLists` hands over: `.planes`,
        `.feats`, `.count`); every per-pair output equals the un-padded single-pair call, assignment matrices are zero and
        log-scores -inf outside the pair's block."""
        device = planeParam1.device
        B = planeParam1.shape[0]
        NQ = self.num_queries
        trans_list, rot_list = [], []
        # (no torch.tensor([...], device=...) here: that is a pageable H2D copy per call and not CUDA-graph capturable)
        zero_rot = torch.zeros(1, 4, device=device)
        zero_rot[:, 0] = 1.0
        output_cameras = {"camera_zero": {"tran": torch.zeros(1, 3, device=device), "rot": zero_rot}}
        out_cam_type = self.inference_out_cam_type if self.cam_ref_on else "initial"

        single_call = (self.use_stage_entry and self.cam_rec_on and self.plane_matcher_on and out_cam_type != "initial" and not want_diag
                       and assignment_override is None and getattr(matching_net, "use_stage_entry", False)
                       and not (out_cam_type == "max-score" and result_exchange is not None))
        from .backbone import RawImages
        raw = cam_feats1 if isinstance(cam_feats1, RawImages) else None
        if raw is not None and not (single_call and initial_pose is None and raw.single_call_ok):
            cam_feats1, raw = raw.features(), None           # backbone first, then the head on its feature planes
        if single_call and raw is not None:
            # ------------------------------------------------ stage set S5 behind ONE C call (nsac_model_forward): ResNet-50 on both
            # views' uint8 images -> the whole head + matcher
            PW, MW, RW = self.pixel_weights(), matching_net.match_weights(), self.refine_weights()
            BW = raw.backbone.backbone_weights(raw.images.shape[2], raw.images.shape[3])
            PW.passes = RW.passes = self.tc_passes
            BW.passes = raw.backbone.tc_passes
            MW.passes, MW.sinkhorn_iterations = matching_net.tc_passes, int(matching_net.sinkhorn_iterations)
            if (plane_count1 is None) != (plane_count2 is None):
                raise ValueError("plane_count1 and plane_count2 go together")
            if raw.num_images != 2 * B:
                raise ValueError(f"{raw.num_images} images for {B} pairs (expected both views stacked: 2B)")
            r = ops.model_forward(BW, PW, MW, RW, raw.images, planeParam1, planeParam2, planeApp1.float(), planeApp2.float(), NQ,
                                  self.matching_score_threshold, out_cam_type, plane_count1, plane_count2, hyp_pairs,
                                  exchange=result_exchange)
            return self._pack_single_call(r, output_cameras, trans_list, rot_list)
        if single_call:
            # ------------------------------------------------ the WHOLE head + matcher behind ONE C call (nsac_head_forward,
            # csrc/forward.cu: nsac_pixel_forward -> nsac_match_forward -> nsac_refine_forward on the current stream)
            PW, MW, RW = self.pixel_weights(), matching_net.match_weights(), self.refine_weights()
            PW.passes = RW.passes = self.tc_passes
            MW.passes, MW.sinkhorn_iterations = matching_net.tc_passes, int(matching_net.sinkhorn_iterations)
            if (plane_count1 is None) != (plane_count2 is None):
                raise ValueError("plane_count1 and plane_count2 go together")
            if initial_pose is None:
                _, lv = self._feature_planes(cam_feats1, cam_feats2)
                (p3, H3, W3), (p4, H4, W4), (p5, H5, W5) = lv["res3"], lv["res4"], lv["res5"]
                if (H4, W4, H5, W5) != (H3 // 2, W3 // 2, H3 // 4, W3 // 4) or H3 % 4 or W3 % 4:
                    raise ValueError(f"feature map sizes must halve level to level: res3 {H3}x{W3}, res4 {H4}x{W4}, res5 {H5}x{W5}")
            else:
                p3 = p4 = p5 = None
                H3 = W3 = 0
            r = ops.head_forward(PW, MW, RW, p3, p4, p5, B, H3, W3, planeParam1, planeParam2, planeApp1.float(), planeApp2.float(), NQ,
                                 self.matching_score_threshold, out_cam_type, plane_count1, plane_count2, hyp_pairs, initial_pose,
                                 exchange=result_exchange)
            return self._pack_single_call(r, output_cameras, trans_list, rot_list)

        if self.use_stage_entry and self.cam_rec_on and self.plane_matcher_on:
            # K1 + w >= 0 + K2 behind ONE C call (nsac_pixel_forward, csrc/forward.cu); the Python branch below issues the same launches
            W = self.pixel_weights()
            W.passes = self.tc_passes
            if initial_pose is None:
                _, lv = self._feature_planes(cam_feats1, cam_feats2)
                (p3, H3, W3), (p4, H4, W4), (p5, H5, W5) = lv["res3"], lv["res4"], lv["res5"]
                if (H4, W4, H5, W5) != (H3 // 2, W3 // 2, H3 // 4, W3 // 4) or H3 % 4 or W3 % 4:
                    raise ValueError(f"feature map sizes must halve level to level: res3 {H3}x{W3}, res4 {H4}x{W4}, res5 {H5}x{W5}")
                r = ops.pixel_forward(W, p3, p4, p5, B, H3, W3)
            else:
                r = ops.pixel_forward(W, None, None, None, B, 0, 0, initial_pose=initial_pose)
            initial_trans, initial_rot = r["init_tran"], r["init_rot"]
            q0, rot_feat0, t0, trans_feat0 = r["q0"], r["rot_feat0"], r["t0"], r["trans_feat0"]
            trans_list += [initial_trans, t0]
            rot_list += [initial_rot, q0]
            output_cameras["camera_init"] = {"tran": initial_trans, "rot": initial_rot}
            output_cameras["camera_initRec"] = {"tran": t0, "rot": q0}
        else:
            if initial_pose is None:
                initial_trans, initial_rot, pix_tfeat, pix_rfeat = self._forward_pixel_camera_head(cam_feats1, cam_feats2)
            else:
                initial_trans, initial_rot = initial_pose
                pix_tfeat = pix_rfeat = None
            # w >= 0, per pair (the reference flips the whole batch by sample 0, :436-437)
            initial_rot = torch.where(initial_rot[:, 0:1] < 0, -initial_rot, initial_rot)
            trans_list.append(initial_trans)
            rot_list.append(initial_rot)
            output_cameras["camera_init"] = {"tran": initial_trans, "rot": initial_rot}
            if not self.plane_matcher_on:
                output_cameras["camera"] = {"tran": trans_list[-1], "rot": rot_list[-1]}
                return output_cameras, trans_list, rot_list, [], {}, None

            if self.cam_rec_on:
                q0, rot_feat0, t0, trans_feat0 = self._forward_rec_heads(initial_rot, initial_trans)
                trans_list.append(t0)
                rot_list.append(q0)
                output_cameras["camera_initRec"] = {"tran": t0, "rot": q0}
            else:
                if pix_rfeat is None:
                    raise ValueError("initial_pose override needs CAM_REC_ON (the pixel features are skipped)")
                q0, rot_feat0, t0, trans_feat0 = initial_rot, pix_rfeat, initial_trans, pix_tfeat

        # ------------------------------------------------------------ matching (:493-503)
        if matching_net is None or not hasattr(matching_net, "match"):
            raise RuntimeError("matching_net must be a nopesac_b200.MatchingHead")
        cam = torch.cat([t0, q0], dim=-1)
        log_scores_padded, assignment = matching_net.match(planeApp1, planeApp2, cam, planeParam1, planeParam2,
                                                           match_threshold=self.matching_score_threshold,
                                                           plane_count1=plane_count1, plane_count2=plane_count2)
        output_planeAss = {"pred_assignment_beforeRef0": assignment.clone()}
        if out_cam_type == "initial":
            output_planeAss["pred_assignment"] = assignment.clone()
            output_cameras["camera"] = {"tran": trans_list[0], "rot": rot_list[0]}
            return output_cameras, trans_list, rot_list, [log_scores_padded], output_planeAss, None

        # ------------------------------------------------------------ geo sequences (:513-569)
        if assignment_override is not None and tuple(assignment_override.shape) != tuple(assignment.shape):
            raise ValueError(f"assignment_override must be [B,n1,n2] = {tuple(assignment.shape)}, got {tuple(assignment_override.shape)}")
        # 'max-score' picks argmax of the scores themselves: a discrete decision, so it takes the exact-fp32 scoring path
        # (the single-pass fp16 score MLPs of the tensor-core path move scores by up to ~4e-5 and could flip a near-tie)
        if out_cam_type == "max-score" and result_exchange is not None:
            raise NotImplementedError("INFERENCE_OUT_CAM_TYPE='max-score' runs the exact-fp32 scoring kernels, which have no fused "
                                      "result exchange: gather the rows with nopesac_b200.dist.gather_results instead")
        if not want_diag and self.use_stage_entry:
            # ---------------------------------------------------- K6 .. K10 behind ONE C call (nsac_refine_forward, csrc/forward.cu):
            # geo sequences (:513-569) -> hypothesis MLP chain (:957-986) -> per-hypothesis poses -> scoring + selection
            # (:964-1115) -> pruning (:605-629); launch for launch the sequence of the Python branch below
            W = self.refine_weights()
            W.passes = self.tc_passes
            r = ops.refine_forward(W, planeParam1, planeParam2, assignment if assignment_override is None else assignment_override,
                                   t0, q0, rot_feat0, trans_feat0, NQ, out_cam_type, hyp_pairs=hyp_pairs,
                                   prune=assignment_override is None, exchange=result_exchange)
            geo_local, geo_global, sig, matched_num, pair_idx = r["geo_local"], r["geo_global"], r["sig"], r["matched_num"], r["pair_idx"]
            q_h, t_h, pose = r["q_h"], r["t_h"], r["pose"]
            res = {"score_rot": r["score_rot"], "score_tran": r["score_tran"], "sel_idx": r["sel_idx"], "diag": None}
            pruned = r["assign_pruned"] if assignment_override is None else ops.prune_assignment(assignment, planeParam1, planeParam2, pose)
        else:
            geo_local, geo_global, sig, geo8, matched_num, pair_idx = ops.geo_sequence(
                planeParam1, planeParam2, assignment if assignment_override is None else assignment_override,
                t0, q0, NQ, hyp_pairs=hyp_pairs)

            # ------------------------------------------------------------ refinement head (:925-1115)
            fused_rot, fused_tran = self._hypothesis_features(geo8, rot_feat0, trans_feat0, B, NQ)
            q_h, t_h = ops.pose_heads(fused_rot, fused_tran, self.rots.weight, self.rots.bias,
                                      self.trans.weight, self.trans.bias)
            pk = self.prepare_tc()
            precision = "fp32" if out_cam_type == "max-score" else "fp16"
            res = ops.score_aggregate(geo_local, q_h.view(B, NQ, 4), t_h.view(B, NQ, 3), q0, t0,
                                      fused_rot.view(B, NQ, 256), fused_tran.view(B, NQ, 256), rot_feat0, trans_feat0,
                                      matched_num, pk["normal_score_proj"], pk["param_score_proj"],
                                      self.rots.weight, self.rots.bias, self.trans.weight, self.trans.bias,
                                      out_cam_type=out_cam_type, want_scores=True, want_diag=want_diag,
                                      precision=precision, pack=pk["score_pack"], exchange=result_exchange,
                                      vecs_host=pk["score_vecs_host"])
            pose = res["pose"]
            # ------------------------------------------------------------ assignment pruning (:605-629)
            # (the sign flip of :600-601 does not change R, which is quadratic in q)
            pruned = ops.prune_assignment(assignment, planeParam1, planeParam2, pose)
        return self._pack_refined(output_cameras, trans_list, rot_list, log_scores_padded, output_planeAss, pose, pruned, q0, t0, q_h, t_h,
                                  res, sig, matched_num, pair_idx, geo_local, geo_global, want_diag)

    def _pack_single_call(self, r, output_cameras, trans_list, rot_list):
        """The reference's return value from the output dict of ops.head_forward / ops.model_forward."""
        trans_list += [r["init_tran"], r["t0"]]
        rot_list += [r["init_rot"], r["q0"]]
        output_cameras["camera_init"] = {"tran": r["init_tran"], "rot": r["init_rot"]}
        output_cameras["camera_initRec"] = {"tran": r["t0"], "rot": r["q0"]}
        output_planeAss = {"pred_assignment_beforeRef0": r["assign"]}
        res = {"score_rot": r["score_rot"], "score_tran": r["score_tran"], "sel_idx": r["sel_idx"], "diag": None}
        return self._pack_refined(output_cameras, trans_list, rot_list, r["log_scores_padded"], output_planeAss, r["pose"],
                                  r["assign_pruned"], r["q0"], r["t0"], r["q_h"], r["t_h"], res, r["sig"], r["matched_num"],
                                  r["pair_idx"], r["geo_local"], r["geo_global"], False)

    def _pack_refined(self, output_cameras, trans_list, rot_list, log_scores_padded, output_planeAss, pose, pruned, q0, t0, q_h, t_h,
                      res, sig, matched_num, pair_idx, geo_local, geo_global, want_diag):
        """The reference's return value of inference_Joint (:589-640) from the tensors of the refinement stage."""
        B, NQ = pose.shape[0], self.num_queries
        ref_trans, ref_rot = pose[:, 0:3], pose[:, 3:7]
        avg_trans, avg_rot = pose[:, 7:10], pose[:, 10:14]
        trans_list += [avg_trans, ref_trans]
        rot_list += [avg_rot, ref_rot]
        output_cameras["camera_avgRef0"] = {"tran": avg_trans, "rot": avg_rot}
        output_cameras["camera_softRef0"] = {"tran": ref_trans, "rot": ref_rot}

        output_planeAss["pred_assignment_afterRef0"] = pruned.clone()
        output_planeAss["pred_assignment"] = pruned.clone()
        output_cameras["camera"] = {"tran": ref_trans, "rot": ref_rot}

        all_rots = torch.cat([q0.unsqueeze(1), q_h.view(B, NQ, 4)], 1)
        all_trans = torch.cat([t0.unsqueeze(1), t_h.view(B, NQ, 3)], 1)
        output_cameras["camera_onePP"] = {"tran": all_trans, "rot": all_rots}   # padded to NQ+1; see matched_num
        pose_ref_outputs = {
            "pred_trans": ref_trans, "pred_rot": ref_rot, "pred_trans_avg": avg_trans, "pred_rot_avg": avg_rot,
            "all_pred_trans": all_trans, "all_pred_rots": all_rots,
            "score_soft_rot": res["score_rot"], "score_soft_offset": res["score_tran"],
            "sig_seq": sig, "matched_num": matched_num, "sel_idx": res["sel_idx"], "pair_idx": pair_idx,
            "geo_local": geo_local, "geo_global": geo_global, "pose": pose,
        }
        if want_diag:
            pose_ref_outputs.update(l2_dist=res["diag"][0], normal_dist=res["diag"][1], offset_dist=res["diag"][2])
        return output_cameras, trans_list, rot_list, [log_scores_padded], output_planeAss, pose_ref_outputs
