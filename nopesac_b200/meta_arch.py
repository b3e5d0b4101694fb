"""META_ARCH boundary of the path: `PlaneTR_NopeSAC` under `META_ARCH_REGISTRY`, the thing that CALLS the
hot path in the reference (meta_arch/siamese_planeTR.py:33-34, 338-450).

Only the hot-path children are built natively — `matching_head` and `camera_head_list.0`, under the
reference's attribute names so the `matching_head.*` / `camera_head_list.0.*` keys of a reference
checkpoint load unchanged (`load_reference_state_dict`).  The single-view plane detector (backbone +
PlaneTRHead + mask post-processing, siamese_planeTR.py:452-473, 625-803) is upstream of the path and out
of scope this round: its per-view outputs enter through `batched_inputs`:

    batched_inputs = [{"0": view, "1": view}, ...]      # any batch size (the reference asserts 1, :340)
    view = {"pred_plane": [n,3], "pred_plane_feats": [n,256] or [1,n,256],
            "cam_feats": {"res2".."res5": [C,h,w] or [1,C,h,w]}}

and `forward` returns the reference's per-pair result dicts (:411-431): every `camera*` key with
{"tran","rot"} and the `pred_assignment*` matrices (as device tensors; `.cpu().numpy()` packing of
:384-399 is left to the caller so that no host sync happens inside).  Pairs of one call may have different
numbers of planes: they are zero-padded to the largest count and the kernels get the per-pair counts (ragged batch).
"""
from __future__ import annotations

from typing import Dict, List

import torch
from torch import nn

from .backbone import build_backbone
from .camera_head import build_camera_head
from .compat import Registry, ShapeSpec
from .matching_head import build_matching_head
from .planeTR_head import build_planeTR_head
from .plane_postprocess import PlaneLists, postprocess_plane_head_mask

__all__ = ["META_ARCH_REGISTRY", "PlaneTR_NopeSAC", "build_model"]

META_ARCH_REGISTRY = Registry("META_ARCH")

# detectron2 build_resnet_backbone(R50).output_shape() for OUT_FEATURES res2..res5
RESNET50_OUTPUT_SHAPE = {
    "res2": ShapeSpec(channels=256, stride=4), "res3": ShapeSpec(channels=512, stride=8),
    "res4": ShapeSpec(channels=1024, stride=16), "res5": ShapeSpec(channels=2048, stride=32),
}


def build_model(cfg):
    """detectron2.modeling.build_model: META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg)."""
    model = META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg)
    return model.to(torch.device(cfg.MODEL.DEVICE)) if torch.cuda.is_available() or cfg.MODEL.DEVICE == "cpu" else model


@META_ARCH_REGISTRY.register()
class PlaneTR_NopeSAC(nn.Module):
    def __init__(self, cfg, with_backbone: bool = False, with_plane_head: bool = False):
        """`with_backbone=True` also builds `self.backbone` (row f2: the R-50 of `build_backbone(cfg)`, siamese_planeTR.py
        `__init__`; `backbone.*` checkpoint keys then load too) so that `inference_from_images` starts at RGB.
        `with_plane_head=True` (needs the backbone) also builds `self.sem_seg_head` (row f1: `build_planeTR_head`,
        siamese_planeTR.py:64): `forward` then accepts the reference's own batched_inputs with `image` entries and runs the
        whole model, `inference_from_rgb`."""
        super().__init__()
        self.cfg = cfg
        self.backbone = build_backbone(cfg) if with_backbone else None
        if with_plane_head and not with_backbone:
            raise ValueError("with_plane_head=True needs with_backbone=True")
        self.sem_seg_head = build_planeTR_head(cfg, self.backbone.output_shape()) if with_plane_head else None
        self.mask_on = cfg.MODEL.MASK_ON
        self.embedding_on = cfg.MODEL.EMBEDDING_ON
        self.camera_on = cfg.MODEL.CAMERA_ON
        self.camera_refine_on = cfg.MODEL.CAMERA_HEAD.REFINE_ON
        self.num_queries = cfg.MODEL.SEM_SEG_HEAD.NUM_OBJECT_QUERIES
        self.overlap_threshold = cfg.TEST.OVERLAP_THRESHOLD                  # siamese_planeTR.py:92-94
        self.plane_score_threshold = cfg.TEST.PLANE_SCORE_THRESHOLD
        self.mask_prob_threshold = cfg.TEST.MASK_PROB_THRESHOLD
        # camCls kmeans pickles (siamese_planeTR.py:119-128) are loaded by the reference but never used in
        # forward, and cannot be unpickled without sklearn 0.21 / spherecluster: deliberately skipped.
        self.matching_head = build_matching_head(cfg) if self.embedding_on else None
        self.camera_head_list = nn.ModuleList()
        if self.camera_on:
            self.camera_head_list.append(build_camera_head(cfg, RESNET50_OUTPUT_SHAPE))
        self.eval()

    @property
    def device(self):
        return next(self.parameters()).device

    def load_reference_state_dict(self, state_dict: Dict[str, torch.Tensor]):
        """Loads the hot-path keys of a full reference checkpoint; returns the keys that were ignored
        (backbone / sem_seg_head / criterion)."""
        mine = {k: v for k, v in state_dict.items()
                if k.startswith("matching_head.") or k.startswith("camera_head_list.0.") or
                (self.backbone is not None and k.startswith("backbone.")) or
                (self.sem_seg_head is not None and k.startswith("sem_seg_head."))}
        missing, unexpected = self.load_state_dict(mine, strict=False)
        if missing:
            raise KeyError(f"reference checkpoint lacks hot-path keys: {missing[:5]} ...")
        return sorted(set(state_dict) - set(mine))

    def plane_lists(self, planeTR_outputs: Dict[str, torch.Tensor], query_feat: torch.Tensor, height: int = 480,
                    width: int = 640) -> PlaneLists:
        """Row f1, batched and sync-free: PlaneTRHead outputs -> device-resident plane lists (csrc/planes.cu)."""
        return postprocess_plane_head_mask(planeTR_outputs, query_feat, height, width, self.plane_score_threshold,
                                           self.mask_prob_threshold, self.overlap_threshold)

    def _postprocess_planeHeadMask(self, planeTR_outputs, pred_depth, batched_inputs, image_sizes, query_feat_in,
                                   mask_threshold=0.5, nms=False):
        """The reference's method (siamese_planeTR.py:625-803), same arguments and per-image result dicts; all images of the
        call must share one output size (the reference asserts batch size 1, :454).  One host sync (the plane counts)."""
        sizes = {(bi.get("height", sz[0]), bi.get("width", sz[1])) for bi, sz in zip(batched_inputs, image_sizes)}
        if len(sizes) != 1:
            raise ValueError(f"_postprocess_planeHeadMask: images of one call must share one output size, got {sorted(sizes)}")
        height, width = next(iter(sizes))
        lists = self.plane_lists(planeTR_outputs, query_feat_in, height, width)
        return lists.to_reference_results(batched_inputs, with_rle=True)

    @torch.no_grad()
    def inference_from_images(self, images1: torch.Tensor, images2: torch.Tensor, planeParam1, planeParam2, planeApp1, planeApp2,
                              **head_kwargs):
        """RGB -> backbone -> camera head (rows f2 + a2..a15): `images*` [B,3,H,W] uint8 (fast stem) or float in 0..255
        (normalisation is fused into the stem), both views go through the backbone as one batch of 2B images, `res3..res5` feed the pixel pose CNN, the
        plane lists come from the caller (PlaneTRHead is not built here).  Returns the camera head's 6-tuple."""
        if self.backbone is None:
            raise RuntimeError("PlaneTR_NopeSAC was built without a backbone (with_backbone=True)")
        # `images2 is None`: `images1` already holds both views stacked, [2B,3,H,W] = first views then second views
        images = images1 if images2 is None else torch.cat([images1, images2], 0)
        # both views stacked; the head runs backbone + head behind ONE C call (nsac_model_forward) when it can, otherwise the
        # backbone's NHWC hi/lo planes are handed over (no NCHW round trip either way)
        from .backbone import RawImages
        feats = RawImages(self.backbone, images)
        return self.camera_head_list[0](feats, None, planeParam1, planeParam2, planeApp1=planeApp1, planeApp2=planeApp2,
                                        matching_net=self.matching_head, **head_kwargs)

    @torch.no_grad()
    def inference_from_plane_heads(self, planeTR_outputs1, query_feat1, planeTR_outputs2, query_feat2, cam_feats1, cam_feats2,
                                   height: int = 480, width: int = 640, max_planes: int = None, **head_kwargs):
        """Rows f1 + a2..a15 without a host round trip: the PlaneTRHead outputs of both views of B pairs -> plane lists
        (`plane_lists`) -> camera head on the padded lists with per-pair plane counts (`plane_count1/2`).  This replaces
        `inference_single`'s post-processing + the `.unsqueeze(0).to(device)` hand-off of siamese_planeTR.py:364-383, which
        goes through the host for every plane.  `max_planes` (a caller-side bound on planes per view, e.g. 20) trims the padded
        lists from NUM_OBJECT_QUERIES rows to that many — counts are clamped to it — so the matcher does not work on padding.
        Returns (head outputs 6-tuple, PlaneLists view 1, PlaneLists view 2); everything stays on the device."""
        l1 = self.plane_lists(planeTR_outputs1, query_feat1, height, width)
        l2 = self.plane_lists(planeTR_outputs2, query_feat2, height, width)
        P = self.num_queries if max_planes is None else min(int(max_planes), self.num_queries)
        c1, c2 = l1.count.clamp(max=P), l2.count.clamp(max=P)
        out = self.camera_head_list[0](cam_feats1, cam_feats2, l1.planes[:, :P].contiguous(), l2.planes[:, :P].contiguous(),
                                       planeApp1=l1.feats[:, :P].contiguous(), planeApp2=l2.feats[:, :P].contiguous(),
                                       matching_net=self.matching_head, plane_count1=c1, plane_count2=c2, **head_kwargs)
        return out, l1, l2

    @staticmethod
    def _stack_views(batched_inputs, view: str, device):
        """Per-view inputs of B pairs -> padded [B,Pmax,3] / [B,Pmax,256] tensors + plane counts (None if all pairs have the
        same number of planes in this view) + stacked camera feature maps.  Plane counts come from tensor SHAPES (host-side
        metadata): no device synchronisation."""
        planes = [bi[view]["pred_plane"].reshape(-1, 3) for bi in batched_inputs]
        feats = [bi[view]["pred_plane_feats"].reshape(-1, 256) for bi in batched_inputs]
        ns = [p.shape[0] for p in planes]
        if any(f.shape[0] != n for f, n in zip(feats, ns)) or min(ns) < 1:
            raise ValueError(f"view {view}: pred_plane / pred_plane_feats disagree or are empty (planes per pair: {ns})")
        count = None
        if len(set(ns)) > 1:
            pmax = max(ns)
            pad = lambda t, n: torch.cat([t, t.new_zeros(pmax - n, t.shape[1])]) if n < pmax else t
            planes = [pad(p, n) for p, n in zip(planes, ns)]
            feats = [pad(f, n) for f, n in zip(feats, ns)]
            count = torch.tensor(ns, dtype=torch.int32).to(device, non_blocking=True)
        planes = torch.stack(planes).to(device).float()
        feats = torch.stack(feats).to(device).float()
        cam = {}
        for k in batched_inputs[0][view]["cam_feats"]:
            cam[k] = torch.stack([bi[view]["cam_feats"][k].reshape(bi[view]["cam_feats"][k].shape[-3:])
                                  for bi in batched_inputs]).to(device).float()
        return planes, feats, cam, count

    @torch.no_grad()
    def inference_from_rgb(self, batched_inputs: List[dict], max_planes: int = None, **head_kwargs):
        """The reference's whole inference (siamese_planeTR.py:338-450 with inference_single :452-473) for ANY number of pairs in
        one call: `batched_inputs[i]["0" / "1"]["image"]` = uint8 (or float 0..255) [3,H,W] of one size ->
        preprocess (normalisation fused into the stem) -> backbone on all 2B images -> PlaneTRHead -> plane lists on the device
        (`plane_lists`, no per-plane host round trip) -> camera head on the padded lists with per-pair plane counts.
        Returns (results, lists1, lists2): `results[i]` has the reference's camera / assignment keys (:411-431) as device tensors
        (assignment matrices padded to `max_planes` or NUM_OBJECT_QUERIES rows / columns: see lists*.count), `lists*` are the
        PlaneLists of the first / second views (`.to_reference_results()` gives the reference's per-view dicts)."""
        if self.backbone is None or self.sem_seg_head is None:
            raise RuntimeError("PlaneTR_NopeSAC was built without backbone / plane head (with_backbone=True, with_plane_head=True)")
        B, dev = len(batched_inputs), self.device
        imgs = [bi[v]["image"] for v in ("0", "1") for bi in batched_inputs]               # first views, then second views
        images = torch.stack([im.to(dev, non_blocking=True) for im in imgs])
        if images.dtype != torch.uint8:
            images = images.float()
        H, W = images.shape[2:]
        feats = self.backbone(images, planes=True)
        outputs, query_feat = self.sem_seg_head(feats)
        lists = self.plane_lists(outputs, query_feat, H, W)
        P = self.num_queries if max_planes is None else min(int(max_planes), self.num_queries)
        cnt = lists.count.clamp(max=P)
        pl, ft = lists.planes[:, :P], lists.feats[:, :P]
        out = self.camera_head_list[0](feats, None, pl[:B].contiguous(), pl[B:].contiguous(), planeApp1=ft[:B].contiguous(),
                                       planeApp2=ft[B:].contiguous(), matching_net=self.matching_head, plane_count1=cnt[:B].contiguous(),
                                       plane_count2=cnt[B:].contiguous(), **head_kwargs)
        cams, planeAss = out[0], out[4]
        results = []
        for i in range(B):
            r = {"pred_aff": None, "depth": {"0": None, "1": None}}
            for key, value in cams.items():
                j = i if value["tran"].shape[0] == B else 0      # camera_zero is [1,3] regardless of B
                r[key] = {"tran": value["tran"][j], "rot": value["rot"][j]}
            for key, value in planeAss.items():
                r[key] = value[i]
            results.append(r)
        sl = lambda a, b: PlaneLists(*[getattr(lists, f)[a:b] for f in ("count", "flags", "ori_idx", "planes", "feats", "scores", "centers",
                                                                         "bboxes", "areas", "seg")])
        return results, sl(0, B), sl(B, 2 * B), out

    @torch.no_grad()
    def forward(self, batched_inputs: List[dict]):
        if self.sem_seg_head is not None and "image" in batched_inputs[0]["0"]:
            return self.inference_from_rgb(batched_inputs)[0]
        return self.inference(batched_inputs)

    @torch.no_grad()
    def inference(self, batched_inputs: List[dict]):
        assert not self.training
        B = len(batched_inputs)
        if not self.camera_on:
            zero = {"camera": {"tran": torch.zeros(3), "rot": torch.tensor([1., 0., 0., 0.])}}
            return [dict(zero) for _ in range(B)]
        dev = self.device
        p1, a1, f1, c1 = self._stack_views(batched_inputs, "0", dev)
        p2, a2, f2, c2 = self._stack_views(batched_inputs, "1", dev)
        if (c1 is None) != (c2 is None):           # ragged in one view only: the other view's counts are all Pmax
            full = lambda p: torch.full((B,), p.shape[1], dtype=torch.int32, device=dev)
            c1, c2 = (full(p1) if c1 is None else c1), (full(p2) if c2 is None else c2)
        cams, _, _, _, planeAss, _ = self.camera_head_list[0](
            f1, f2, p1, p2, planeApp1=a1, planeApp2=a2, batched_inputs=batched_inputs,
            matching_net=self.matching_head, plane_count1=c1, plane_count2=c2)
        n1s = [bi["0"]["pred_plane"].reshape(-1, 3).shape[0] for bi in batched_inputs]
        n2s = [bi["1"]["pred_plane"].reshape(-1, 3).shape[0] for bi in batched_inputs]
        results = []
        for i in range(B):
            r = {"0": batched_inputs[i]["0"], "1": batched_inputs[i]["1"], "pred_aff": None,
                 "depth": {"0": None, "1": None}}
            for key, value in cams.items():
                j = i if value["tran"].shape[0] == B else 0      # camera_zero is [1,3] regardless of B
                r[key] = {"tran": value["tran"][j], "rot": value["rot"][j]}
            for key, value in planeAss.items():
                r[key] = value[i, :n1s[i], :n2s[i]]        # the pair's own [n1, n2] block of the padded batch
            results.append(r)
        return results
