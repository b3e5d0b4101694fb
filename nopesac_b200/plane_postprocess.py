"""Plane-list extraction from the PlaneTRHead outputs — the step before the hot path.

Mirrors `PlaneTR_NopeSAC._postprocess_planeHeadMask` (meta_arch/siamese_planeTR.py:625-803): same inputs (`pred_logits`,
`pred_params`, `pred_mask_logits`, `query_feat`), same thresholds (`cfg.TEST.PLANE_SCORE_THRESHOLD`, `MASK_PROB_THRESHOLD`,
`OVERLAP_THRESHOLD`), same per-image result keys.  The reference handles one image with a Python loop over the queries and a
device->host copy per plane; `postprocess_plane_head_mask` runs a whole batch through `nsac_plane_postprocess`
(csrc/planes.cu) and returns a `PlaneLists` of device tensors — no host sync, so the plane parameters and query features can be
handed to the camera head directly.  `PlaneLists.to_reference_results()` is the (synchronising) conversion into the
reference's list of dicts for callers that want exactly that format.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib, ops

FLAG_ZERO, FLAG_FALLBACK, FLAG_PATCH00 = 1, 2, 4
NO_PLANE = 255


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr())


@dataclass
class PlaneLists:
    """Device-resident result for B images, every per-plane tensor padded to NQ rows (kept planes first, in query order)."""
    count: torch.Tensor      # int32 [B]        number of kept planes n_b
    flags: torch.Tensor      # int32 [B]        FLAG_* bits
    ori_idx: torch.Tensor    # int32 [B,NQ]     query index of every kept plane (-1 = padding)  -> pred_plane_oriIdxs
    planes: torch.Tensor     # fp32  [B,NQ,3]   -> pred_plane
    feats: torch.Tensor      # fp32  [B,NQ,C]   -> pred_plane_feats
    scores: torch.Tensor     # fp32  [B,NQ]     -> instances[j]["score"]
    centers: torch.Tensor    # fp32  [B,NQ,2]   -> pred_plane_ins_center
    bboxes: torch.Tensor     # fp32  [B,NQ,4]   -> instances[j]["bbox"] (x, y, w, h), bbox_mode 1
    areas: torch.Tensor      # int32 [B,NQ]
    seg: torch.Tensor        # uint8 [B,H,W]    kept-list index per pixel, 255 = none; pred_plane_masks[j] = (seg == j)

    def masks(self, b: int, n: Optional[int] = None) -> torch.Tensor:
        """bool [n,H,W] = the reference's `pred_plane_masks` of image b (n read from the device if not given)."""
        n = int(self.count[b]) if n is None else n
        return self.seg[b].unsqueeze(0) == torch.arange(n, device=self.seg.device, dtype=torch.uint8).view(n, 1, 1)

    def to_reference_results(self, batched_inputs: Optional[List[dict]] = None, with_rle: bool = False) -> List[Dict]:
        """The reference's per-image dicts (:792-801).  Synchronises once (reads `count`).  `instances[j]["segmentation"]` holds
        COCO RLE `counts` only if `with_rle` (encoded on the host from the label map, `rle_counts`)."""
        counts = self.count.cpu().tolist()
        H, W = self.seg.shape[1:]
        out = []
        for b, n in enumerate(counts):
            meta = batched_inputs[b] if batched_inputs is not None else {}
            res = {"image_id": meta.get("image_id", b), "file_name": meta.get("file_name", "")}
            masks = self.masks(b, n)
            scores = self.scores[b, :n].cpu().tolist()
            bboxes = self.bboxes[b, :n].cpu().tolist()
            masks_host = masks.cpu().numpy() if with_rle else None
            res["instances"] = []
            for j in range(n):
                seg = {"size": [H, W]}
                if with_rle:
                    seg["counts"] = rle_counts(masks_host[j])
                res["instances"].append({"image_id": res["image_id"], "file_name": res["file_name"], "category_id": 0,
                                         "score": scores[j], "segmentation": seg, "bbox": bboxes[j], "bbox_mode": 1})
            res["pred_plane"] = self.planes[b, :n]
            res["pred_plane_feats"] = self.feats[b, :n].unsqueeze(0).contiguous()
            res["pred_plane_oriIdxs"] = list(self.ori_idx[b, :n].unbind(0))
            res["pred_plane_masks"] = masks
            res["pred_plane_ins_center"] = self.centers[b, :n].reshape(-1, 2)
            out.append(res)
        return out


def postprocess_plane_head_mask(planeTR_outputs: Dict[str, torch.Tensor], query_feat: torch.Tensor, height: int = 480,
                                width: int = 640, plane_score_threshold: float = 0.6, mask_prob_threshold: float = 0.5,
                                overlap_threshold: float = 0.6) -> PlaneLists:
    """Batched `_postprocess_planeHeadMask`: `planeTR_outputs` = {'pred_logits' [B,NQ,2], 'pred_params' [B,NQ,3],
    'pred_mask_logits' [B,NQ,h,w]}, `query_feat` [B,NQ,C]; CUDA fp32."""
    logits, params, masks = planeTR_outputs["pred_logits"], planeTR_outputs["pred_params"], planeTR_outputs["pred_mask_logits"]
    logits, params, masks, query_feat = (ops._chk(t.detach(), n).contiguous() for t, n in (
        (logits, "pred_logits"), (params, "pred_params"), (masks, "pred_mask_logits"), (query_feat, "query_feat")))
    B, NQ = logits.shape[:2]
    h, w = masks.shape[-2:]
    Cf = query_feat.shape[-1]
    assert logits.shape == (B, NQ, 2) and params.shape == (B, NQ, 3) and masks.shape == (B, NQ, h, w) and query_feat.shape == (B, NQ, Cf)
    dev = logits.device
    L = _lib.lib()
    ws_bytes = L.nsac_plane_post_workspace_bytes(B, NQ, height, width)
    if ws_bytes == 0:
        raise RuntimeError(f"nsac_plane_post_workspace_bytes: bad shape B={B} NQ={NQ} H={height} W={width}")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    out = PlaneLists(count=torch.empty(B, **i32), flags=torch.empty(B, **i32), ori_idx=torch.empty(B, NQ, **i32),
                     planes=torch.empty(B, NQ, 3, **f32), feats=torch.empty(B, NQ, Cf, **f32), scores=torch.empty(B, NQ, **f32),
                     centers=torch.empty(B, NQ, 2, **f32), bboxes=torch.empty(B, NQ, 4, **f32), areas=torch.empty(B, NQ, **i32),
                     seg=torch.empty(B, height, width, dtype=torch.uint8, device=dev))
    st = L.nsac_plane_postprocess(_ptr(logits), _ptr(params), _ptr(masks), _ptr(query_feat), B, NQ, Cf, h, w, height, width,
                                  float(plane_score_threshold), float(mask_prob_threshold), float(overlap_threshold),
                                  _ptr(out.count), _ptr(out.flags), _ptr(out.ori_idx), _ptr(out.planes), _ptr(out.feats),
                                  _ptr(out.scores), _ptr(out.centers), _ptr(out.bboxes), _ptr(out.areas), _ptr(out.seg), _ptr(ws),
                                  ops._stream())
    _lib.check(st, "nsac_plane_postprocess")
    ops._count(4)
    return out


def rle_counts(mask: np.ndarray) -> bytes:
    """COCO compressed RLE `counts` of a [H,W] bool mask on the host (what `pycocotools.mask.encode(np.asfortranarray(m))["counts"]`
    returns, siamese_planeTR.py:703): column-major run lengths starting with a zero run, delta-coded against counts[i-2] for
    i > 2, 5 payload bits + continuation bit per character, offset 48."""
    flat = np.asarray(mask, dtype=np.uint8).reshape(-1, order="F")
    if flat.size == 0:
        return b""
    edges = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    runs = np.diff(np.concatenate(([0], edges, [flat.size]))).astype(np.int64)
    if flat[0]:
        runs = np.concatenate(([0], runs))
    delta = runs.copy()
    delta[3:] -= runs[1:-2]
    out = bytearray()
    for x in delta.tolist():
        while True:
            c = x & 0x1F
            x >>= 5
            more = (x != -1) if (c & 0x10) else (x != 0)
            out.append((c | 0x20 if more else c) + 48)
            if not more:
                break
    return bytes(out)
