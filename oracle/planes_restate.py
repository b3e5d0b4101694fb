"""TEST INFRASTRUCTURE — not product code.

CPU restatement (plain PyTorch / numpy, reference op order, one image at a time) of the step right BEFORE the hot path,
SURVEY.md §8 row f1: `PlaneTR_NopeSAC._postprocess_planeHeadMask`, meta_arch/siamese_planeTR.py:625-803 — from the PlaneTRHead
outputs (`pred_logits [B,NQ,2]`, `pred_params [B,NQ,3]`, `pred_mask_logits [B,NQ,h,w]`, `query_feat [B,NQ,C]`) to the
per-view plane list the camera head consumes (`pred_plane`, `pred_plane_feats`, masks, centres, instances).

Pinned (tests/test_oracle_planes.py): every tensor / number of the result is compared with the reference's own method, cut
out of the unmodified source file with `ast` and executed (oracle/ref_planes_loader.py), and with the golden fixture
tests/golden/planes_post.golden generated from it (tests/golden/make_planes_golden.py).

Third-party arithmetic that is ABSENT here: `pycocotools.mask.encode / toBbox` (pycocotools 2.0.x, `common/maskApi.c`
rleEncode / rleToBbox / rleToString; not installed, not vendored by the reference).  `rle_encode`, `rle_to_bbox` and
`rle_to_string` restate the published algorithm; the reference pin runs the reference method with these three as its
`mask_util`, so for `bbox` and `segmentation.counts` parity is UNPINNED (restated algorithm only); everything else is pinned.

Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------------------------
# pycocotools maskApi.c, restated
# ---------------------------------------------------------------------------------------------------------------------
def rle_encode(mask: np.ndarray) -> list:
    """rleEncode: run lengths of the column-major (Fortran order) mask, starting with a run of zeros."""
    flat = np.asarray(mask).astype(np.uint8).reshape(-1, order="F")
    if flat.size == 0:
        return []
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    bounds = np.concatenate([[0], change, [flat.size]])
    counts = np.diff(bounds).tolist()
    if flat[0] == 1:
        counts = [0] + counts
    return counts


def rle_to_bbox(counts: list, h: int, w: int) -> list:
    """rleToBbox: [x, y, w, h] of the ones (tight box; all zeros -> [0,0,0,0])."""
    m = (len(counts) // 2) * 2
    if m == 0:
        return [0.0, 0.0, 0.0, 0.0]
    xs, ys, xe, ye, cc, xp = w, h, 0, 0, 0, 0
    for j in range(m):
        cc += counts[j]
        t = cc - j % 2
        y = t % h
        x = (t - y) // h
        if j % 2 == 0:
            xp = x
        elif xp < x:
            ys, ye = 0, h - 1
        xs, xe, ys, ye = min(xs, x), max(xe, x), min(ys, y), max(ye, y)
    return [float(xs), float(ys), float(xe - xs + 1), float(ye - ys + 1)]


def rle_to_string(counts: list) -> bytes:
    """rleToString: LEB128-like, 6 bits per character (ascii 48..111), counts[i>2] delta-coded against counts[i-2]."""
    out = bytearray()
    for i, c in enumerate(counts):
        x = int(c)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(ch + 48)
    return bytes(out)


class mask_util:
    """The two pycocotools.mask calls of siamese_planeTR.py:703-704 on top of the restated algorithm."""

    @staticmethod
    def encode(mask_fortran: np.ndarray) -> dict:
        h, w = mask_fortran.shape
        return {"size": [h, w], "counts": rle_to_string(rle_encode(mask_fortran)), "_raw": rle_encode(mask_fortran)}

    @staticmethod
    def toBbox(rle: dict) -> np.ndarray:
        return np.asarray(rle_to_bbox(rle["_raw"], rle["size"][0], rle["size"][1]), dtype=np.float64)


# ---------------------------------------------------------------------------------------------------------------------
# siamese_planeTR.py:804-812
# ---------------------------------------------------------------------------------------------------------------------
def normalized_xy_map(h: int = 480, w: int = 640) -> np.ndarray:
    xy = np.zeros((2, h, w), dtype=np.float32)
    xy[0] = (np.arange(w, dtype=np.float64) / w)[None, :]
    xy[1] = (np.arange(h, dtype=np.float64) / h)[:, None]
    return xy


# ---------------------------------------------------------------------------------------------------------------------
# siamese_planeTR.py:625-803
# ---------------------------------------------------------------------------------------------------------------------
def _center(mask_np: np.ndarray, xy: np.ndarray, eps: float) -> np.ndarray:
    """:726-739 / :775-788 (float64 sums of the float32 maps; eps = 1e-10 on the regular branch, 0 on the fallback)."""
    plane_mask = mask_np.astype(np.float64)
    pixel_num = plane_mask.sum()
    with np.errstate(invalid="ignore", divide="ignore"):
        cx = (xy[0] * plane_mask).sum() / (pixel_num + eps)
        cy = (xy[1] * plane_mask).sum() / (pixel_num + eps)
    c = np.zeros([2], dtype=np.float32)
    c[0], c[1] = cx, cy
    return c


def postprocess_plane_head_mask(pred_logits, pred_params, pred_mask_logits, query_feat, height, width,
                                plane_score_threshold=0.6, mask_prob_threshold=0.5, overlap_threshold=0.6):
    """Returns one dict per image: pred_plane [n,3], pred_plane_feats [1,n,C], pred_plane_oriIdxs (list of int),
    pred_plane_masks bool [n,H,W], pred_plane_ins_center [n,2], scores (list), bboxes (list of [x,y,w,h]), counts (list
    of RLE strings), areas (list), zero_flag, fallback."""
    bs, nq = pred_logits.shape[:2]
    xy = normalized_xy_map(height, width)
    results = []
    for i in range(bs):
        logits, param = pred_logits[i], pred_params[i]
        prob_hw = torch.sigmoid(pred_mask_logits[i])                                                     # :646
        prob_hw = F.interpolate(prob_hw[:, None], size=(height, width), mode="bilinear", align_corners=False)[:, 0]
        ori_idx = torch.arange(0, nq)
        pred_prob = F.softmax(logits, dim=-1)                                                            # :652
        score, labels = pred_prob.max(dim=-1)
        label_mask = (labels == 0) & (score > plane_score_threshold)                                     # :654
        zero_flag = False
        if int(label_mask.sum()) == 0:                                                                   # :657-661
            _, max_pro_idx = pred_prob[:, 0].max(dim=0)
            label_mask[max_pro_idx] = 1
            score[max_pro_idx] = pred_prob[max_pro_idx, 0]
            zero_flag = True
        valid_param = param[label_mask, :]
        valid_prob = score[label_mask]
        valid_ori = prob_hw[label_mask]
        valid_w = valid_prob.view(-1, 1, 1) * valid_ori                                                  # :667
        valid_feat = query_feat[i, label_mask]
        valid_idx = ori_idx[label_mask]
        ids = valid_w.argmax(0)                                                                          # :674
        out = {k: [] for k in ("plane", "feat", "idx", "mask", "center", "score", "bbox", "counts", "area")}
        max_overlap_id, max_overlap, fallback = 0, 0.0, False

        def emit(pi, mask_np, eps):
            counts = rle_encode(mask_np)
            out["plane"].append(valid_param[pi])
            out["feat"].append(valid_feat[pi])
            out["idx"].append(int(valid_idx[pi]))
            out["mask"].append(torch.from_numpy(np.ascontiguousarray(mask_np)))
            out["center"].append(_center(mask_np, xy, eps))
            out["score"].append(float(valid_prob[pi]))
            out["bbox"].append(rle_to_bbox(counts, height, width))
            out["counts"].append(rle_to_string(counts))
            out["area"].append(int(mask_np.sum()))

        for pi in range(valid_param.shape[0]):                                                           # :684-739
            mask_pi = (ids == pi) & (valid_w[pi] > mask_prob_threshold)
            mask_np = mask_pi.numpy().copy()
            mask_area = int(mask_pi.sum())
            original_area = int((valid_ori[pi] >= mask_prob_threshold).sum())
            if not zero_flag:
                if mask_area < 1 or original_area < 1:
                    continue
                overlap = mask_area / original_area
                if overlap > max_overlap:
                    max_overlap, max_overlap_id = overlap, pi
                if overlap < overlap_threshold:
                    continue
            elif mask_area == 0:
                mask_np[0, 0] = True
            emit(pi, mask_np, 1e-10)
        if len(out["plane"]) == 0:                                                                       # :741-790
            fallback = True
            emit(max_overlap_id, (ids == max_overlap_id).numpy().copy(), 0.0)
        results.append({
            "pred_plane": torch.stack(out["plane"], 0),
            "pred_plane_feats": torch.stack(out["feat"], 0).unsqueeze(0).contiguous(),
            "pred_plane_oriIdxs": out["idx"],
            "pred_plane_masks": torch.stack(out["mask"], 0),
            "pred_plane_ins_center": torch.from_numpy(np.stack(out["center"], 0)).reshape(-1, 2),
            "scores": out["score"], "bboxes": out["bbox"], "counts": out["counts"], "areas": out["area"],
            "zero_flag": zero_flag, "fallback": fallback,
            # margins for conditioning-aware comparisons against a device implementation (not part of the reference's result)
            "_valid_w": valid_w, "_valid_ori": valid_ori, "_valid_idx": valid_idx,
        })
    return results
