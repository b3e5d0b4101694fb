"""Seeded synthetic image-pair generator (SURVEY.md §8(d) "Synthetic pair").

Everything is drawn from a CPU ``torch.Generator().manual_seed(20260 + pair_idx)`` in fp32 so the
CUDA path and the CPU oracle see identical bits.  Planted correspondences: view-2 planes are the
view-1 planes warped by a ground-truth pose (the reference's own warp formula,
camera_head.py:1427-1456) plus a little noise, shuffled by a seeded permutation; appearance
embeddings of planted matches differ by 0.05·N(0, I).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

SEED_BASE = 20260
FLIP = (1.0, -1.0, -1.0)  # suncg2habitat axis flip used throughout the reference


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """(w,x,y,z) -> R, same element formulas as camera_head.py:1148-1173."""
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    rows = [
        1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * x * z + 2 * w * y,
        2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x,
        2 * x * z - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y,
    ]
    return torch.stack(rows, dim=-1).reshape(q.shape[:-1] + (3, 3))


def warp_planes(p: torch.Tensor, q: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """Plane params p [n,3] of view 1 expressed in the global frame of pose (q [4], t [3]).
    end = R (p*flip) + t ; b = end - t ; pi = (end.b / (|b|+1e-5)^2) b   (camera_head.py:1446-1453)."""
    flip = torch.tensor(FLIP, dtype=p.dtype)
    R = quat_to_rotmat(q)
    end = (R @ (p * flip).T).T + t
    b = end - t
    k = (end * b).sum(-1) / (b.norm(dim=-1) + 1e-5) ** 2
    return k[:, None] * b


def rotvec_to_quat(rv: torch.Tensor) -> torch.Tensor:
    ang = rv.norm()
    if float(ang) < 1e-12:
        return torch.tensor([1.0, 0.0, 0.0, 0.0])
    axis = rv / ang
    q = torch.cat([torch.cos(ang / 2)[None], axis * torch.sin(ang / 2)])
    if q[0] < 0:
        q = -q
    return q / q.norm()


@dataclass
class PairBatch:
    planes1: torch.Tensor   # [B,P,3]
    planes2: torch.Tensor   # [B,P,3]
    app1: torch.Tensor      # [B,P,256]
    app2: torch.Tensor      # [B,P,256]
    gt_quat: torch.Tensor   # [B,4]  (w,x,y,z), w>=0
    gt_tran: torch.Tensor   # [B,3]
    perm: torch.Tensor      # [B,P] int64: view-1 plane i is planted at view-2 index perm[i]
    feats1: Optional[Dict[str, torch.Tensor]] = None   # res2..res5, [B,C,h,w]
    feats2: Optional[Dict[str, torch.Tensor]] = None

    def to(self, device, non_blocking=False):
        def mv(x):
            if x is None:
                return None
            if isinstance(x, dict):
                return {k: v.to(device, non_blocking=non_blocking) for k, v in x.items()}
            return x.to(device, non_blocking=non_blocking)
        return PairBatch(*[mv(getattr(self, f)) for f in self.__dataclass_fields__])


FEATURE_SHAPES = {"res2": (256, 4), "res3": (512, 8), "res4": (1024, 16), "res5": (2048, 32)}


def make_pair(pair_idx: int, planes_per_view: int = 16, with_features: bool = False,
              height: int = 480, width: int = 640, negative_k: bool = False):
    """One planted pair. `negative_k` pulls every other plane to within a few centimetres of the camera so
    that the (re-embedded) initial translation lies beyond it and some warp factors k = 1 + t.b/|b|^2 are
    negative (sig_seq = -1 coverage, camera_head.py:568-569)."""
    g = torch.Generator().manual_seed(SEED_BASE + pair_idx)
    P = planes_per_view
    n = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=-1)
    d = torch.rand(P, 1, generator=g) * 3.0 + 0.5
    if negative_k:
        d[::2] = d[::2] * 0.02
    p1 = n * d
    rv = (torch.rand(3, generator=g) * 2 - 1) * 0.6
    q = rotvec_to_quat(rv)
    t = (torch.rand(3, generator=g) * 2 - 1) * 0.4
    flip = torch.tensor(FLIP)
    perm = torch.randperm(P, generator=g)
    p2 = torch.empty(P, 3)
    p2[perm] = warp_planes(p1, q, t) * flip + 0.01 * torch.randn(P, 3, generator=g)
    a1 = torch.randn(P, 256, generator=g)
    a2 = torch.empty(P, 256)
    a2[perm] = a1 + 0.05 * torch.randn(P, 256, generator=g)
    feats = [None, None]
    if with_features:
        for v in range(2):
            feats[v] = {
                k: torch.relu(torch.randn(c, height // s, width // s, generator=g))
                for k, (c, s) in FEATURE_SHAPES.items()
            }
    return p1, p2, a1, a2, q, t, perm, feats[0], feats[1]


def make_batch(first_pair: int, num_pairs: int, planes_per_view: int = 16, with_features: bool = False,
               negative_k: bool = False, height: int = 480, width: int = 640) -> PairBatch:
    items = [make_pair(first_pair + i, planes_per_view, with_features, height, width, negative_k)
             for i in range(num_pairs)]
    stack = lambda j: torch.stack([it[j] for it in items], 0)
    f1 = f2 = None
    if with_features:
        f1 = {k: torch.stack([it[7][k] for it in items], 0) for k in FEATURE_SHAPES}
        f2 = {k: torch.stack([it[8][k] for it in items], 0) for k in FEATURE_SHAPES}
    return PairBatch(stack(0), stack(1), stack(2), stack(3), stack(4), stack(5), stack(6), f1, f2)


def device_features(num_pairs: int, device, seed: int = 7, height: int = 480, width: int = 640):
    """Backbone feature maps drawn directly on the device (bench only: 34 MB/pair is too much to
    draw on the host for B=64; parity tests use make_batch(with_features=True))."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = []
    for _ in range(2):
        out.append({
            k: torch.relu(torch.randn(num_pairs, c, height // s, width // s, generator=g, device=device))
            for k, (c, s) in FEATURE_SHAPES.items() if k != "res2"
        })
        # res2 is dropped by the pixel decoder (camera_modules.py:262-263); keep a 1-element stand-in
        out[-1]["res2"] = torch.zeros(num_pairs, 1, 1, 1, device=device)
    return out[0], out[1]


def all_pairs_hypotheses(planes_per_view: int, num_hyp: int, seed: int = 0) -> torch.Tensor:
    """Index list [H,2] of (i,j) candidate plane pairs in torch.nonzero (row-major) order
    (SURVEY.md §8(d) "Mapping P planes x H hypotheses"): H == P*P -> all pairs; H < P*P -> the first
    H of a seeded permutation, sorted row-major; H > P*P -> all pairs repeated cyclically."""
    P = planes_per_view
    grid = torch.stack(torch.meshgrid(torch.arange(P), torch.arange(P), indexing="ij"), -1).reshape(-1, 2)
    if num_hyp == P * P:
        return grid
    if num_hyp < P * P:
        g = torch.Generator().manual_seed(977 + seed)
        sel = torch.randperm(P * P, generator=g)[:num_hyp].sort().values
        return grid[sel]
    reps = math.ceil(num_hyp / (P * P))
    return grid.repeat(reps, 1)[:num_hyp]


def pose_errors(pred_tran, pred_rot, gt_tran, gt_rot):
    """Reference pose-error formulas (mp3d_evaluation.py:389-391, 463-465):
    T = |t - t_gt|_2 ; R = 2 acos(clip(|q.q_gt|, -1, 1)) * 180/pi."""
    te = (pred_tran - gt_tran).norm(dim=-1)
    dot = (torch.nn.functional.normalize(pred_rot, dim=-1) * gt_rot).sum(-1).abs().clamp(-1, 1)
    re = 2 * torch.acos(dot) * 180.0 / math.pi
    return te, re


def make_weights(shapes, seed: int = 40):
    """Deterministic, reference-independent weights for parity tests and benches: name -> tensor for a
    {name: shape} spec (a module's state_dict shapes).  He-uniform matrices (keeps ReLU chains O(1) through
    the ~28-layer refinement MLPs so errors are not hidden by vanishing activations), small non-zero biases,
    non-trivial norm statistics.  Drawn in sorted-name order from one CPU generator."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t = torch.zeros(shape, dtype=torch.int64)
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif name == "bin_score" or name.endswith(".bin_score"):
            t = torch.ones(shape)
        elif len(shape) <= 1 and leaf == "weight":      # norm scales
            t = torch.rand(shape, generator=g) + 0.5
        elif leaf == "bias":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = math.sqrt(6.0 / max(fan_in, 1))
            if name.endswith("trans.weight") and "convs" not in name and "fc_" not in name and "decoder" not in name:
                bound *= 0.1       # `trans` head: keep regressed translations in the metre range (abs 1e-4 bar)
            if name.endswith("convs_backbone.7.0.weight"):
                bound *= 0.05      # keeps the 300-way correlation-softmax logits O(10) (He gain gives ~3500 and a
                                   # softmax whose fp32-vs-fp64 noise alone is 3e-5 on the initial pose)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        out[name] = t
    return out


def make_backbone_weights(shapes, seed: int = 8, conv3_scale: float = 0.3):
    """Seeded random-init state of the R-50 backbone (BASELINE.json configs[1]: "random-init ResNet50") for a {name: shape} spec
    (`ResNet50Backbone().state_dict()` shapes): He-normal (fan_out) convolutions like c2_msra_fill, FrozenBN statistics drawn
    non-trivially, and the last norm scale of every block damped by `conv3_scale` so that 16 residual blocks of a RANDOM network
    keep res5 O(1) like a trained one (otherwise activations grow ~2x per block and the pixel network's correlation softmax
    saturates).  Both arms of the bench load this same state."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        if name.endswith("running_var"):
            t = torch.rand(shape, generator=g) + 0.5
        elif name.endswith("norm.weight"):
            t = torch.rand(shape, generator=g) * 0.5 + 0.5
            if name.endswith("conv3.norm.weight"):
                t = t * conv3_scale
        elif name.endswith("norm.bias") or name.endswith("running_mean"):
            t = torch.randn(shape, generator=g) * 0.1
        else:
            t = torch.randn(shape, generator=g) * (2.0 / (shape[0] * shape[2] * shape[3])) ** 0.5
        out[name] = t
    return out


def make_images(seed: int, num_images: int, height: int = 480, width: int = 640) -> torch.Tensor:
    """Synthetic RGB, uint8 U{0..255} [N,3,H,W] (SURVEY.md 8d "RGB (full-model benches)"), seeded on the host."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (num_images, 3, height, width), generator=g, dtype=torch.uint8)


# ---------------------------------------------------------------------------------------------------------------------
# Row f1: synthetic PlaneTRHead outputs (the inputs of `_postprocess_planeHeadMask`, siamese_planeTR.py:625-803)
# ---------------------------------------------------------------------------------------------------------------------
PLANE_HEAD_CASES = ("regular", "zero", "zero_empty", "fallback_empty", "fallback_overlap", "noise")


def make_plane_head_outputs(image_idx: int, num_queries: int = 50, mask_h: int = 120, mask_w: int = 160,
                            channels: int = 256, case: str = "regular", planes: Optional[int] = None):
    """One image's `pred_logits [NQ,2]`, `pred_params [NQ,3]`, `pred_mask_logits [NQ,h,w]`, `query_feat [NQ,C]` (CPU fp32,
    seeded by the image index).  Plane queries get smooth blob-shaped mask logits that tile the image; the cases drive the
    branches of the reference's post-processing:
      regular           several confident planes (+ two decoys: a duplicate that loses every pixel, a low-overlap one)
      zero              no query passes the plane-score test            -> `zero_flag` (:657-661)
      zero_empty        as above and its thresholded mask is empty      -> pixel (0,0) is set (:699-702)
      fallback_empty    one plane, thresholded mask empty               -> `len(instances) == 0` branch (:741-790)
      fallback_overlap  one plane, overlap below OVERLAP_THRESHOLD      -> same branch, non-trivial overlap
      noise             noisy patchwork masks (ragged regions, many near-ties between planes)
    """
    assert case in PLANE_HEAD_CASES
    g = torch.Generator().manual_seed(77000 + image_idx)
    nq = num_queries
    ys = torch.arange(mask_h, dtype=torch.float32).view(-1, 1) / mask_h
    xs = torch.arange(mask_w, dtype=torch.float32).view(1, -1) / mask_w
    logits = torch.empty(nq, 2)
    logits[:, 0] = -3.0 + 0.5 * torch.randn(nq, generator=g)
    logits[:, 1] = 3.0 + 0.5 * torch.randn(nq, generator=g)
    masks = -4.0 + 0.5 * torch.randn(nq, mask_h, mask_w, generator=g)
    params = torch.randn(nq, 3, generator=g)
    feats = torch.randn(nq, channels, generator=g)
    if planes is None:
        planes = 4 + int(torch.randint(0, 13, (1,), generator=g))
    planes = min(planes, nq)
    order = torch.randperm(nq, generator=g)

    def blob(scale=14.0, sharp=40.0):
        cy, cx = torch.rand(2, generator=g).tolist()
        rad = 0.10 + 0.12 * float(torch.rand(1, generator=g))
        d2 = ((ys - cy) ** 2 + (xs - cx) ** 2) / (rad * rad)
        return scale - sharp * d2 / 4.0 + 0.3 * torch.randn(mask_h, mask_w, generator=g)

    def patches():
        coarse = 4.0 * torch.randn(1, 1, mask_h // 8, mask_w // 8, generator=g)
        up = torch.nn.functional.interpolate(coarse, size=(mask_h, mask_w), mode="bilinear", align_corners=False)[0, 0]
        return up + 1.5 * torch.randn(mask_h, mask_w, generator=g)

    if case in ("regular", "noise"):
        for k in range(planes):
            q = int(order[k])
            logits[q, 0] = 2.0 + 2.0 * float(torch.rand(1, generator=g))
            logits[q, 1] = -1.0 + 0.5 * float(torch.randn(1, generator=g))
            masks[q] = blob() if case == "regular" else patches()
        if case == "regular" and planes + 2 <= nq:
            dup, weak = int(order[planes]), int(order[planes + 1])
            logits[dup] = torch.tensor([0.6, -0.6])           # passes the score test, loses every pixel to its twin
            masks[dup] = masks[int(order[0])] - 1.0
            logits[weak] = torch.tensor([0.3, -0.3])          # score ~0.65: most of its area falls below the mask threshold
            masks[weak] = blob(scale=1.2, sharp=2.0)
    elif case in ("zero", "zero_empty"):
        q = int(order[0])
        logits[q] = torch.tensor([0.1, 0.0])                  # best p0 ~0.52 < PLANE_SCORE_THRESHOLD
        masks[q] = blob() if case == "zero" else -3.0 + 0.2 * torch.randn(mask_h, mask_w, generator=g)
    elif case == "fallback_empty":
        q = int(order[0])
        logits[q] = torch.tensor([0.25, -0.25])               # score ~0.62
        masks[q] = 0.7 + 0.1 * torch.randn(mask_h, mask_w, generator=g)      # prob ~0.67: 0.62 * 0.67 < 0.5 everywhere
    elif case == "fallback_overlap":
        q = int(order[0])
        logits[q] = torch.tensor([0.4, -0.4])                 # score ~0.69
        masks[q] = blob(scale=1.5, sharp=3.0)                 # prob mostly 0.5 .. 0.8: area(0.69 p > 0.5) << area(p >= 0.5)
    return {"pred_logits": logits, "pred_params": params, "pred_mask_logits": masks, "query_feat": feats}


def make_plane_head_batch(first_image: int, num_images: int, cases=("regular",), **kw):
    """Stacked `[B, ...]` tensors of `make_plane_head_outputs` (case i = cases[i % len(cases)])."""
    items = [make_plane_head_outputs(first_image + i, case=cases[i % len(cases)], **kw) for i in range(num_images)]
    return {k: torch.stack([it[k] for it in items]).contiguous() for k in items[0]}
