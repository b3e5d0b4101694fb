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
