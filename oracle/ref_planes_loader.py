"""TEST INFRASTRUCTURE — not product code.

Runs the reference's own `PlaneTR_NopeSAC._postprocess_planeHeadMask` and `precompute_xy_map`
(meta_arch/siamese_planeTR.py:625-812) without importing the module (it pulls in detectron2 / pycocotools / cv2 model code at
import time): the two methods are cut out of the *unmodified* source file with `ast` and compiled as they are.  Nothing is
copied into the repo.  Harness-side substitutions, all outside the arithmetic under test:
  * `mask_util` (pycocotools, not installed)  -> oracle.planes_restate.mask_util (restated maskApi.c)
  * `np.float` (removed in numpy >= 1.24)     -> float
`available()` says whether the reference is present (it is not on the GPU box).
"""
from __future__ import annotations

import ast
import os
import textwrap
import types

import numpy as np
import torch
import torch.nn.functional as F

from . import planes_restate

REF_ROOT = os.environ.get("NSAC_REFERENCE_ROOT", "/root/reference")
_FILE = os.path.join(REF_ROOT, "NopeSAC_Net", "modeling", "meta_arch", "siamese_planeTR.py")


def available() -> bool:
    return os.path.isfile(_FILE)


class _NumpyWithFloat:
    float = float

    def __getattr__(self, name):
        return getattr(np, name)


def load(num_queries: int, height: int = 480, width: int = 640, plane_score_threshold: float = 0.6,
         mask_prob_threshold: float = 0.5, overlap_threshold: float = 0.6):
    """Returns f(planeTR_outputs, query_feat) -> the reference's list of per-image result dicts."""
    with open(_FILE) as f:
        src = f.read()
    found = {}
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name in ("_postprocess_planeHeadMask", "precompute_xy_map"):
            found.setdefault(node.name, ast.get_source_segment(src, node))
    if len(found) != 2:
        raise RuntimeError("reference post-processing code not found")
    ns = {"np": _NumpyWithFloat(), "torch": torch, "F": F, "mask_util": planes_restate.mask_util}
    for s in found.values():
        exec(compile(textwrap.dedent(s), _FILE, "exec"), ns)
    self = types.SimpleNamespace(num_queries=num_queries, plane_score_threshold=plane_score_threshold,
                                 mask_prob_threshold=mask_prob_threshold, overlap_threshold=overlap_threshold,
                                 device=torch.device("cpu"))
    ns["precompute_xy_map"](self, h=height, w=width)

    def run(planeTR_outputs, query_feat):
        bs = planeTR_outputs["pred_logits"].shape[0]
        batched_inputs = [{"image_id": i, "file_name": f"synthetic_{i}", "height": height, "width": width} for i in range(bs)]
        return ns["_postprocess_planeHeadMask"](self, planeTR_outputs, [None] * bs, batched_inputs, [(height, width)] * bs, query_feat)

    return run
